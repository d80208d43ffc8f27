"""emdee.jl_b200 -- B200-native (sm_100a) implementation of EmDee.jl's nonbonded hot path behind the
reference's own interface: cell binning -> cutoff pair loop -> Lennard-Jones energy/force/virial
(-> velocity-Verlet).  Import as `emdee_jl_b200` (see ../emdee_jl_b200.py).

Layout: csrc/ holds the CUDA kernels and the C ABI (include/emdee_b200.h); api.py mirrors the
reference's exported names over that ABI; workloads.py generates the BASELINE configurations.
Loading fails loudly when the CUDA library has not been built: there is no CPU fallback.
"""
from ._lib import EmDeeError, build_library, LIB_PATH  # noqa: F401
from .api import (  # noqa: F401
    ALLPAIRS_REFERENCE,
    CUTOFF,
    ENERGIES,
    FORCES,
    VIRIALS,
    WARPSIZE,
    Cells,
    berendsen_,
    cells_per_dimension,
    Context,
    LennardJonesAtom,
    LennardJonesModel,
    NonbondedSystem,
    comm_unique_id,
    compute_nonbonded_,
    default_context,
    index2voxel,
    naively_compute_nonbonded_,
    nonbonded_computation_tiles,
    step_,
    stencil_vectors,
    surrounding_cells,
    update_cells_,
)
from . import modelling, trajectory, workloads  # noqa: F401
