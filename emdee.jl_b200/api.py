"""api.py -- host-side mirror of EmDee.jl's interface for the nonbonded path.

Julia is not installed here, so the host side above the C ABI is Python + ctypes and
mirrors the reference's exported names, argument order and meaning one to one; julia/EmDee.jl holds
the equivalent `ccall` shim a maintainer would drop into the reference.  Differences forced by the
language: `f!` is spelled `f_`, `Val(bitmask)` is a plain int, arrays are numpy (a 3xN column-major
Julia matrix is an (N,3) C-contiguous numpy array; (3,N) Fortran-ordered arrays are accepted too).

Everything computes on the GPU through libemdee_b200.so; nothing here falls back to the CPU.
Reference citations are file:line relative to the reference tree.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import EmDeeError, call

# src/nonbonded.jl:12-16
FORCES = 1 << 0
ENERGIES = 1 << 1
VIRIALS = 1 << 2
WARPSIZE = 32

# pair-set semantics
CUTOFF = 0
ALLPAIRS_REFERENCE = 1


class LennardJonesModel:
    """LennardJonesModel(cutoff, switch) -- src/lennard_jones.jl:6-11.
    Fields rc2, rs2, inv_delta2 are the reference's rc², rs², δ⁻² (FP64 here, Float32 there)."""

    def __init__(self, cutoff, switch):
        cutoff, switch = float(cutoff), float(switch)
        if not (0.0 < switch < cutoff):
            raise ValueError("LennardJonesModel: need 0 < switch < cutoff")
        self.cutoff, self.switch = cutoff, switch
        self.rc2 = cutoff ** 2
        self.rs2 = switch ** 2
        self.inv_delta2 = 1.0 / (cutoff ** 2 - switch ** 2)

    def __repr__(self):
        return "LennardJonesModel(%r, %r)" % (self.cutoff, self.switch)


def LennardJonesAtom(eps, sigma):
    """LennardJonesAtom(ε, σ) = LJAtom(0.5σ, 2*sqrt(ε)) -- src/lennard_jones.jl:13,15-18.
    Argument order is (ε, σ).  Returns the (half_σ, twice_sqrt_ε) pair; a vector of atoms is an
    (N,2) float64 array (`np.tile(LennardJonesAtom(1, 1), (N, 1))` for `fill(LennardJonesAtom(1,1), N)`)."""
    return np.array([0.5 * float(sigma), 2.0 * math.sqrt(float(eps))])


def nonbonded_computation_tiles(N):
    """nonbonded_computation_tiles(N) -- src/nonbonded.jl:18-26: (I,J) 1-based block pairs, I<=J,
    ordered by diagonal offset.  Integer work: identical to the reference."""
    n = -(-int(N) // WARPSIZE)
    out = np.empty((n * (n + 1) // 2, 2), dtype=np.int32)
    k = 0
    for i in range(n):
        m = n - i
        j = np.arange(1, m + 1, dtype=np.int32)
        out[k:k + m, 0] = j
        out[k:k + m, 1] = j + i
        k += m
    return out


def _as_3xN(a, name):
    """Accept (N,3) C-order or (3,N) F-order float64; return a C-contiguous (N,3) view/copy."""
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError("%s must be a 3xN matrix" % name)
    if a.shape[0] == 3 and a.shape[1] != 3:
        a = a.T
    elif a.shape[0] == 3 and a.shape[1] == 3 and a.flags.f_contiguous and not a.flags.c_contiguous:
        a = a.T
    if a.shape[1] != 3:
        raise ValueError("%s must be a 3xN matrix" % name)
    return np.ascontiguousarray(a, dtype=np.float64)


def _out_view(a, N, width, name):
    """Writable C-contiguous float64 view of a caller-owned output array (3xN or length N)."""
    if not isinstance(a, np.ndarray) or a.dtype != np.float64:
        raise TypeError("%s must be a float64 numpy array" % name)
    if width == 1:
        if a.shape != (N,) or not a.flags.c_contiguous:
            raise ValueError("%s must be a contiguous vector of length N" % name)
        return a
    if a.shape == (N, 3) and a.flags.c_contiguous:
        return a
    if a.shape == (3, N) and a.flags.f_contiguous:
        return a.T
    raise ValueError("%s must be (N,3) C-contiguous or (3,N) Fortran-contiguous" % name)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU (emdee_create).  `device` defaults to LOCAL_RANK-style index 0."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        call("emdee_create", C.byref(self._h), int(device))
        self.device = int(device)

    def close(self):
        if self._h:
            _lib.load().emdee_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        call("emdee_device_info", self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem))
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), mem_bytes=mem.value)

    def measure_fp64_peak(self):
        """Measured DFMA throughput in FLOP/s (2 per FMA) -- the FP64 roofline denominator."""
        f, ms = C.c_double(), C.c_double()
        call("emdee_measure_fp64_peak", self._h, C.byref(f), C.byref(ms))
        return f.value

    def launch_count(self):
        n = C.c_int64()
        call("emdee_launch_count", self._h, C.byref(n))
        return n.value

    def timer_start(self):
        call("emdee_timer_start", self._h)

    def timer_stop(self):
        ms = C.c_double()
        call("emdee_timer_stop", self._h, C.byref(ms))
        return ms.value

    def comm_init(self, rank, nranks, unique_id):
        call("emdee_comm_init", self._h, int(rank), int(nranks), unique_id)


def comm_unique_id():
    buf = C.create_string_buffer(128)
    call("emdee_comm_unique_id", buf)
    return buf.raw


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class NonbondedSystem:
    """Device-resident system handle (emdee_system_*): positions, velocities and forces stay on the
    GPU across calls (form ii).  Additive to the reference API, which has no
    simulation loop."""

    def __init__(self, N, L, ctx=None):
        self.ctx = ctx or default_context()
        self.N, self.L = int(N), float(L)
        self._h = C.c_void_p()
        call("emdee_system_create", self.ctx._h, self.N, self.L, C.byref(self._h))

    def close(self):
        if self._h:
            _lib.load().emdee_system_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setters ---------------------------------------------------------------------------------
    def set_model(self, model):
        call("emdee_set_model", self._h, model.cutoff, model.switch)
        self.model = model

    def set_atoms(self, atoms):
        a = np.ascontiguousarray(atoms, dtype=np.float64)
        if a.shape != (self.N, 2):
            raise ValueError("atoms must be an (N,2) array of (half_sigma, twice_sqrt_eps)")
        call("emdee_set_lj_atoms", self._h, _ptr(a))

    def set_positions(self, positions):
        p = _as_3xN(positions, "positions")
        if p.shape[0] != self.N:
            raise ValueError("positions must hold N atoms")
        call("emdee_set_positions", self._h, _ptr(p))

    def update_cells(self):
        """Incremental form of update_cells! (src/cells.jl:196-222) after set_positions: returns the number of atoms whose cell
        changed; 0 means the sorted order, the cell table (and, within skin/2, the pair list) were kept as they are."""
        m = C.c_int64()
        call("emdee_update_cells", self._h, C.byref(m))
        return int(m.value)

    def local_id_range(self):
        """(first id, count): the smallest cyclic window of global ids -- first, first+1, ... (mod N) -- covering every atom this
        rank owns."""
        a, n = C.c_int64(), C.c_int64()
        call("emdee_get_local_id_range", self._h, C.byref(a), C.byref(n))
        return int(a.value), int(n.value)

    def set_positions_range(self, id_first, count, positions):
        """Upload only rows id_first, ... (mod N), `count` of them, of the full (N, 3) position array (a slab rank moves the
        rows of the atoms it owns; bin() refreshes the ghosts)."""
        p = _as_3xN(positions, "positions")
        if p.shape[0] != self.N:
            raise ValueError("positions must hold N atoms")
        call("emdee_set_positions_range", self._h, int(id_first), int(count), _ptr(p))

    def _range_out(self, out, shape):
        if not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous and out.shape == shape):
            raise ValueError("out must be a C-contiguous float64 array of shape %r (the full id-ordered array)" % (shape,))
        return out

    def forces_range(self, id_first, count, out):
        """Write the window's rows of the full (N, 3) array `out` (forces of the atoms this rank owns); other rows are untouched."""
        call("emdee_get_forces_range", self._h, int(id_first), int(count), _ptr(self._range_out(out, (self.N, 3))))
        return out

    def energies_range(self, id_first, count, out):
        call("emdee_get_energies_range", self._h, int(id_first), int(count), _ptr(self._range_out(out, (self.N,))))
        return out

    def virials_range(self, id_first, count, out):
        call("emdee_get_virials_range", self._h, int(id_first), int(count), _ptr(self._range_out(out, (self.N,))))
        return out

    def set_velocities(self, velocities):
        v = _as_3xN(velocities, "velocities")
        call("emdee_set_velocities", self._h, _ptr(v))

    def set_masses(self, masses):
        m = np.ascontiguousarray(masses, dtype=np.float64)
        if m.shape != (self.N,):
            raise ValueError("masses must have length N")
        call("emdee_set_masses", self._h, _ptr(m))

    def set_exclusions(self, base, mask):
        if base is None:
            call("emdee_set_exclusions", self._h, None, None)
            return
        b = np.ascontiguousarray(base, dtype=np.int32)
        m = np.ascontiguousarray(mask, dtype=np.uint64)
        call("emdee_set_exclusions", self._h, _ptr(b), _ptr(m))

    def set_pairs14(self, pairs, scale):
        """Pairs three bonds apart (workloads.pairs14) interact with `scale` (the force field's lj14scale, src/modelling.jl:199)
        times the ordinary Lennard-Jones interaction.  pairs=None clears."""
        if pairs is None:
            call("emdee_set_pairs14", self._h, None, 0, 1.0)
            return
        ij = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        call("emdee_set_pairs14", self._h, _ptr(ij), ij.shape[0], float(scale))

    def set_skin(self, skin):
        call("emdee_set_skin", self._h, float(skin))

    def set_tiles(self, tiles):
        if tiles is None:
            call("emdee_set_tiles", self._h, None, 0)
            return
        t = np.ascontiguousarray(tiles, dtype=np.int32).reshape(-1, 2)
        call("emdee_set_tiles", self._h, _ptr(t), t.shape[0])

    # -- compute ---------------------------------------------------------------------------------
    def bin(self, ndiv=2):
        call("emdee_bin", self._h, int(ndiv))

    def compute(self, mode=CUTOFF, bitmask=FORCES | ENERGIES | VIRIALS):
        call("emdee_compute_nonbonded", self._h, int(mode), int(bitmask))

    def compute_into(self, mode, bitmask, forces=None, energies=None, virials=None):
        """compute_nonbonded!(forces, energies, virials, ...) with the reference's output arguments (src/nonbonded.jl:109-120):
        the evaluation and the transfer of the selected outputs into caller-owned id-ordered host arrays as ONE call
        (emdee_compute_nonbonded_into: on one GPU the rows of a finished chunk of z planes travel while the next chunk computes)."""
        f = _out_view(forces, self.N, 3, "forces") if bitmask & FORCES else None
        e = _out_view(energies, self.N, 1, "energies") if bitmask & ENERGIES else None
        w = _out_view(virials, self.N, 1, "virials") if bitmask & VIRIALS else None
        call("emdee_compute_nonbonded_into", self._h, int(mode), int(bitmask), _ptr(f), _ptr(e), _ptr(w))

    def vv_step(self, dt, nsteps, rebin_every=1):
        """nsteps velocity-Verlet steps; rebin_every > 0: re-bin at that cadence, 0: never, < 0 (with a skin):
        adaptively, when an atom has moved more than skin/2 since the last binning."""
        call("emdee_vv_step", self._h, float(dt), int(nsteps), int(rebin_every))

    def synchronize(self):
        call("emdee_synchronize", self._h)

    def profile_begin(self):
        call("emdee_profile_begin", self._h)

    def profile_end(self):
        """(summed force-kernel milliseconds, number of force-kernel launches) since profile_begin."""
        ms, n = C.c_double(), C.c_int64()
        call("emdee_profile_end", self._h, C.byref(ms), C.byref(n))
        return ms.value, n.value

    def profile_kind(self, kind):
        """(milliseconds, launches) of one kernel in the last profile: 0 k_force_cells, 1 k_list_build, 2 k_force_list."""
        ms, n = C.c_double(), C.c_int64()
        call("emdee_profile_kind", self._h, int(kind), C.byref(ms), C.byref(n))
        return ms.value, n.value

    # -- getters ---------------------------------------------------------------------------------
    def _get3(self, fn, out=None):
        out = np.empty((self.N, 3)) if out is None else out
        call(fn, self._h, _ptr(out))
        return out

    def _get1(self, fn, out=None, dtype=np.float64, n=None):
        out = np.empty(self.N if n is None else n, dtype=dtype) if out is None else out
        call(fn, self._h, _ptr(out))
        return out

    def positions(self, out=None):
        return self._get3("emdee_get_positions", out)

    def velocities(self, out=None):
        return self._get3("emdee_get_velocities", out)

    def forces(self, out=None):
        return self._get3("emdee_get_forces", out)

    def energies(self, out=None):
        return self._get1("emdee_get_energies", out)

    def virials(self, out=None):
        return self._get1("emdee_get_virials", out)

    def totals(self, pairs=True):
        E, W, n = C.c_double(), C.c_double(), C.c_int64()
        call("emdee_get_totals", self._h, C.byref(E), C.byref(W), C.byref(n) if pairs else None)
        return E.value, W.value, (n.value if pairs else None)

    def kinetic_energy(self):
        K = C.c_double()
        call("emdee_kinetic_energy", self._h, C.byref(K))
        return K.value

    def step_config(self):
        """How the stepping path is configured after the last bin(): brick shape, kernel variant (host-only query)."""
        o = np.zeros(8, dtype=np.int32)
        call("emdee_get_step_config", self._h, _ptr(o))
        return dict(brick=(int(o[0]), int(o[1]), int(o[2])), brick_capacity=int(o[3]), pair_list=bool(o[4]),
                    persistent=bool(o[5] & 1), tma=bool(o[5] & 2), compacted=bool(o[5] & 4), shallow_stacks=bool(o[5] & 8), split_lists=bool(o[5] & 16),
                    fused_vv=bool(o[6]), list_chunks=int(o[7]))

    def step_counters(self):
        """Counters of the fused stepping loop: re-binnings, and stepping launches by list mode (full walk, prune, replay of the
        inner list)."""
        o = np.zeros(4, dtype=np.int64)
        call("emdee_get_step_counters", self._h, _ptr(o))
        return dict(rebins=int(o[0]), walk=int(o[1]), prune=int(o[2]), replay=int(o[3]))

    def scale_velocities(self, factor):
        call("emdee_scale_velocities", self._h, float(factor))

    def checkpoint(self):
        """Host copy of everything a restart needs, in atom-id order (additive)."""
        self.synchronize()      # raises if a device-side check failed during the steps (skin violated, list overflow): never checkpoint such a state
        return dict(N=self.N, L=self.L, positions=self.positions(), velocities=self.velocities())

    def restore(self, ckpt):
        """Positions and velocities from checkpoint(); the caller re-bins and re-evaluates the forces
        (bin(); compute(CUTOFF, FORCES)) before stepping on."""
        if int(ckpt["N"]) != self.N or float(ckpt["L"]) != self.L:
            raise ValueError("checkpoint belongs to a different system (N or L differ)")
        self.set_positions(ckpt["positions"])
        self.set_velocities(ckpt["velocities"])

    def cells_per_dimension(self):
        M = C.c_int32()
        call("emdee_get_cells_per_dimension", self._h, C.byref(M))
        return M.value

    def cell_index(self):
        return self._get1("emdee_get_cell_index", dtype=np.int32)

    def cell_population(self):
        M = self.cells_per_dimension()
        return self._get1("emdee_get_cell_population", dtype=np.int32, n=M ** 3)

    def cell_order(self):
        M = self.cells_per_dimension()
        nloc, _ = self.local_count()
        perm = np.empty(nloc, dtype=np.int32)
        start = np.empty(M ** 3 + 1, dtype=np.int32)
        call("emdee_get_cell_order", self._h, _ptr(perm), _ptr(start))
        return perm, start

    def local_count(self):
        a, b = C.c_int64(), C.c_int64()
        call("emdee_get_local_count", self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def local_ids(self):
        n, _ = self.local_count()
        ids = np.empty(n, dtype=np.int32)
        call("emdee_get_local_ids", self._h, _ptr(ids))
        return ids

    def pair_set(self, cap=None):
        """Sorted (i<j) pair list of the CUTOFF pair set, 0-based ids (small N)."""
        if cap is None:
            cap = int(self.pair_set_digest()[0])
        ij = np.empty((max(cap, 1), 2), dtype=np.int32)
        n = C.c_int64()
        call("emdee_pair_set", self._h, _ptr(ij), cap, C.byref(n))
        return ij[:n.value]

    def pair_set_digest(self):
        d = np.zeros(3, dtype=np.uint64)
        call("emdee_pair_set_digest", self._h, _ptr(d))
        return d

    def list_pair_count(self):
        """Pairs (i<j) the pair-list stepping kernel evaluates inside the cutoff at the current positions
        (-1 when the system steps without a list).  Audit of the list pre-culls."""
        n = C.c_int64()
        call("emdee_list_pair_count", self._h, C.byref(n))
        return int(n.value)


def _system_for(positions, L, model, atoms, ctx=None):
    p = _as_3xN(positions, "positions")
    s = NonbondedSystem(p.shape[0], L, ctx)
    s.set_model(model)
    s.set_atoms(atoms)
    s.set_positions(p)
    return s


def compute_nonbonded_(forces, energies, virials, positions, L, tiles, model, atoms, bitmask,
                       mode=ALLPAIRS_REFERENCE, ndiv=2):
    """compute_nonbonded!(forces, energies, virials, positions, L, tiles, model, atoms, Val(bitmask))
    -- src/nonbonded.jl:109-120.  Caller-owned outputs are overwritten for the selected bits only
    (the reference zeroes and accumulates the selected ones, :112-114).  The default mode is the
    reference's own semantics: every minimum-image pair in `tiles`, no cutoff cull.
    mode=CUTOFF evaluates the pair set r² <= rc² through the cell list and ignores `tiles`."""
    s = _system_for(positions, L, model, atoms)
    try:
        N = s.N
        if mode == ALLPAIRS_REFERENCE:
            s.set_tiles(tiles)
        else:
            s.bin(ndiv)
        s.compute_into(mode, bitmask, forces, energies, virials)
    finally:
        s.close()
    return None


def naively_compute_nonbonded_(forces, energies, virials, positions, L, model, atoms):
    """naively_compute_nonbonded!(forces, energies, virials, positions, L, model, atoms)
    -- src/nonbonded.jl:122-155: all N(N-1)/2 pairs, always all three outputs, no tiles, no bitmask.
    The reference runs this checker as a serial CPU loop; the drop-in evaluates the same sums on
    the GPU (the CPU restatement of the loop lives in oracle/, as test infrastructure)."""
    compute_nonbonded_(forces, energies, virials, positions, L, None, model, atoms,
                       FORCES | ENERGIES | VIRIALS, mode=ALLPAIRS_REFERENCE)


def index2voxel(index, M):
    """index2voxel(index, M) -- src/cells.jl:22-26: 1-based linear cell index -> 0-based voxel [i, j, k], x fastest."""
    k, l = divmod(int(index) - 1, M * M)
    j, i = divmod(l, M)
    return np.array([i, j, k], dtype=np.int64)


def stencil_vectors(rc, action):
    """stencil_vectors(rc, action) -- src/cells.jl:28-34, restated literally (integer work, identical output and order):
    the (2*ceil(rc)+1)^3 block around a cell is split by linear order into the first half (`action`) and the second
    half (reaction), the centre excluded, and filtered by sum((|d|-1)^2) < rc^2 with rc in cell units.  Note that a
    zero component contributes (0-1)^2 = 1 to that sum, as in the reference (it drops needed cells
    for rc* < sqrt(3); the kernels here do not use these tables)."""
    nmax = int(math.ceil(rc))
    M = 1 + 2 * nmax
    half = M ** 3 // 2
    rng = range(1, half + 1) if action else range(half + 2, M ** 3 + 1)
    out = []
    for i in rng:
        v = index2voxel(i, M) - nmax
        if int(((np.abs(v) - 1) ** 2).sum()) < rc * rc:
            out.append(v)
    return np.array(out, dtype=np.int64).reshape(-1, 3)


def cells_per_dimension(L, cutoff, ndiv):
    """cells_per_dimension(L, cutoff, ndiv) = floor(Int32, ndiv*L/cutoff) -- src/cells.jl:36."""
    return int(np.int32(math.floor(ndiv * L / cutoff)))


def surrounding_cells(L, cutoff, M, action):
    """surrounding_cells(L, cutoff, M, action) -- src/cells.jl:38-44: (n_vec, M^3) Int32 table of the 1-based cells
    reached from every cell by the stencil vectors, one periodic image (:39).  Host table for small grids only:
    62 x M^3 x 4 B is 597 MB at 4 M atoms, which is why the kernels derive neighbours from cell coordinates instead."""
    M = int(M)
    vectors = stencil_vectors(M * cutoff / L, action)
    idx = np.arange(M ** 3, dtype=np.int64)
    vox = np.stack([idx % M, (idx // M) % M, idx // (M * M)], axis=1)                  # index2voxel of every cell
    t = vox[None, :, :] + vectors[:, None, :]
    t = np.where(t < 0, t + M, np.where(t >= M, t - M, t))                              # pbc(x), one image only
    return (1 + t[..., 0] + M * t[..., 1] + M * M * t[..., 2]).astype(np.int32)


class Cells:
    """Cells(r, L, cutoff; ndiv=2, num_threads=256) -- src/cells.jl:6-20,176-194.

    Fields kept from the reference: M, cutoff, index (1-based cell of every atom, :85),
    population, head/next (linked lists in descending atom order as distribute! builds them,
    :46-60; derived on demand from the sorted order).  The stencil tables action_cells /
    reaction_cells are not materialised (597 MB at 4M atoms; the kernels derive neighbours from
    the cell coordinates).  num_threads is accepted for signature compatibility and unused."""

    def __init__(self, r, L, cutoff, ndiv=2, num_threads=256, ctx=None):
        p = _as_3xN(r, "r")
        self.cutoff = float(cutoff)
        self.ndiv = int(ndiv)
        self.num_threads = int(num_threads)
        self._sys = NonbondedSystem(p.shape[0], L, ctx)
        # only the cutoff matters for binning; the switch distance is irrelevant here
        self._sys.set_model(LennardJonesModel(self.cutoff, 0.5 * self.cutoff))
        self._refresh(p)

    def _refresh(self, p, incremental=False):
        self._sys.set_positions(p)
        if incremental:
            self.movers = self._sys.update_cells()      # atoms that changed cell (src/cells.jl:79-85); 0: nothing to relink
            if self.movers == 0:
                return
        else:
            self._sys.bin(self.ndiv)
        self.M = self._sys.cells_per_dimension()
        self.index = self._sys.cell_index()
        self.population = self._sys.cell_population()
        self.perm, self.cell_start = self._sys.cell_order()
        self._lists = None

    def _linked_lists(self):
        if self._lists is None:
            N = self.index.shape[0]
            head = np.zeros(self.M ** 3, dtype=np.int32)
            nxt = np.zeros(N, dtype=np.int32)
            ids1 = self.perm + 1                       # 1-based atom ids sorted by (cell, id)
            cs = self.cell_start
            nonempty = np.nonzero(cs[1:] > cs[:-1])[0]
            head[nonempty] = ids1[cs[nonempty + 1] - 1]          # largest id of the cell
            prev = np.zeros(N, dtype=np.int32)
            prev[1:] = ids1[:-1]
            prev[cs[nonempty]] = 0                                # first of each cell ends the list
            nxt[ids1 - 1] = prev
            self._lists = (head, nxt)
        return self._lists

    @property
    def action_cells(self):
        """The reference's stencil table (src/cells.jl:185), built on the host on demand (small grids only)."""
        return surrounding_cells(self._sys.L, self.cutoff, self.M, True)

    @property
    def reaction_cells(self):
        return surrounding_cells(self._sys.L, self.cutoff, self.M, False)

    @property
    def head(self):
        return self._linked_lists()[0]

    @property
    def next(self):
        return self._linked_lists()[1]

    def close(self):
        self._sys.close()


def update_cells_(cells, r, L):
    """update_cells!(cells, r, L) -- src/cells.jl:196-222: re-bin after the atoms moved.  Incremental like the reference's: the
    cell of every atom is recomputed on the device and only if some atom changed cell are the lists rebuilt
    (cells.movers holds the count)."""
    if float(L) != cells._sys.L:
        raise ValueError("update_cells!: L differs from the box the cells were built for")
    cells._refresh(_as_3xN(r, "r"), incremental=True)
    return None


def berendsen_(system, kT, tau, elapsed, ndof=None):
    """Additive: one Berendsen velocity rescaling towards temperature kT (energy units) with coupling time tau after
    `elapsed` time of dynamics: lambda = sqrt(1 + elapsed/tau (kT/kT_now - 1)), kT_now = 2K/ndof (ndof = 3N - 3).
    K comes from the device (emdee_kinetic_energy), the scaling runs on the device (emdee_scale_velocities).
    Returns (kT_now, lambda)."""
    ndof = 3 * system.N - 3 if ndof is None else ndof
    now = 2.0 * system.kinetic_energy() / ndof
    lam = math.sqrt(max(0.0, 1.0 + elapsed / tau * (kT / now - 1.0))) if now > 0 else 1.0
    system.scale_velocities(lam)
    return now, lam


def step_(system, nsteps, dt, rebin_every=1):
    """Additive: velocity-Verlet on a NonbondedSystem (the reference has no integrator)."""
    system.vv_step(dt, nsteps, rebin_every)
