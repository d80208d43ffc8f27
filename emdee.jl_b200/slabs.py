"""slabs.py -- host-side logic of the 1-D slab decomposition (mirrors do_bin_slab in csrc/emdee_b200.cu).

The cell grid's slowest index (z, src/cells.jl:85: index = 1 + v1 + (v2 + v3*M)*M) is split into
contiguous plane ranges, one per rank (one process per GPU).  Every rank needs the R boundary planes
of its two ring neighbours as ghosts.  Pure integer logic, shared by the tests (gloo, CPU) and by
callers that want to know who owns what."""


def plane_range(rank, nranks, M):
    """Global z planes [z0, z1) owned by `rank` (same integer formula as the library)."""
    return rank * M // nranks, (rank + 1) * M // nranks


def owner_of_plane(z, nranks, M):
    z %= M
    for r in range(nranks):
        z0, z1 = plane_range(r, nranks, M)
        if z0 <= z < z1:
            return r
    raise ValueError("plane %d has no owner" % z)


def neighbours(rank, nranks):
    """(lower, upper) ranks on the periodic ring."""
    return (rank - 1) % nranks, (rank + 1) % nranks


def check(nranks, M, R):
    """The decomposition is valid iff the grid supports a cell neighbourhood (M >= 2R+1) and every
    slab is at least R planes thick (ghosts then come from the immediate neighbours only)."""
    return M >= 2 * R + 1 and M // nranks >= max(R, 1)


def halo_plan(rank, nranks, M, R):
    """Global plane lists: what this rank sends down / up and what it receives as lower / upper ghosts.
    Message order inside one group (also with 2 ranks, where both neighbours are the same peer):
    sends [to lower, to upper], receives [from upper, from lower]."""
    z0, z1 = plane_range(rank, nranks, M)
    return {
        "send_to_lower": [z0 + k for k in range(R)],               # my bottom planes -> lower's upper ghosts
        "send_to_upper": [z1 - R + k for k in range(R)],           # my top planes    -> upper's lower ghosts
        "recv_lower_ghosts": [(z0 - R + k) % M for k in range(R)],
        "recv_upper_ghosts": [(z1 + k) % M for k in range(R)],
        "order": ("send_to_lower", "send_to_upper", "recv_upper_ghosts", "recv_lower_ghosts"),
    }


def migration_targets(zplane, rank, nranks, M):
    """Where the atoms a rank owns go at a re-binning, from the global z plane of their new cell -- the rule of
    k_cell_index_slab (csrc/slab.cuh): with dz = (z - z0) mod M, an atom stays if dz < nz, moves up if it is nearer
    to the top of the slab (dz - nz < M - dz) and down otherwise.  0 = stays, +1 = to the upper neighbour, -1 = to
    the lower one.  An atom that lands beyond the neighbour's slab is an error (between re-binnings an atom moves at
    most skin/2, less than one plane)."""
    z0, z1 = plane_range(rank, nranks, M)
    nz = z1 - z0
    out = []
    for z in zplane:
        dz = (int(z) - z0) % M
        if dz < nz:
            out.append(0)
            continue
        step = 1 if dz - nz < M - dz else -1
        t0, t1 = plane_range((rank + step) % nranks, nranks, M)
        if not t0 <= int(z) % M < t1:
            raise ValueError("an atom moved to plane %d, beyond the neighbours of rank %d" % (int(z) % M, rank))
        out.append(step)
    return out
