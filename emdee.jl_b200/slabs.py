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
