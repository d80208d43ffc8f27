"""workloads.py -- synthetic inputs of the BASELINE configs and the exclusion
bitmask builder.  Pure numpy host code; no physics is evaluated here.

The generator is stateless and counter-based (splitmix64 finaliser in uint64 wrap-around
arithmetic), so any atom can be generated on any rank and the same numbers come out in C, Python
and Julia.
"""
import numpy as np

RHO_STAR = 0.8442
SEED = 87287
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(counter, seed=SEED):
    """u = (splitmix64_finalise(seed + GOLDEN*counter) >> 11) * 2^-53, counter >= 1."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + _GOLDEN * np.asarray(counter, dtype=np.uint64)
    return (mix64(z) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def fcc_box(n, rho=RHO_STAR):
    """Lattice constant a = (4/rho)^(1/3) and box edge L = n*a."""
    a = (4.0 / rho) ** (1.0 / 3.0)
    return a, n * a


def fcc_lattice(n, rho=RHO_STAR, amplitude=0.1, seed=SEED, ids=None):
    """FCC lattice of n^3 conventional cells, N = 4 n^3, atom id = 4*(ix + n*(iy + n*iz)) + b
    with basis {(0,0,0),(.5,.5,0),(.5,0,.5),(0,.5,.5)}*a, each coordinate c of atom `id` displaced by
    amplitude*(2u-1), u = uniform01(3*id + c + 1).  Returns (positions (N,3) float64, L).
    `ids` selects a subset (for per-rank generation)."""
    a, L = fcc_box(n, rho)
    N = 4 * n ** 3
    if ids is None:
        ids = np.arange(N, dtype=np.int64)
    ids = np.asarray(ids, dtype=np.int64)
    b = ids % 4
    cell = ids // 4
    ix = cell % n
    iy = (cell // n) % n
    iz = cell // (n * n)
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    pos = (np.stack([ix, iy, iz], axis=1) + basis[b]) * a
    if amplitude != 0.0:
        ctr = (3 * ids[:, None] + np.arange(3)[None, :] + 1).astype(np.uint64)
        pos = pos + amplitude * (2.0 * uniform01(ctr, seed) - 1.0)
    return np.ascontiguousarray(pos), L


def lj_fluid_atoms(N, eps=1.0, sigma=1.0):
    """fill(LennardJonesAtom(eps, sigma), N) as an (N,2) array (src/lennard_jones.jl:13)."""
    return np.tile(np.array([0.5 * sigma, 2.0 * np.sqrt(eps)]), (N, 1))


def maxwell_velocities(N, temperature=1.44, mass=1.0, seed=SEED):
    """Box-Muller on the same generator (stream offset 2^32), net momentum removed in id order."""
    ids = np.arange(N, dtype=np.uint64)
    base = np.uint64(1 << 32)
    c = (np.uint64(6) * ids[:, None] + np.arange(6, dtype=np.uint64)[None, :] + base + np.uint64(1))
    u = uniform01(c, seed)
    u1 = np.maximum(u[:, 0:3], 1e-300)
    u2 = u[:, 3:6]
    g = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    m = np.broadcast_to(np.asarray(mass, dtype=np.float64), (N,))
    v = g * np.sqrt(temperature / m)[:, None]
    p = (v * m[:, None]).sum(axis=0)
    v = v - p[None, :] / m.sum()
    return np.ascontiguousarray(v)


def exclusion_masks(N, bonds, max_distance=2):
    """Exclusion bitmask from the bond graph: pairs at graph distance 1..max_distance
    (1-2 and 1-3 for the default) are excluded, X = A | (A.A > 0) with the diagonal cleared.
    bonds: (nb,2) 0-based atom ids.  Atoms of a molecule must be contiguous (the reference regroups
    atoms residue-contiguously, src/modelling.jl:330-348) and span fewer than 64 ids.
    Returns (base int32 (N,), mask uint64 (N,)): pair (i,j) excluded iff bit (j-base[i]) of mask[i]."""
    bonds = np.asarray(bonds, dtype=np.int64).reshape(-1, 2)
    nbr = [set() for _ in range(N)]
    for a, b in bonds:
        nbr[a].add(int(b))
        nbr[b].add(int(a))
    base = np.arange(N, dtype=np.int64)
    mask = np.zeros(N, dtype=np.uint64)
    for i in range(N):
        if not nbr[i]:
            continue
        seen = {i}
        frontier = {i}
        for _ in range(max_distance):
            nxt = set()
            for k in frontier:
                nxt |= nbr[k]
            frontier = nxt - seen
            seen |= frontier
        seen.discard(i)
        lo = min(min(seen), i)
        if max(max(seen), i) - lo >= 64:
            raise ValueError("exclusion window of atom %d spans 64 or more ids" % i)
        base[i] = lo
        m = 0
        for j in seen:
            m |= 1 << (j - lo)
        mask[i] = np.uint64(m)
    return base.astype(np.int32), mask


def pairs14(N, bonds):
    """Pairs of atoms exactly three bonds apart (graph distance 3: the "1-4" pairs whose Lennard-Jones interaction a force
    field scales by lj14scale, src/modelling.jl:199), as an (n,2) int32 array with i<j in lexicographic order."""
    bonds = np.asarray(bonds, dtype=np.int64).reshape(-1, 2)
    nbr = [set() for _ in range(N)]
    for a, b in bonds:
        nbr[a].add(int(b))
        nbr[b].add(int(a))
    out = []
    for i in range(N):
        if not nbr[i]:
            continue
        seen = {i}
        frontier = {i}
        for _ in range(3):
            nxt = set()
            for k in frontier:
                nxt |= nbr[k]
            frontier = nxt - seen
            seen |= frontier
        out.extend((i, j) for j in sorted(frontier) if j > i)      # the last frontier: distance exactly 3
    return np.asarray(out, dtype=np.int32).reshape(-1, 2)


def replicate_molecular(pos, box, bonds, per_atom, reps, jitter=0.01, seed=SEED):
    """Replicate a molecular configuration reps^3 times into a cubic box of edge reps*box
    (config 4: the reference's test PDB replicated 9x9x9).  Copy k shifts every atom by the copy
    offset plus a per-copy jitter so that copies are not bit-identical.  per_atom: dict of (N0,...)
    arrays tiled along the atom axis.  Returns (positions, L, bonds, per_atom)."""
    N0 = pos.shape[0]
    ncopy = reps ** 3
    k = np.arange(ncopy)
    off = np.stack([k % reps, (k // reps) % reps, k // (reps * reps)], axis=1).astype(np.float64) * box
    ctr = (3 * k[:, None] + np.arange(3)[None, :] + 1 + (1 << 40)).astype(np.uint64)
    jit = jitter * (2.0 * uniform01(ctr, seed) - 1.0)
    allpos = (pos[None, :, :] + off[:, None, :] + jit[:, None, :]).reshape(ncopy * N0, 3)
    allbonds = (np.asarray(bonds, dtype=np.int64)[None, :, :] + (k * N0)[:, None, None]).reshape(-1, 2)
    out = {name: np.tile(arr, (ncopy,) + (1,) * (arr.ndim - 1)) for name, arr in per_atom.items()}
    return np.ascontiguousarray(allpos), reps * box, allbonds, out


def molecular_system(fixture, reps=9, jitter=0.01, seed=SEED):
    """BASELINE config 4: the reference's molecular test system
    (test/data/dibenzo-p-dioxin-in-water.{pdb,xml}: 1519 atoms, 500 residues, cubic box 24.56 A) replicated
    reps^3 times -- reps = 9 gives 1,107,351 atoms in a box of 221.04 A.  `fixture` is the parsed form of
    those two files (tests/golden/dioxin_water.npz, or modelling.System(...).fixture()).  Lengths in A
    (sigma converted nm -> A), per-type LJ parameters, 1-2/1-3 exclusions from the bond graph as bitmasks;
    cutoff 10 A, switch 9 A.  Returns a dict: positions, L, atoms (N,2), masses, excl = (base, mask), bonds,
    cutoff, switch."""
    pos0 = np.asarray(fixture["positions"], dtype=np.float64)
    box = float(fixture["box"])
    tidx = np.asarray(fixture["type_index"])
    N0 = pos0.shape[0]
    sig = np.asarray(fixture["type_sigma_nm"])[tidx] * 10.0
    eps = np.asarray(fixture["type_epsilon"])[tidx]
    atoms0 = np.stack([0.5 * sig, 2.0 * np.sqrt(eps)], axis=1)
    mass0 = np.asarray(fixture["type_mass"])[tidx]
    base0, mask0 = exclusion_masks(N0, fixture["bonds"])
    pos, L, bonds, per = replicate_molecular(pos0, box, fixture["bonds"], {"atoms": atoms0, "mass": mass0, "mask": mask0},
                                            reps, jitter, seed)
    # the exclusion window of copy k is the window of copy 0 shifted by k*N0 atom ids
    base = (base0.astype(np.int64)[None, :] + (np.arange(reps ** 3, dtype=np.int64) * N0)[:, None]).reshape(-1)
    # 1-4 pairs of copy k are those of copy 0 shifted by k*N0; the scale is the force field's lj14scale
    p14_0 = pairs14(N0, fixture["bonds"]).astype(np.int64)
    p14 = (p14_0[None, :, :] + (np.arange(reps ** 3, dtype=np.int64) * N0)[:, None, None]).reshape(-1, 2).astype(np.int32)
    lj14 = float(fixture["lj14scale"]) if "lj14scale" in fixture else 1.0
    return dict(positions=pos, L=L, atoms=np.ascontiguousarray(per["atoms"]), masses=np.ascontiguousarray(per["mass"]),
                excl=(base.astype(np.int32), np.ascontiguousarray(per["mask"])), bonds=bonds, cutoff=10.0, switch=9.0,
                pairs14=(p14, lj14))
