"""Size-independent properties of the CPU oracle (the checker the GPU parity tests lean on): with no reference binary
to pin it (Julia is absent), every symmetry the domain offers is checked here on seeded random fluids.  CPU only."""
import numpy as np
import pytest


def _fluid(n_side, L, seed, jitter=0.25):
    """Jittered simple-cubic fluid: no overlaps, no pair near the cutoff by accident of symmetry."""
    rng = np.random.default_rng(seed)
    g = (np.arange(n_side) + 0.5) * (L / n_side)
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    return pos + rng.uniform(-jitter, jitter, pos.shape)


def _mixed_atoms(N, seed):
    rng = np.random.default_rng(seed)
    sig = rng.choice([0.9, 1.0, 1.1], N)
    eps = rng.choice([0.5, 1.0, 1.5], N)
    return np.stack([0.5 * sig, 2.0 * np.sqrt(eps)], axis=1)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_translation_permutation_newton(oracle, seed):
    L, rc, rs = 9.0, 2.5, 2.0
    pos = _fluid(8, L, seed)
    N = pos.shape[0]
    atoms = _mixed_atoms(N, seed)
    a = oracle.cutoff_cells(pos, L, rc, rs, atoms, ndiv=1)
    frms = np.sqrt((a["forces"] ** 2).sum(axis=1).mean())
    # Newton's third law and per-atom halves: sum of forces vanishes, sum of per-atom energies is the total
    assert np.abs(a["forces"].sum(axis=0)).max() < 1e-11 * frms * N
    assert abs(a["energies"].sum() - a["E"]) <= 1e-12 * abs(a["E"]) and abs(a["virials"].sum() - a["W"]) <= 1e-12 * abs(a["W"])
    # a rigid translation (with periodic re-entry) changes nothing physical
    shift = np.array([0.37, -5.21, 13.9])
    b = oracle.cutoff_cells(pos + shift, L, rc, rs, atoms, ndiv=1)
    assert b["npairs"] == a["npairs"]
    assert abs(b["E"] - a["E"]) <= 1e-11 * abs(a["E"]) and abs(b["W"] - a["W"]) <= 1e-11 * abs(a["W"])
    assert np.abs(b["forces"] - a["forces"]).max() <= 1e-10 * frms
    # relabelling the atoms permutes the per-atom outputs and leaves the pair count alone
    perm = np.random.default_rng(seed + 10).permutation(N)
    c = oracle.cutoff_cells(pos[perm], L, rc, rs, atoms[perm], ndiv=1)
    assert c["npairs"] == a["npairs"]
    assert np.abs(c["forces"] - a["forces"][perm]).max() <= 1e-10 * frms
    assert np.abs(c["energies"] - a["energies"][perm]).max() <= 1e-10 * np.abs(a["energies"]).max()
    # the two cell geometries and the brute-force pair set agree bit for bit on the pair set
    d = oracle.cutoff_cells(pos, L, rc, rs, atoms, ndiv=2)
    _, dig = oracle.pair_set_brute(pos, L, rc * rc)
    assert np.array_equal(d["digest"], a["digest"]) and np.array_equal(dig, a["digest"])


def test_length_scaling(oracle):
    """LJ is scale-free: r, L, rc, rs, sigma -> lambda * (...) leaves E and W unchanged and divides forces by lambda
    (lambda a power of two: the floating-point sequence is the same up to exponents, so this holds to rounding)."""
    L, rc, rs, lam = 9.0, 2.5, 2.0, 4.0
    pos = _fluid(8, L, 5)
    N = pos.shape[0]
    atoms = _mixed_atoms(N, 5)
    a = oracle.cutoff_cells(pos, L, rc, rs, atoms, ndiv=1)
    scaled = atoms.copy()
    scaled[:, 0] *= lam
    b = oracle.cutoff_cells(pos * lam, L * lam, rc * lam, rs * lam, scaled, ndiv=1)
    assert np.array_equal(a["digest"], b["digest"])
    assert abs(a["E"] - b["E"]) <= 1e-13 * abs(a["E"]) and abs(a["W"] - b["W"]) <= 1e-13 * abs(a["W"])
    assert np.abs(a["forces"] - lam * b["forces"]).max() <= 1e-13 * np.abs(a["forces"]).max()


def test_exclusions_remove_exactly_their_pairs(oracle, em):
    """E(all pairs) - E(with exclusions) = sum of interaction() over the excluded pairs inside the cutoff."""
    L, rc, rs = 9.0, 2.5, 2.0
    pos = _fluid(8, L, 7)
    N = pos.shape[0]
    atoms = _mixed_atoms(N, 7)
    rng = np.random.default_rng(7)
    # "molecules" of 4 consecutive ids bonded in a chain: 1-2 and 1-3 excluded, 1-4 kept
    bonds = np.array([(4 * m + k, 4 * m + k + 1) for m in range(N // 4) for k in range(3)])
    base, mask = em.workloads.exclusion_masks(N, bonds)
    full = oracle.cutoff_cells(pos, L, rc, rs, atoms, ndiv=1)
    ex = oracle.cutoff_cells(pos, L, rc, rs, atoms, ndiv=1, excl=(base, mask))
    model = oracle.lj_model(rc, rs)
    dE, removed = 0.0, 0
    for m in range(N // 4):
        for i, j in ((0, 1), (1, 2), (2, 3), (0, 2), (1, 3)):
            a, b = 4 * m + i, 4 * m + j
            d = pos[a] / L - pos[b] / L
            d -= np.rint(d)
            v = L * d
            r2 = float(v @ v)
            if r2 <= rc * rc:
                dE += oracle.interaction(r2, model, atoms[a], atoms[b])[0]
                removed += 1
    assert full["npairs"] - ex["npairs"] == removed and removed > 0
    assert abs((full["E"] - ex["E"]) - dE) <= 1e-10 * max(abs(dE), 1.0)
    # the brute-force pair set honours the same masks
    ij, dig = oracle.pair_set_brute(pos, L, rc * rc, excl=(base, mask))
    assert ij.shape[0] == ex["npairs"] and np.array_equal(dig, ex["digest"])
    same = (ij[:, 0] // 4 == ij[:, 1] // 4)
    assert set(map(tuple, (ij[same] % 4).tolist())) <= {(0, 3)}


def test_exclusion_masks_against_graph_distances(em):
    """workloads.exclusion_masks vs all-pairs shortest paths on random small molecules (bit j-base[i] of mask[i])."""
    rng = np.random.default_rng(11)
    N, bonds, first = 0, [], []
    for _ in range(40):
        n = int(rng.integers(1, 9))
        first.append((N, n))
        for k in range(1, n):                       # random tree + an occasional ring closure
            bonds.append((N + k, N + int(rng.integers(0, k))))
        if n > 3 and rng.random() < 0.5:
            bonds.append((N, N + n - 1))
        N += n
    bonds = np.array(sorted(set((min(a, b), max(a, b)) for a, b in bonds)))
    for maxd in (1, 2, 3):
        base, mask = em.workloads.exclusion_masks(N, bonds, max_distance=maxd)
        dist = np.full((N, N), 99)
        np.fill_diagonal(dist, 0)
        for a, b in bonds:
            dist[a, b] = dist[b, a] = 1
        for k in range(N):                          # Floyd-Warshall (N ~ 180)
            dist = np.minimum(dist, dist[:, k:k + 1] + dist[k:k + 1, :])
        for i in range(N):
            want = {j for j in range(N) if 1 <= dist[i, j] <= maxd}
            got = {int(base[i]) + b for b in range(64) if (int(mask[i]) >> b) & 1}
            assert got == want, (i, maxd)


def test_float32_instantiation_tracks_float64(oracle):
    """The Float32 instantiation (the reference's own type) agrees with FP64 to single-precision accuracy."""
    L, rc, rs = 9.0, 3.0, 2.5
    pos = _fluid(6, L, 9)
    N = pos.shape[0]
    atoms = np.tile(oracle.lj_atom(1, 1), (N, 1))
    f64 = oracle.naive_allpairs(pos, L, oracle.lj_model(rc, rs), atoms)
    f32 = oracle.naive_allpairs(pos.astype(np.float32), L, oracle.lj_model(rc, rs, np.float32), atoms.astype(np.float32))
    assert f32[0].dtype == np.float32
    scale = np.abs(f64[0]).max()
    assert np.abs(f32[0] - f64[0]).max() <= 2e-5 * scale
    assert abs(f32[1].sum() - f64[1].sum()) <= 2e-5 * abs(f64[1].sum())
