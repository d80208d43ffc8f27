"""CPU tests: the oracle against the golden vectors, its numpy twin and analytic properties.
(The reference's only fixture is lj_sample.xyz; it stores no numeric answers -- see make_golden.py.)"""
import numpy as np
import pytest

from oracle import oracle_np as on


def test_model_and_atom(oracle):
    m = oracle.lj_model(3.0, 2.5)
    assert m.tolist() == [9.0, 6.25, 1.0 / 2.75]              # src/lennard_jones.jl:10
    assert oracle.lj_atom(1, 1).tolist() == [0.5, 2.0]        # src/lennard_jones.jl:13 (eps, sigma) order
    assert oracle.lj_atom(4.0, 3.0).tolist() == [1.5, 4.0]
    assert np.array_equal(m, on.lj_model(3.0, 2.5))


def test_interaction_properties(oracle):
    m = oracle.lj_model(3.0, 2.5)
    a = oracle.lj_atom(1, 1)
    # below the switch radius: plain LJ, W = -r dE/dr = 24 (2 r^-12 - r^-6)
    for r in (0.95, 1.0, 1.5, 2.4):
        E, W = oracle.interaction(r * r, m, a, a)
        assert E == pytest.approx(4 * (r ** -12 - r ** -6), rel=1e-14)
        assert W == pytest.approx(24 * (2 * r ** -12 - r ** -6), rel=1e-13)
    # W = -r dE/dr inside the switching region, by central differences
    for r in (2.55, 2.7, 2.9, 2.99):
        h = 1e-6
        Ep = oracle.interaction((r + h) ** 2, m, a, a)[0]
        Em = oracle.interaction((r - h) ** 2, m, a, a)[0]
        W = oracle.interaction(r * r, m, a, a)[1]
        assert W == pytest.approx(-r * (Ep - Em) / (2 * h), rel=1e-6)
    # g -> 0 approaching the cutoff from below; full unswitched LJ beyond it
    assert abs(oracle.interaction(9.0 * (1 - 1e-12), m, a, a)[0]) < 1e-15
    r = 3.2
    assert oracle.interaction(r * r, m, a, a)[0] == pytest.approx(4 * (r ** -12 - r ** -6), rel=1e-14)
    # numpy twin, scalar by scalar, bit for bit
    for r2 in (0.9, 1.3, 6.25, 6.3, 7.7, 8.99, 9.0, 9.5, 20.0):
        Ec, Wc = oracle.interaction(r2, m, a, a)
        En, Wn = on.interaction(np.array([r2]), m, a[0], a[1], a[0], a[1])
        assert Ec == En[0] and Wc == Wn[0]


def test_tiles(oracle, lj_sample):
    t = oracle.tiles(800)
    assert t.shape == (325, 2)
    assert np.array_equal(t, lj_sample["tiles"])
    assert t[0].tolist() == [1, 1] and t[24].tolist() == [25, 25] and t[25].tolist() == [1, 2] and t[-1].tolist() == [1, 25]
    assert oracle.tiles(1).tolist() == [[1, 1]] and oracle.tiles(33).tolist() == [[1, 1], [2, 2], [1, 2]]


def test_allpairs_golden(oracle, lj_sample):
    g = lj_sample
    pos = g["positions"]
    model = oracle.lj_model(float(g["cutoff"]), float(g["switch"]))
    atoms = np.tile(oracle.lj_atom(1, 1), (800, 1))
    f, e, w = oracle.naive_allpairs(pos, float(g["L"]), model, atoms)
    assert np.array_equal(f, g["allpairs_forces"]) and np.array_equal(e, g["allpairs_energies"])
    assert np.array_equal(w, g["allpairs_virials"])
    # SURVEY Appendix B numbers (independent throw-away restatement made during the survey)
    assert e.sum() == pytest.approx(-4466.170939617827, rel=1e-13)
    assert w.sum() == pytest.approx(-2000.679519390219, rel=1e-13)
    assert np.sqrt((f ** 2).sum(1).mean()) == pytest.approx(26.24427184765313, rel=1e-13)
    # tile order == naive order up to summation order
    ft, et, wt = oracle.tiles_allpairs(pos, float(g["L"]), g["tiles"], model, atoms)
    assert np.abs(ft - f).max() < 1e-11 and np.abs(et - e).max() < 1e-12 and np.abs(wt - w).max() < 1e-11


def test_allpairs_numpy_twin_bitwise(oracle, lj_sample):
    g = lj_sample
    n = 160   # the twin's sequential accumulation is slow; a prefix of the fixture is enough
    pos = g["positions"][:n]
    model = oracle.lj_model(3.0, 2.5)
    atoms = np.tile(oracle.lj_atom(1, 1), (n, 1))
    f, e, w = oracle.naive_allpairs(pos, 10.0, model, atoms)
    f2, e2, w2 = on.naive_allpairs_f64(pos, 10.0, model, atoms)
    assert np.array_equal(f, f2) and np.array_equal(e, e2) and np.array_equal(w, w2)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_allpairs_scalar_loop_bitwise(oracle, lj_sample, dtype):
    """The C oracle (both instantiations) against a literal scalar-loop restatement of naively_compute_nonbonded!
    with the reference's mixed-precision accumulators: bit for bit, N = 60 atoms of the fixture."""
    n = 60
    pos = np.ascontiguousarray(lj_sample["positions"][:n]).astype(dtype)
    model = oracle.lj_model(3.0, 2.5, dtype)
    atoms = np.tile(oracle.lj_atom(1, 1, dtype), (n, 1))
    f, e, w = oracle.naive_allpairs(pos, 10.0, model, atoms)
    f2, e2, w2 = on.naive_allpairs_loops(pos, 10.0, model, atoms)
    assert f.dtype == dtype and f2.dtype == dtype
    assert np.array_equal(f, f2) and np.array_equal(e, e2) and np.array_equal(w, w2)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [64, 50])
def test_tile_schedule_scalar_loop_bitwise(oracle, lj_sample, dtype, n):
    """compute_tile!'s lane schedule (partner lane, returning lane, epilogue order) restated a second time in plain
    loops: the C oracle's tile simulation matches it bit for bit (n = 64: full tiles; n = 50: masked tail lanes)."""
    pos = np.ascontiguousarray(lj_sample["positions"][:n]).astype(dtype)
    model = oracle.lj_model(3.0, 2.5, dtype)
    atoms = np.tile(oracle.lj_atom(1, 1, dtype), (n, 1))
    tiles = oracle.tiles(n)
    f, e, w = oracle.tiles_allpairs(pos, 10.0, tiles, model, atoms)
    f2, e2, w2 = on.tiles_allpairs_loops(pos, 10.0, tiles, model, atoms)
    assert np.array_equal(f, f2) and np.array_equal(e, e2) and np.array_equal(w, w2)
    # and the schedule visits every ordered (lane, partner) exactly once: tile order == naive order up to summation order
    fn, en, wn = oracle.naive_allpairs(pos, 10.0, model, atoms)
    tol = 2e-4 if dtype == np.float32 else 1e-11
    assert np.abs(f - fn).max() < tol * max(1.0, np.abs(fn).max()) and abs(e.sum() - en.sum()) < tol * abs(en.sum())


def test_float32_reference_criterion(oracle, lj_sample):
    """The reference's own test (test/runtests.jl:39-41): tile kernel vs naive loop < 1e-4, Float32."""
    g = lj_sample
    x32 = g["positions"].astype(np.float32)
    m32 = oracle.lj_model(3.0, 2.5, np.float32)
    a32 = np.tile(oracle.lj_atom(1, 1, np.float32), (800, 1))
    fn, en, wn = oracle.naive_allpairs(x32, 10.0, m32, a32)
    ft, et, wt = oracle.tiles_allpairs(x32, 10.0, g["tiles"], m32, a32)
    assert (ft - fn).max() < 1e-4 and (et - en).max() < 1e-4 and (wt - wn).max() < 1e-4
    assert np.abs(ft - fn).max() < 1e-4
    # Float32 differs from Float64 at the 1e-5 F_rms level: parity is vs FP64
    frms = np.sqrt((g["allpairs_forces"] ** 2).sum(1).mean())
    assert 1e-7 < np.abs(fn - g["allpairs_forces"]).max() / frms < 1e-4


def test_cutoff_golden(oracle, lj_sample):
    g = lj_sample
    atoms = np.tile(oracle.lj_atom(1, 1), (800, 1))
    for ndiv in (1, 2):
        r = oracle.cutoff_cells(g["positions"], 10.0, 3.0, 2.5, atoms, ndiv=ndiv)
        assert r["npairs"] == 35677 and np.array_equal(r["digest"], g["cutoff_digest"])
        assert np.abs(r["forces"] - g["cutoff_forces"]).max() < 1e-11
        assert np.abs(r["energies"] - g["cutoff_energies"]).max() < 1e-12
    assert r["E"] == pytest.approx(-4292.184041039327, rel=1e-13)
    assert r["W"] == pytest.approx(-957.3126906742807, rel=1e-13)
    pairs, dig = oracle.pair_set_brute(g["positions"], 10.0, 9.0)
    assert np.array_equal(pairs, g["cutoff_pairs"]) and np.array_equal(dig, g["cutoff_digest"])
    assert np.array_equal(oracle.pair_set_cells(g["positions"], 10.0, 3.0, 1), pairs)
    assert np.array_equal(oracle.pair_set_cells(g["positions"], 10.0, 3.0, 2), pairs)
    assert np.array_equal(on.pair_digest(pairs), dig)


def test_cutoff_numpy_twin(oracle, lj_sample):
    g = lj_sample
    n = 400
    pos = g["positions"][:n]
    atoms = np.tile(oracle.lj_atom(1, 1), (n, 1))
    f, e, w, ij = on.cutoff_compute(pos, 10.0, 3.0, 2.5, atoms)
    r = oracle.cutoff_cells(pos, 10.0, 3.0, 2.5, atoms, ndiv=1)
    pairs, dig = oracle.pair_set_brute(pos, 10.0, 9.0)
    assert np.array_equal(ij, pairs) and np.array_equal(on.pair_digest(ij), r["digest"])
    assert np.abs(f - r["forces"]).max() < 1e-12 and np.abs(e - r["energies"]).max() < 1e-13
    assert np.abs(w - r["virials"]).max() < 1e-12


def test_cell_index_golden(oracle, lj_sample):
    g = lj_sample
    for ndiv, key, first in ((1, "cell_index_ndiv1", [21, 9, 12, 2, 15]), (2, "cell_index_ndiv2", [186, 35, 77, 45, 89])):
        M = oracle.cells_per_dimension(10.0, 3.0, ndiv)
        assert M == 3 * ndiv
        idx = oracle.cell_index(g["positions"], 10.0, M)
        assert np.array_equal(idx, g[key]) and idx[:5].tolist() == first
        assert np.array_equal(idx, on.cell_index(g["positions"], 10.0, M))
        assert np.array_equal(idx, oracle.cell_index(g["positions"].astype(np.float32), 10.0, M))
    # Q7: a tiny negative coordinate rounds s - floor(s) to 1.0; the index must stay in range
    pos = np.array([[-1e-18, 0.0, 9.999999999]])
    assert oracle.cell_index(pos, 10.0, 6).tolist() == [1 + 5 + (0 + 5 * 6) * 6]
    assert oracle.cells_per_dimension(16.79596, 2.5, 1) == 6 and oracle.cells_per_dimension(16.79596, 2.5, 2) == 13


def test_fcc_zero_force_and_lattice_energy(oracle, em):
    pos, L = em.workloads.fcc_lattice(6, amplitude=0.0)
    N = pos.shape[0]
    atoms = em.workloads.lj_fluid_atoms(N)
    r = oracle.cutoff_cells(pos, L, 2.5, 2.0, atoms, ndiv=1)
    assert np.abs(r["forces"]).max() < 1e-11
    assert np.ptp(r["energies"]) < 1e-12
    # every atom of a perfect FCC lattice sees the same shells: 12, 6, 24, 12, 24 neighbours within 2.5 sigma
    a = (4 / 0.8442) ** (1 / 3)
    shells = [(12, a / np.sqrt(2)), (6, a), (24, a * np.sqrt(1.5)), (12, a * np.sqrt(2)), (24, a * np.sqrt(2.5))]
    m = oracle.lj_model(2.5, 2.0)
    e = 0.5 * sum(n * oracle.interaction(d * d, m, atoms[0], atoms[0])[0] for n, d in shells if d <= 2.5)
    assert r["energies"][0] == pytest.approx(e, rel=1e-12)
    assert r["npairs"] == N * sum(n for n, d in shells if d <= 2.5) // 2


def test_exclusions(oracle, em, dioxin_water):
    g = dioxin_water
    N = g["positions"].shape[0]
    base, mask = em.workloads.exclusion_masks(N, g["bonds"])
    # water: every atom excludes the other two atoms of its molecule (1-2 and 1-3)
    ow = 22 + 1          # first water is atoms 22,23,24 (H, O, H)
    for i in (22, 23, 24):
        ex = {int(base[i]) + k for k in range(64) if (int(mask[i]) >> k) & 1}
        assert ex == {22, 23, 24} - {i}
    # symmetric
    for i in range(0, 60):
        for k in range(64):
            if (int(mask[i]) >> k) & 1:
                j = int(base[i]) + k
                assert (int(mask[j]) >> (i - int(base[j]))) & 1
    # the oracle drops exactly the excluded pairs that are inside the cutoff
    pos = g["positions"]
    L = float(g["box"])
    full, _ = oracle.pair_set_brute(pos, L, 100.0)
    kept, _ = oracle.pair_set_brute(pos, L, 100.0, excl=(base, mask))
    dropped = set(map(tuple, full.tolist())) - set(map(tuple, kept.tolist()))
    assert dropped and all((int(mask[i]) >> (j - int(base[i]))) & 1 for i, j in dropped)
    nexcl = sum(bin(int(m)).count("1") for m in mask) // 2
    assert len(dropped) == nexcl        # bonded neighbours are all well inside 10 A


def _molecular_inputs(em, g):
    sig = g["type_sigma_nm"][g["type_index"]] * 10.0
    eps = g["type_epsilon"][g["type_index"]]
    atoms = np.stack([0.5 * sig, 2.0 * np.sqrt(eps)], axis=1)
    base, mask = em.workloads.exclusion_masks(g["positions"].shape[0], g["bonds"])
    return atoms, base, mask


def test_molecular_golden(oracle, em, dioxin_water):
    """Config 4's single box (1519 atoms, 5 LJ classes, 1-2/1-3 exclusions, rc = 10 A): stored pair digest, pair count,
    E, W and forces; the exclusion masks themselves are part of the fixture."""
    g = dioxin_water
    atoms, base, mask = _molecular_inputs(em, g)
    assert np.array_equal(base, g["excl_base"]) and np.array_equal(mask, g["excl_mask"])
    r = oracle.cutoff_cells(g["positions"], float(g["box"]), 10.0, 9.0, atoms, ndiv=1, excl=(base, mask))
    assert r["npairs"] == int(g["cutoff_npairs"]) == 323944 and np.array_equal(r["digest"], g["cutoff_digest"])
    assert r["E"] == pytest.approx(float(g["cutoff_E"]), rel=1e-13) and r["W"] == pytest.approx(float(g["cutoff_W"]), rel=1e-13)
    assert np.abs(r["forces"] - g["cutoff_forces"]).max() <= 1e-12 * np.abs(g["cutoff_forces"]).max()
    fast = oracle.cutoff_cells(g["positions"], float(g["box"]), 10.0, 9.0, atoms, ndiv=1, excl=(base, mask), fast=True)
    assert np.array_equal(fast["digest"], r["digest"]) and abs(fast["E"] - r["E"]) <= 1e-11 * abs(r["E"])


def test_molecular_numpy_twin(oracle, em, dioxin_water):
    """The same evaluation by the independent numpy restatement: identical pair set, forces and energies to rounding."""
    g = dioxin_water
    atoms, base, mask = _molecular_inputs(em, g)
    f, e, w, ij = on.cutoff_compute(g["positions"], float(g["box"]), 10.0, 9.0, atoms, base, mask)
    assert ij.shape[0] == int(g["cutoff_npairs"]) and np.array_equal(on.pair_digest(ij), g["cutoff_digest"])
    assert np.abs(f - g["cutoff_forces"]).max() < 1e-11 and abs(e.sum() - float(g["cutoff_E"])) <= 1e-12 * abs(float(g["cutoff_E"]))


def test_vv_energy_conservation(oracle, em):
    pos, L = em.workloads.fcc_lattice(5)
    N = pos.shape[0]
    atoms = em.workloads.lj_fluid_atoms(N)
    vel = em.workloads.maxwell_velocities(N, 1.0)
    mass = np.ones(N)
    r0 = oracle.cutoff_cells(pos, L, 2.5, 2.0, atoms, ndiv=1)
    p, v, f = oracle.vv_steps(pos, vel, r0["forces"], mass, L, 2.5, 2.0, atoms, 0.002, 50)
    r1 = oracle.cutoff_cells(p, L, 2.5, 2.0, atoms, ndiv=1)
    H0 = r0["E"] + 0.5 * (vel ** 2).sum()
    H1 = r1["E"] + 0.5 * (v ** 2).sum()
    assert abs(H1 - H0) / N < 1e-4                     # O(dt^2) fluctuation of the VV shadow Hamiltonian
    p2, v2, _ = oracle.vv_steps(pos, vel, r0["forces"], mass, L, 2.5, 2.0, atoms, 0.001, 100)
    H2 = oracle.cutoff_cells(p2, L, 2.5, 2.0, atoms, ndiv=1)["E"] + 0.5 * (v2 ** 2).sum()
    assert abs(H2 - H0) < 0.5 * abs(H1 - H0)           # halving dt cuts the energy error ~4x
    assert np.abs(f - r1["forces"]).max() < 1e-10
    assert np.abs(v.sum(axis=0)).max() < 1e-9          # momentum conserved


def test_vv_numpy_twin(oracle, em):
    """The oracle's velocity-Verlet against the numpy twin (independent force evaluation, unfused updates), unequal
    masses, five steps: positions and velocities to rounding."""
    pos, L = em.workloads.fcc_lattice(4)
    N = pos.shape[0]
    atoms = em.workloads.lj_fluid_atoms(N)
    vel = em.workloads.maxwell_velocities(N, 1.2)
    mass = 1.0 + 0.5 * em.workloads.uniform01(np.arange(1, N + 1) + (1 << 50))
    f0 = oracle.cutoff_cells(pos, L, 2.5, 2.0, atoms, ndiv=1)["forces"]
    p, v, f = oracle.vv_steps(pos, vel, f0, mass, L, 2.5, 2.0, atoms, 0.004, 5)
    p2, v2, f2 = on.vv_steps_twin(pos, vel, f0, mass, L, 2.5, 2.0, atoms, 0.004, 5)
    assert np.abs(p - p2).max() < 1e-13 and np.abs(v - v2).max() < 1e-12 and np.abs(f - f2).max() < 1e-10


def test_workload_generator(em):
    pos, L = em.workloads.fcc_lattice(10)
    assert pos.shape == (4000, 3) and L == pytest.approx(16.79596, abs=1e-5)
    sub, _ = em.workloads.fcc_lattice(10, ids=np.array([0, 5, 3999]))
    assert np.array_equal(sub, pos[[0, 5, 3999]])        # stateless: any atom on any rank
    p0, _ = em.workloads.fcc_lattice(10, amplitude=0.0)
    assert np.abs(pos - p0).max() <= 0.1 and np.abs(pos - p0).max() > 0.09
    # splitmix64 known answer: first output of the generator seeded with 0 (public test vector)
    z = np.uint64(0x9E3779B97F4A7C15)
    assert int(em.workloads.mix64(z)) == 0xE220A8397B1DCDAF
    v = em.workloads.maxwell_velocities(4000, 1.44)
    assert np.abs(v.sum(axis=0)).max() < 1e-9 and (v ** 2).mean() == pytest.approx(1.44, rel=0.05)
