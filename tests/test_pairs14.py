"""1-4 scaling (lj14scale) in the oracle: the pair list from the bond graph, the C correction against the numpy twin and the
golden vectors, and the corrected force as the gradient of the corrected energy (CPU only)."""
import numpy as np
import pytest

from oracle import oracle_np as on


def _fixture_system(g):
    import emdee_jl_b200 as em

    pos, L = g["positions"], float(g["box"])
    tidx = g["type_index"]
    atoms = np.stack([0.5 * g["type_sigma_nm"][tidx] * 10.0, 2.0 * np.sqrt(g["type_epsilon"][tidx])], axis=1)
    base, mask = em.workloads.exclusion_masks(pos.shape[0], g["bonds"])
    return em, pos, L, atoms, (base, mask)


def test_pairs14_are_the_pairs_three_bonds_apart(dioxin_water):
    g = dioxin_water
    em, pos, L, atoms, (base, mask) = _fixture_system(g)
    N = pos.shape[0]
    p14 = em.workloads.pairs14(N, g["bonds"])
    assert np.array_equal(p14, g["pairs14"]) and p14.shape == (47, 2)         # dibenzo-p-dioxin: 22 atoms, 47 such pairs; water has none
    assert np.all(p14[:, 0] < p14[:, 1]) and p14.max() < 22
    # not excluded (distance 1, 2), but inside the distance-3 neighbourhood
    b3, m3 = em.workloads.exclusion_masks(N, g["bonds"], max_distance=3)
    for i, j in p14:
        assert not (int(mask[i]) >> (int(j) - int(base[i]))) & 1
        assert (int(m3[i]) >> (int(j) - int(b3[i]))) & 1
    # every distance-3 neighbour is listed exactly once
    n3 = sum(bin(int(m3[i]) & ~(int(mask[i]) << (int(base[i]) - int(b3[i])) if base[i] >= b3[i] else int(mask[i]) >> (int(b3[i]) - int(base[i])))).count("1")
             for i in range(22))
    assert n3 == 2 * p14.shape[0]


def test_pairs14_correction_c_vs_numpy_vs_golden(oracle, dioxin_water):
    g = dioxin_water
    em, pos, L, atoms, excl = _fixture_system(g)
    scale = float(g["lj14scale"])
    assert scale == 0.5                                                      # test/data/dibenzo-p-dioxin-in-water.xml:84
    plain = oracle.cutoff_cells(pos, L, 10.0, 9.0, atoms, ndiv=1, excl=excl)
    with14 = oracle.cutoff_cells(pos, L, 10.0, 9.0, atoms, ndiv=1, excl=excl, pairs14=(g["pairs14"], scale))
    assert np.array_equal(with14["digest"], plain["digest"]) and with14["n14_inside"] == 47      # the pair SET is not changed
    df, de, dw, n = on.pairs14_correction(pos, L, 10.0, 9.0, atoms, g["pairs14"], scale)
    assert n == 47
    assert np.abs(plain["forces"] + df - with14["forces"]).max() < 1e-11
    assert np.abs(plain["energies"] + de - with14["energies"]).max() < 1e-12 and np.abs(plain["virials"] + dw - with14["virials"]).max() < 1e-11
    assert with14["E"] == pytest.approx(float(g["cutoff14_E"]), rel=1e-14) and with14["W"] == pytest.approx(float(g["cutoff14_W"]), rel=1e-14)
    assert np.array_equal(with14["forces"], g["cutoff14_forces"])
    assert abs(with14["E"] - plain["E"]) > 1e-3                              # the correction is not a no-op on this fixture
    # scale = 1 is the identity, scale = 0 removes the pairs' energy entirely
    same = oracle.cutoff_cells(pos, L, 10.0, 9.0, atoms, ndiv=1, excl=excl, pairs14=(g["pairs14"], 1.0))
    assert same["E"] == plain["E"] and np.array_equal(same["forces"], plain["forces"])
    gone = oracle.cutoff_cells(pos, L, 10.0, 9.0, atoms, ndiv=1, excl=excl, pairs14=(g["pairs14"], 0.0))
    assert gone["E"] - plain["E"] == pytest.approx(2.0 * (with14["E"] - plain["E"]), rel=1e-9)


def test_pairs14_force_is_the_gradient(oracle, dioxin_water):
    g = dioxin_water
    em, pos, L, atoms, excl = _fixture_system(g)
    p14 = (g["pairs14"], 0.5)
    f = oracle.cutoff_cells(pos, L, 10.0, 9.0, atoms, ndiv=1, excl=excl, pairs14=p14)["forces"]
    h = 1e-5
    for atom, c in ((0, 0), (3, 1), (16, 2)):                                # atoms of the solute that take part in 1-4 pairs
        p = pos.copy(); p[atom, c] += h
        ep = oracle.cutoff_cells(p, L, 10.0, 9.0, atoms, ndiv=1, excl=excl, pairs14=p14)["E"]
        p[atom, c] -= 2 * h
        em_ = oracle.cutoff_cells(p, L, 10.0, 9.0, atoms, ndiv=1, excl=excl, pairs14=p14)["E"]
        assert f[atom, c] == pytest.approx(-(ep - em_) / (2 * h), rel=2e-6, abs=1e-6)
