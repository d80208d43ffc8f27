"""Host-side logic that decides correctness on the GPU but needs no GPU to be checked."""
import ctypes as C

import numpy as np
import pytest


def _threshold(em, hext, rcut):
    h = (C.c_double * 3)(*hext)
    t = C.c_float()
    em._lib.call("emdee_fp16_threshold", h, float(rcut), C.byref(t))
    return float(t.value)


def _r2_fp16(ci, cj):
    """The list kernels' arithmetic on FP16 coordinates (k_force_list_p `test`): d = cj - ci, packed squares, one sum;
    every operation rounds to FP16."""
    f16 = np.float16
    ci, cj = ci.astype(np.float32).astype(f16), cj.astype(np.float32).astype(f16)      # staged through FP32, like the kernels
    d = (cj - ci).astype(f16)
    sq = (d * d).astype(f16)
    s = (d[:, 2] * d[:, 2] + sq[:, 0]).astype(f16)          # hfma2(dzw, dzw, hmul2(dxy, dxy)).low
    return (s + sq[:, 1]).astype(f16)


@pytest.mark.parametrize("hext,rcut", [((9.05, 6.1, 6.1), 2.5), ((9.05, 6.1, 6.1), 2.95), ((15.2, 6.3, 3.4), 3.0),
                                       ((30.0, 30.0, 30.0), 10.0), ((3.9, 3.9, 3.9), 1.2)])
def test_fp16_precull_never_drops_a_pair(em, hext, rcut):
    """For coordinates anywhere in the staged box and every separation with r <= rcut (many of them exactly at rcut),
    the FP16 distance test of the list kernels accepts the pair: the threshold is conservative."""
    thr = np.float16(np.nextafter(np.float32(_threshold(em, hext, rcut)), np.float32(np.inf)))
    thr = thr if float(thr) >= _threshold(em, hext, rcut) else np.nextafter(thr, np.float16(np.inf))   # __float2half_ru
    rng = np.random.default_rng(7)
    n = 400000
    h = np.asarray(hext)
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    r = np.where(rng.random(n) < 0.5, rcut, rcut * rng.random(n) ** (1.0 / 3.0))
    dvec = u * r[:, None]
    ci = (rng.random((n, 3)) * 2 - 1) * h
    cj = ci + dvec
    ok = (np.abs(cj) <= h).all(axis=1)
    ci, cj = ci[ok], cj[ok]
    r2h = _r2_fp16(ci, cj)
    assert (r2h <= thr).all(), float((r2h.astype(np.float64) - float(thr)).max())
    # ... and it is tight enough to be useful: pairs 4 % beyond rcut are rejected more often than not
    cj2 = ci + 1.04 * (cj - ci)
    assert (_r2_fp16(ci, cj2) > thr).mean() > 0.5


def test_cutoff_band_classification():
    """pair_in_range (lj_pair.cuh): t = hi(r2) - (hi(rc2) - 1); t < 0 inside, t > 2 outside, else the exact path.
    A decided pair must stay decided under any perturbation of r2 far larger than the ~1e-13 relative difference
    between the brick-frame r2 and the oracle's rounding sequence."""
    def hi(x):
        return (np.asarray(x, dtype=np.float64).view(np.int64) >> 32).astype(np.int64)

    rng = np.random.default_rng(3)
    for rc in (2.5, 3.0, 10.0, 1.0, 2.0 ** 0.5):
        rc2 = rc * rc
        r2 = rc2 * (1.0 + (rng.random(200000) - 0.5) * 2e-5)
        r2 = np.concatenate([r2, rc2 * (1.0 + (rng.random(200000) - 0.5) * 2.0)])
        t = hi(r2) - (hi(rc2) - 1)
        inside, outside = t < 0, t > 2
        eps = 1e-9
        assert (r2[inside] * (1 + eps) < rc2).all()
        assert (r2[outside] * (1 - eps) > rc2).all()
        band = ~inside & ~outside
        assert (np.abs(r2[band] / rc2 - 1.0) < 4e-6).all()      # the exact path is taken only within ~3 * 2^-20 of rc2


def test_stencil_tables_mirror(em):
    """index2voxel / stencil_vectors / surrounding_cells (src/cells.jl:22-44): the vectorised host mirror against a
    literal loop restatement; 62 + 62 vectors for rc* in [1.74, 2]; action and reaction stencils are
    point reflections of each other, so every cell pair is visited once."""
    from oracle import oracle_np as on

    assert em.index2voxel(1, 5).tolist() == [0, 0, 0] and em.index2voxel(5 * 5 * 2 + 5 * 3 + 4 + 1, 5).tolist() == [4, 3, 2]
    assert em.cells_per_dimension(10.0, 3.0, 2) == 6 and em.cells_per_dimension(167.96, 2.5, 1) == 67
    for rc in (1.0, 1.5, 1.74, 1.9, 2.0, 2.3, 3.0):
        for action in (True, False):
            v = em.stencil_vectors(rc, action)
            assert [tuple(x) for x in v.tolist()] == on.stencil_vectors_loops(rc, action), (rc, action)
        a, r = em.stencil_vectors(rc, True), em.stencil_vectors(rc, False)
        assert sorted(map(tuple, (-a).tolist())) == sorted(map(tuple, r.tolist()))
    assert len(em.stencil_vectors(1.9, True)) == 62 and len(em.stencil_vectors(2.0, False)) == 62
    for L, cutoff, M in ((10.0, 3.0, 6), (10.0, 2.0, 10), (7.0, 2.5, 5)):
        for action in (True, False):
            t = em.surrounding_cells(L, cutoff, M, action)
            assert t.dtype == np.int32 and t.tolist() == on.surrounding_cells_loops(L, cutoff, M, action)
