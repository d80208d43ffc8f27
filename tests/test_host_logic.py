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
