"""oracle_c.py -- ctypes binding of oracle/liboracle.so (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Arrays are numpy; positions are (N,3) C-contiguous (== the reference's 3xN
column-major layout, src/nonbonded.jl:60), atoms are (N,2) {half_sigma, twice_sqrt_eps}.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

_d = C.c_double
_f = C.c_float
_i32 = C.c_int32
_i64 = C.c_int64
_u64 = C.c_uint64
_p = C.c_void_p


def build(force=False):
    """Compile liboracle.so / liboracle_fast.so with oracle/Makefile (building the checker is not using it)."""
    if force:
        subprocess.check_call(["make", "-s", "-C", _HERE, "clean"])
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib(fast=False):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name not in _LIBS:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.restype = None
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_mix64.restype = _u64
        L.oracle_mix64.argtypes = [_u64]
        L.oracle_tiles.restype = _i64
        L.oracle_tiles.argtypes = [_i64, _p]
        L.oracle_cells_per_dimension.restype = _i32
        L.oracle_cells_per_dimension.argtypes = [_d, _d, C.c_int]
        for suf, r in (("f64", _d), ("f32", _f)):
            getattr(L, "oracle_lj_model_" + suf).argtypes = [_d, _d, _p]
            getattr(L, "oracle_lj_atom_" + suf).argtypes = [_d, _d, _p]
            getattr(L, "oracle_interaction_" + suf).argtypes = [r, _p, _p, _p, _p]
            getattr(L, "oracle_naive_allpairs_" + suf).argtypes = [_i64, _p, r, _p, _p, _p, _p, _p]
            getattr(L, "oracle_tiles_allpairs_" + suf).argtypes = [_i64, _p, r, _p, _i64, _p, _p, C.c_int, _p, _p, _p]
            getattr(L, "oracle_cell_index_" + suf).argtypes = [_i64, _p, r, _i32, _p]
        L.oracle_pair_set_brute.restype = _i64
        L.oracle_pair_set_brute.argtypes = [_i64, _p, _d, _d, _p, _p, _p, _i64, _p]
        L.oracle_pair_set_cells.restype = _i64
        L.oracle_pair_set_cells.argtypes = [_i64, _p, _d, _d, C.c_int, _p, _p, _p, _i64]
        L.oracle_cutoff_cells.restype = C.c_int
        L.oracle_cutoff_cells.argtypes = [_i64, _p, _d, _d, _d, _p, C.c_int, _p, _p, C.c_int, _p, _p, _p, _p, _p, _p]
        L.oracle_pairs14_correction.restype = _i64
        L.oracle_pairs14_correction.argtypes = [_i64, _p, _d, _d, _d, _p, _p, _i64, _d, C.c_int, _p, _p, _p, _p]
        L.oracle_vv_steps.restype = C.c_int
        L.oracle_vv_steps.argtypes = [_i64, _p, _p, _p, _p, _d, _d, _d, _p, C.c_int, _p, _p, _d, _i64]
        _LIBS[name] = L
    return _LIBS[name]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_p)


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def num_threads(fast=False):
    return lib(fast).oracle_num_threads()


def set_num_threads(n, fast=False):
    """OpenMP threads of the oracle (torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm wants the host)."""
    lib(fast).oracle_set_num_threads(int(n))


def lj_model(cutoff, switch, dtype=np.float64):
    out = np.empty(3, dtype=dtype)
    getattr(lib(), "oracle_lj_model_" + ("f64" if dtype == np.float64 else "f32"))(cutoff, switch, _ptr(out))
    return out


def lj_atom(eps, sigma, dtype=np.float64):
    out = np.empty(2, dtype=dtype)
    getattr(lib(), "oracle_lj_atom_" + ("f64" if dtype == np.float64 else "f32"))(eps, sigma, _ptr(out))
    return out


def interaction(r2, model, ai, aj):
    dtype = model.dtype
    suf = "f64" if dtype == np.float64 else "f32"
    out = np.empty(2, dtype=dtype)
    getattr(lib(), "oracle_interaction_" + suf)(dtype.type(r2), _ptr(_arr(model, dtype)), _ptr(_arr(ai, dtype)),
                                                _ptr(_arr(aj, dtype)), _ptr(out))
    return out


def tiles(N):
    n = lib().oracle_tiles(N, None)
    out = np.empty((n, 2), dtype=np.int32)
    lib().oracle_tiles(N, _ptr(out))
    return out


def naive_allpairs(pos, L, model, atoms):
    dtype = pos.dtype
    suf = "f64" if dtype == np.float64 else "f32"
    pos = _arr(pos, dtype); atoms = _arr(atoms, dtype); model = _arr(model, dtype)
    N = pos.shape[0]
    f = np.empty((N, 3), dtype=dtype); e = np.empty(N, dtype=dtype); w = np.empty(N, dtype=dtype)
    getattr(lib(), "oracle_naive_allpairs_" + suf)(N, _ptr(pos), dtype.type(L), _ptr(model), _ptr(atoms),
                                                   _ptr(f), _ptr(e), _ptr(w))
    return f, e, w


def tiles_allpairs(pos, L, tile_list, model, atoms, bitmask=7):
    dtype = pos.dtype
    suf = "f64" if dtype == np.float64 else "f32"
    pos = _arr(pos, dtype); atoms = _arr(atoms, dtype); model = _arr(model, dtype)
    tile_list = _arr(tile_list, np.int32)
    N = pos.shape[0]
    f = np.zeros((N, 3), dtype=dtype); e = np.zeros(N, dtype=dtype); w = np.zeros(N, dtype=dtype)
    getattr(lib(), "oracle_tiles_allpairs_" + suf)(N, _ptr(pos), dtype.type(L), _ptr(tile_list), tile_list.shape[0],
                                                   _ptr(model), _ptr(atoms), bitmask, _ptr(f), _ptr(e), _ptr(w))
    return f, e, w


def cells_per_dimension(L, cutoff, ndiv):
    return int(lib().oracle_cells_per_dimension(L, cutoff, ndiv))


def cell_index(pos, L, M):
    dtype = pos.dtype
    suf = "f64" if dtype == np.float64 else "f32"
    pos = _arr(pos, dtype)
    out = np.empty(pos.shape[0], dtype=np.int32)
    getattr(lib(), "oracle_cell_index_" + suf)(pos.shape[0], _ptr(pos), dtype.type(L), M, _ptr(out))
    return out


def _excl(excl):
    if excl is None:
        return None, None
    return _arr(excl[0], np.int32), _arr(excl[1], np.uint64)


def pair_set_brute(pos, L, rc2, excl=None):
    pos = _arr(pos, np.float64)
    N = pos.shape[0]
    eb, em = _excl(excl)
    dig = np.zeros(3, dtype=np.uint64)
    n = lib().oracle_pair_set_brute(N, _ptr(pos), L, rc2, _ptr(eb), _ptr(em), None, 0, _ptr(dig))
    ij = np.empty((n, 2), dtype=np.int32)
    lib().oracle_pair_set_brute(N, _ptr(pos), L, rc2, _ptr(eb), _ptr(em), _ptr(ij), n, _ptr(dig))
    return ij, dig


def pair_set_cells(pos, L, cutoff, ndiv=1, excl=None):
    pos = _arr(pos, np.float64)
    N = pos.shape[0]
    eb, em = _excl(excl)
    n = lib().oracle_pair_set_cells(N, _ptr(pos), L, cutoff, ndiv, _ptr(eb), _ptr(em), None, 0)
    ij = np.empty((n, 2), dtype=np.int32)
    lib().oracle_pair_set_cells(N, _ptr(pos), L, cutoff, ndiv, _ptr(eb), _ptr(em), _ptr(ij), n)
    return ij


def _fma(a, b, c):
    """Element-wise a*b + c in extended precision, rounded to double (within one ulp of the fma() the C oracle uses)."""
    return (np.asarray(a, dtype=np.longdouble) * np.asarray(b, dtype=np.longdouble) + np.asarray(c, dtype=np.longdouble)).astype(np.float64)


def cutoff_cells(pos, L, cutoff, switch, atoms, ndiv=1, excl=None, bitmask=7, fast=False, pairs14=None):
    """Returns dict(forces, energies, virials, E, W, npairs, digest).  pairs14 = (ij (n,2) int32 with i<j, scale): pairs three
    bonds apart interact with `scale` times the ordinary pair's E / W / force (oracle_pairs14_correction); the pair set and
    its digest are those of the unscaled evaluation."""
    pos = _arr(pos, np.float64); atoms = _arr(atoms, np.float64)
    N = pos.shape[0]
    eb, em = _excl(excl)
    f = np.zeros((N, 3)); e = np.zeros(N); w = np.zeros(N)
    tot = np.zeros(2); npairs = _i64(0); dig = np.zeros(3, dtype=np.uint64)
    rc = lib(fast).oracle_cutoff_cells(N, _ptr(pos), L, cutoff, switch, _ptr(atoms), ndiv, _ptr(eb), _ptr(em), bitmask,
                                       _ptr(f), _ptr(e), _ptr(w), _ptr(tot), C.byref(npairs), _ptr(dig))
    if rc:
        raise RuntimeError("oracle_cutoff_cells failed with status %d" % rc)
    n14 = None
    if pairs14 is not None:
        ij = _arr(pairs14[0], np.int32).reshape(-1, 2)
        n14 = int(lib(fast).oracle_pairs14_correction(N, _ptr(pos), L, cutoff, switch, _ptr(atoms), _ptr(ij), ij.shape[0], float(pairs14[1]),
                                                     bitmask, _ptr(f), _ptr(e), _ptr(w), _ptr(tot)))
    return dict(forces=f, energies=e, virials=w, E=float(tot[0]), W=float(tot[1]), npairs=int(npairs.value), digest=dig, n14_inside=n14)


def vv_steps(pos, vel, forces, mass, L, cutoff, switch, atoms, dt, nsteps, ndiv=1, excl=None, fast=False, pairs14=None):
    """In-place velocity-Verlet on copies; returns (pos, vel, forces)."""
    pos = np.array(pos, dtype=np.float64, order="C"); vel = np.array(vel, dtype=np.float64, order="C")
    forces = np.array(forces, dtype=np.float64, order="C")
    mass = _arr(mass, np.float64); atoms = _arr(atoms, np.float64)
    if pairs14 is not None:      # same update as oracle_vv_steps (one fma per line), forces with the 1-4 correction
        h = (0.5 * dt / mass)[:, None]
        for _ in range(int(nsteps)):
            vel = _fma(h, forces, vel)
            pos = _fma(dt, vel, pos)
            forces = cutoff_cells(pos, L, cutoff, switch, atoms, ndiv=ndiv, excl=excl, bitmask=1, fast=fast, pairs14=pairs14)["forces"]
            vel = _fma(h, forces, vel)
        return pos, vel, forces
    eb, em = _excl(excl)
    rc = lib(fast).oracle_vv_steps(pos.shape[0], _ptr(pos), _ptr(vel), _ptr(forces), _ptr(mass), L, cutoff, switch,
                                   _ptr(atoms), ndiv, _ptr(eb), _ptr(em), dt, nsteps)
    if rc:
        raise RuntimeError("oracle_vv_steps failed with status %d" % rc)
    return pos, vel, forces
