"""_lib.py -- loads libemdee_b200.so (the C ABI of include/emdee_b200.h) with ctypes.

There is no CPU fallback: if the CUDA library is missing or cannot be loaded the import of the
product API fails loudly, and every compute entry point fails with EMDEE_ERR_CUDA without a B200.
The library is built in-tree (csrc/libemdee_b200.so) by build_library(), i.e. plain
`nvcc -gencode arch=compute_100a,code=sm_100a`; nvcc cross-compiles without a GPU.
"""
import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libemdee_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "emdee_b200.h")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


class EmDeeError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("emdee_b200 status %d: %s" % (status, message))
        self.status = status


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def build_library(force=False, verbose=False):
    """Compile csrc/emdee_b200.cu for sm_100a into csrc/libemdee_b200.so (skipped when up to date)."""
    srcs = _sources() + [HEADER]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, os.path.join(CSRC, "emdee_b200.cu")]
    cmd += ["-ldl"]      # NCCL is dlopen'ed at run time (see csrc/emdee_b200.cu), never linked
    subprocess.check_call(cmd)
    return LIB_PATH


_p = C.c_void_p
_d = C.c_double
_i = C.c_int
_i64 = C.c_int64

# name -> (argtypes); every function returns int status except the two noted below
SIGNATURES = {
    "emdee_create": [C.POINTER(_p), _i],
    "emdee_destroy": [_p],
    "emdee_comm_unique_id": [C.c_char_p],
    "emdee_comm_init": [_p, _i, _i, C.c_char_p],
    "emdee_device_info": [_p, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i64)],
    "emdee_measure_fp64_peak": [_p, C.POINTER(_d), C.POINTER(_d)],
    "emdee_system_create": [_p, _i64, _d, C.POINTER(_p)],
    "emdee_system_destroy": [_p],
    "emdee_set_model": [_p, _d, _d],
    "emdee_set_lj_atoms": [_p, _p],
    "emdee_set_positions": [_p, _p],
    "emdee_set_velocities": [_p, _p],
    "emdee_set_masses": [_p, _p],
    "emdee_set_exclusions": [_p, _p, _p],
    "emdee_set_pairs14": [_p, _p, _i64, _d],
    "emdee_get_local_id_range": [_p, C.POINTER(_i64), C.POINTER(_i64)],
    "emdee_set_positions_range": [_p, _i64, _i64, _p],
    "emdee_get_forces_range": [_p, _i64, _i64, _p],
    "emdee_get_energies_range": [_p, _i64, _i64, _p],
    "emdee_get_virials_range": [_p, _i64, _i64, _p],
    "emdee_set_skin": [_p, _d],
    "emdee_bin": [_p, _i],
    "emdee_update_cells": [_p, C.POINTER(_i64)],
    "emdee_get_cells_per_dimension": [_p, C.POINTER(C.c_int32)],
    "emdee_get_cell_index": [_p, _p],
    "emdee_get_cell_population": [_p, _p],
    "emdee_get_cell_order": [_p, _p, _p],
    "emdee_compute_nonbonded": [_p, _i, _i],
    "emdee_compute_nonbonded_into": [_p, _i, _i, _p, _p, _p],
    "emdee_get_step_counters": [_p, _p],
    "emdee_set_tiles": [_p, _p, _i64],
    "emdee_get_positions": [_p, _p],
    "emdee_get_velocities": [_p, _p],
    "emdee_get_forces": [_p, _p],
    "emdee_get_energies": [_p, _p],
    "emdee_get_virials": [_p, _p],
    "emdee_get_totals": [_p, C.POINTER(_d), C.POINTER(_d), C.POINTER(_i64)],
    "emdee_pair_set": [_p, _p, _i64, C.POINTER(_i64)],
    "emdee_pair_set_digest": [_p, _p],
    "emdee_list_pair_count": [_p, _p],
    "emdee_profile_kind": [_p, _i, _p, _p],
    "emdee_fp16_threshold": [_p, _d, _p],
    "emdee_vv_step": [_p, _d, _i64, _i],
    "emdee_kinetic_energy": [_p, C.POINTER(_d)],
    "emdee_scale_velocities": [_p, _d],
    "emdee_synchronize": [_p],
    "emdee_launch_count": [_p, C.POINTER(_i64)],
    "emdee_timer_start": [_p],
    "emdee_timer_stop": [_p, C.POINTER(_d)],
    "emdee_profile_begin": [_p],
    "emdee_profile_end": [_p, C.POINTER(_d), C.POINTER(_i64)],
    "emdee_get_step_config": [_p, _p],
    "emdee_get_local_count": [_p, C.POINTER(_i64), C.POINTER(_i64)],
    "emdee_get_local_ids": [_p, _p],
    "emdee_compute_nonbonded_host": [_i64, _p, _d, _d, _d, _p, _p, _i64, _i, _i, _i, _p, _p, _p],
}
NON_STATUS = {"emdee_version": (_i, []), "emdee_last_error": (C.c_char_p, [])}

_lib = None


def load():
    """Load the shared library and declare every entry point of include/emdee_b200.h."""
    global _lib
    if _lib is None:
        path = os.environ.get("EMDEE_B200_LIB", LIB_PATH)      # the same override the Julia shim reads
        if not os.path.exists(path):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback for the nonbonded path)" % path)
        lib = C.CDLL(path)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = _i
            fn.argtypes = args
        for name, (res, args) in NON_STATUS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status):
    if status != 0:
        raise EmDeeError(status, load().emdee_last_error().decode("utf-8", "replace"))


def call(name, *args):
    check(getattr(load(), name)(*args))
