"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/emdee_b200.h
declares, and fails loudly (no CPU fallback) when no B200 is present.  Host-side logic of the mirror."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "emdee_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(emdee_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(em):
    em.build_library()
    lib = ctypes.CDLL(em.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 35
    for name in syms:
        assert hasattr(lib, name), "libemdee_b200.so does not export %s" % name
    # and the ctypes binding declares every one of them
    from emdee_jl_b200 import _lib
    bound = set(_lib.SIGNATURES) | set(_lib.NON_STATUS)
    assert set(syms) == bound


def test_every_entry_point_is_documented():
    """INTEGRATION.md names every entry point the header declares (section 6: the reference symbol each one replaces)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [name for name in declared_symbols() if "`%s`" % name not in text]
    assert not missing, missing


def test_no_cpu_fallback(em):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(em.EmDeeError) as ei:
        em.Context(0)
    assert ei.value.status == 2 and "no CPU fallback" in str(ei.value)
    f = np.zeros((32, 3)); e = np.zeros(32); w = np.zeros(32)
    with pytest.raises(em.EmDeeError):
        em.compute_nonbonded_(f, e, w, np.random.rand(32, 3), 10.0, em.nonbonded_computation_tiles(32),
                              em.LennardJonesModel(3, 2.5), np.tile(em.LennardJonesAtom(1, 1), (32, 1)), 7)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "emdee.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_c" not in text and "liboracle" not in text and "import oracle" not in text, f


def test_host_mirror(em, oracle):
    m = em.LennardJonesModel(3, 2.5)
    assert (m.rc2, m.rs2, m.inv_delta2) == tuple(oracle.lj_model(3.0, 2.5))
    assert np.array_equal(em.LennardJonesAtom(1, 1), oracle.lj_atom(1, 1))
    assert np.array_equal(em.LennardJonesAtom(0.25, 3.0), oracle.lj_atom(0.25, 3.0))
    assert (em.FORCES, em.ENERGIES, em.VIRIALS, em.WARPSIZE) == (1, 2, 4, 32)
    for N in (1, 31, 32, 33, 800, 4000):
        assert np.array_equal(em.nonbonded_computation_tiles(N), oracle.tiles(N))
    with pytest.raises(ValueError):
        em.LennardJonesModel(2.5, 3.0)


def test_julia_shim_binds_declared_symbols():
    """Every ccall in julia/EmDee.jl names a symbol the header declares."""
    path = os.path.join(ROOT, "julia", "EmDee.jl")
    text = open(path).read()
    called = set(re.findall(r"ccall\(\(:(emdee_[a-z0-9_]+),", text))
    assert called and called <= set(declared_symbols())


def test_header_is_plain_c_and_links(em, tmp_path):
    """include/emdee_b200.h compiles as strict C11 (what cgo / ccall / ctypes see), examples/abi_example.c links against
    the shared library, runs, and -- on a box without a B200 -- reports EMDEE_ERR_CUDA instead of computing on the CPU."""
    import shutil
    import subprocess

    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if not gcc:
        pytest.skip("no C compiler")
    em.build_library()
    exe = str(tmp_path / "abi_example")
    libdir = os.path.dirname(em.LIB_PATH)
    subprocess.check_call([gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "abi_example.c"), "-o", exe, "-L", libdir, "-lemdee_b200",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    import torch

    if torch.cuda.is_available():
        assert out.returncode == 0 and "cutoff: E" in out.stdout, out.stdout + out.stderr
    else:
        assert out.returncode == 3 and "status 2" in out.stderr and "libemdee_b200 version" in out.stdout, out.stdout + out.stderr
