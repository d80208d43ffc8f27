"""XYZ frames and restart files (emdee.jl_b200/trajectory.py).  CPU only."""
import os

import numpy as np
import pytest

REF_XYZ = "/root/reference/test/data/lj_sample.xyz"


def test_xyz_round_trip_and_frames(em, tmp_path):
    from emdee_jl_b200 import trajectory as tj

    rng = np.random.default_rng(1)
    a, b = rng.normal(size=(7, 3)) * 10, rng.normal(size=(7, 3)) * 1e-3
    path = str(tmp_path / "t.xyz")
    w = tj.XYZWriter(path, precision=17)
    w.write(a, "step 0")
    w.write(b, "step\n1")
    assert w.frames == 2 and tj.count_xyz_frames(path) == 2
    names, p0, c0 = tj.read_xyz(path)
    _, p1, c1 = tj.read_xyz(path, frame=1)
    assert names == [str(i + 1) for i in range(7)] and (c0, c1) == ("step 0", "step 1")
    assert np.array_equal(p0, a) and np.array_equal(p1, b)                  # 17 significant digits: exact
    with pytest.raises(IndexError):
        tj.read_xyz(path, frame=2)
    with pytest.raises(ValueError):
        tj.XYZWriter(str(tmp_path / "u.xyz"), names=["Ar"] * 3).write(a)
    # the layout of the reference's fixture: count, EMPTY comment line, "index x y z" in %.12E
    ref_like = tmp_path / "r.xyz"
    ref_like.write_text(" 2\n\n 1 -1.126362593256E-01 1.385093082507E+00 -8.842035145736E-01\n 2 -2.463715052470E+00 -1.375803142943E+00 8.391336211681E-01\n")
    names, p, c = tj.read_xyz(str(ref_like))
    assert names == ["1", "2"] and c == "" and p[1, 2] == 8.391336211681E-01


def test_checkpoint_file_round_trip(em, tmp_path):
    from emdee_jl_b200 import trajectory as tj

    rng = np.random.default_rng(2)
    ck = dict(N=5, L=12.5, positions=rng.normal(size=(5, 3)), velocities=rng.normal(size=(5, 3)))
    path = str(tmp_path / "restart.npz")
    tj.save_checkpoint(path, ck, step=np.int64(1200), time=6.0)
    back = tj.load_checkpoint(path)
    assert back["N"] == 5 and back["L"] == 12.5 and int(back["step"]) == 1200 and float(back["time"]) == 6.0
    assert np.array_equal(back["positions"], ck["positions"]) and np.array_equal(back["velocities"], ck["velocities"])
    bad = dict(ck, N=6)
    tj.save_checkpoint(path, bad)
    with pytest.raises(ValueError):
        tj.load_checkpoint(path)


@pytest.mark.skipif(not os.path.exists(REF_XYZ), reason="the reference tree exists in the build container only")
def test_reads_the_reference_fixture(em, lj_sample):
    from emdee_jl_b200 import trajectory as tj

    names, p, _ = tj.read_xyz(REF_XYZ)
    assert len(names) == 800 and np.array_equal(p, lj_sample["positions"])
