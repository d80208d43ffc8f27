"""make_golden.py -- regenerates tests/golden/*.npz.  Run in the build container only (it reads
/root/reference, which does not exist on the GPU box):  python tests/golden/make_golden.py

Inputs are the reference's only fixtures for this path:
  test/data/lj_sample.xyz                      800-atom LJ configuration (test/runtests.jl:58: L=10, rc=3, rs=2.5)
  test/data/dibenzo-p-dioxin-in-water.{pdb,xml} 1519 atoms / 500 residues (test/runtests.jl:48), config 4 input
The reference holds NO stored numeric answers (its test is a GPU-vs-CPU self-consistency check), so the
"expected" numbers stored here come from the CPU oracle after it was cross-checked against the independent
numpy twin (oracle/oracle_np.py) -- "parity unpinned" against reference binaries, see oracle/emdee_oracle.c.
"""
import os
import re
import sys
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/test/data"

from oracle import oracle_c as oc  # noqa: E402
from oracle import oracle_np as on  # noqa: E402


def lj_sample():
    xyz = np.loadtxt(os.path.join(REF, "lj_sample.xyz"), skiprows=2)[:, 1:4]
    assert xyz.shape == (800, 3)
    L, rc, rs = 10.0, 3.0, 2.5
    model = oc.lj_model(rc, rs)
    atoms = np.tile(oc.lj_atom(1, 1), (800, 1))
    f, e, w = oc.naive_allpairs(xyz, L, model, atoms)
    f2, e2, w2 = on.naive_allpairs_f64(xyz, L, model, atoms)
    assert np.array_equal(f, f2) and np.array_equal(e, e2) and np.array_equal(w, w2), "C oracle != numpy twin"
    cut = oc.cutoff_cells(xyz, L, rc, rs, atoms, ndiv=1)
    fc, ec, wc, ij = on.cutoff_compute(xyz, L, rc, rs, atoms)
    pairs, dig = oc.pair_set_brute(xyz, L, rc * rc)
    assert np.array_equal(ij, pairs) and np.array_equal(on.pair_digest(ij), dig) and np.array_equal(dig, cut["digest"])
    assert np.abs(fc - cut["forces"]).max() < 1e-12 and np.abs(ec - cut["energies"]).max() < 1e-12
    np.savez_compressed(
        os.path.join(HERE, "lj_sample.npz"), positions=xyz, L=L, cutoff=rc, switch=rs,
        allpairs_forces=f, allpairs_energies=e, allpairs_virials=w,
        cutoff_forces=cut["forces"], cutoff_energies=cut["energies"], cutoff_virials=cut["virials"],
        cutoff_pairs=pairs, cutoff_digest=dig,
        cell_index_ndiv1=oc.cell_index(xyz, L, oc.cells_per_dimension(L, rc, 1)),
        cell_index_ndiv2=oc.cell_index(xyz, L, oc.cells_per_dimension(L, rc, 2)),
        tiles=oc.tiles(800))
    print("lj_sample: sumE allpairs %.15g cutoff %.15g pairs %d" % (e.sum(), cut["E"], pairs.shape[0]))


def dioxin_water():
    pdb = open(os.path.join(REF, "dibenzo-p-dioxin-in-water.pdb")).read().splitlines()
    box = None
    names, resn, resi, pos, serial = [], [], [], [], []
    bonds = set()
    for ln in pdb:
        if ln.startswith("CRYST1"):
            box = float(ln[6:15])
        elif ln.startswith(("ATOM", "HETATM")):
            serial.append(int(ln[6:11])); names.append(ln[12:16].strip()); resn.append(ln[17:20].strip())
            resi.append(int(ln[22:26])); pos.append([float(ln[30:38]), float(ln[38:46]), float(ln[46:54])])
        elif ln.startswith("CONECT"):
            f = [int(x) for x in re.findall(r"\d+", ln[6:])]
            for b in f[1:]:
                bonds.add((min(f[0], b), max(f[0], b)))
    pos = np.array(pos)
    sid = {s: k for k, s in enumerate(serial)}
    bonds = np.array(sorted((sid[a], sid[b]) for a, b in bonds), dtype=np.int32)
    resi = np.array(resi)
    assert len(names) == 1519 and len(np.unique(resi)) == 500          # test/runtests.jl:48
    assert np.all(np.diff(resi) >= 0), "atoms are already residue-contiguous (src/modelling.jl:330-348)"
    root = ET.parse(os.path.join(REF, "dibenzo-p-dioxin-in-water.xml")).getroot()
    tmass = {t.get("name"): float(t.get("mass")) for t in root.find("AtomTypes")}
    tname = {}
    for r in root.find("Residues"):
        for a in r.findall("Atom"):
            tname[(r.get("name"), a.get("name"))] = a.get("type")
    nb = root.find("NonbondedForce")
    lj = {a.get("type"): (float(a.get("sigma")), float(a.get("epsilon"))) for a in nb.findall("Atom")}
    types = sorted(tmass)
    tidx = np.array([types.index(tname[(rn, an)]) for rn, an in zip(resn, names)], dtype=np.int32)
    # the product's own readers (emdee.jl_b200/modelling.py) must arrive at the same arrays as this inline parser
    from emdee_jl_b200 import modelling as md
    from emdee_jl_b200 import workloads as wl
    ff = md.ForceField(os.path.join(REF, "dibenzo-p-dioxin-in-water.xml"))
    fx = md.System(os.path.join(REF, "dibenzo-p-dioxin-in-water.pdb"), ff).fixture(ff)
    assert np.array_equal(fx["positions"], pos) and np.array_equal(fx["bonds"], bonds) and np.array_equal(fx["type_index"], tidx)
    # CUTOFF evaluation of the single box (rc = 10 A, rs = 9 A, sigma nm -> A, 1-2/1-3 exclusions): C oracle, cross-checked
    # against the numpy twin (exact pair set, forces / energies to rounding) before it is stored
    sig = np.array([lj[t][0] for t in types])[tidx] * 10.0
    eps = np.array([lj[t][1] for t in types])[tidx]
    atoms = np.stack([0.5 * sig, 2.0 * np.sqrt(eps)], axis=1)
    base, mask = wl.exclusion_masks(len(names), bonds)
    cut = oc.cutoff_cells(pos, box, 10.0, 9.0, atoms, ndiv=1, excl=(base, mask))
    f2, e2, w2, ij = on.cutoff_compute(pos, box, 10.0, 9.0, atoms, base, mask)
    assert np.array_equal(on.pair_digest(ij), cut["digest"]) and ij.shape[0] == cut["npairs"]
    assert np.abs(f2 - cut["forces"]).max() < 1e-11 and np.abs(e2 - cut["energies"]).max() < 1e-12
    # the same evaluation with the force field's lj14scale applied to the pairs three bonds apart (C oracle vs numpy twin)
    p14 = wl.pairs14(len(names), bonds)
    lj14 = float(nb.get("lj14scale"))
    cut14 = oc.cutoff_cells(pos, box, 10.0, 9.0, atoms, ndiv=1, excl=(base, mask), pairs14=(p14, lj14))
    df, de, dw, n14 = on.pairs14_correction(pos, box, 10.0, 9.0, atoms, p14, lj14)
    assert n14 == cut14["n14_inside"] == p14.shape[0] and np.array_equal(cut14["digest"], cut["digest"])
    assert np.abs(f2 + df - cut14["forces"]).max() < 1e-11 and np.abs(e2 + de - cut14["energies"]).max() < 1e-12
    assert abs((e2 + de).sum() - cut14["E"]) < 1e-9 and abs((w2 + dw).sum() - cut14["W"]) < 1e-9
    np.savez_compressed(
        os.path.join(HERE, "dioxin_water.npz"), positions=pos, box=box, bonds=bonds, residue=resi.astype(np.int32),
        pairs14=p14, cutoff14_forces=cut14["forces"], cutoff14_E=cut14["E"], cutoff14_W=cut14["W"],
        type_index=tidx, type_names=np.array(types), type_sigma_nm=np.array([lj[t][0] for t in types]),
        type_epsilon=np.array([lj[t][1] for t in types]), type_mass=np.array([tmass[t] for t in types]),
        lj14scale=float(nb.get("lj14scale")),
        cutoff_forces=cut["forces"], cutoff_E=cut["E"], cutoff_W=cut["W"], cutoff_npairs=np.int64(cut["npairs"]),
        cutoff_digest=cut["digest"], excl_base=base, excl_mask=mask)
    print("dioxin_water: %d atoms, %d bonds, box %.3f, types %s" % (len(names), len(bonds), box, types))


if __name__ == "__main__":
    lj_sample()
    dioxin_water()
