"""Host-side bridge ForceField / System (emdee.jl_b200/modelling.py <- src/modelling.jl:145-349): force-field XML and PDB
readers, residue-template matching by coloured-graph isomorphism, residue-contiguous re-ordering, and the arrays the
nonbonded path takes.  CPU only."""
import os

import numpy as np
import pytest

XML = """<ForceField>
  <AtomTypes>
    <Type name="HW" class="HW" element="H" mass="1.008"/>
    <Type name="OW" class="OW" element="O" mass="15.9994"/>
    <Type name="CT" class="CT" element="C" mass="12.01"/>
    <Type name="HC" class="HC" element="H" mass="1.008"/>
    <Type name="OH" class="OH" element="O" mass="16.0"/>
    <Type name="HO" class="HO" element="H" mass="1.008"/>
  </AtomTypes>
  <Residues>
    <Residue name="HOH">
      <Atom name="H1" type="HW" charge="0.417"/>
      <Atom name="O" type="OW" charge="-0.834"/>
      <Atom name="H2" type="HW" charge="0.417"/>
      <Bond atomName1="O" atomName2="H1"/>
      <Bond from="1" to="2"/>
    </Residue>
    <Residue name="MOH">
      <Atom name="C" type="CT" charge="0.1"/>
      <Atom name="HA" type="HC" charge="0.04"/>
      <Atom name="HB" type="HC" charge="0.04"/>
      <Atom name="HC" type="HC" charge="0.04"/>
      <Atom name="O" type="OH" charge="-0.6"/>
      <Atom name="HO" type="HO" charge="0.38"/>
      <Bond atomName1="C" atomName2="HA"/>
      <Bond atomName1="C" atomName2="HB"/>
      <Bond atomName1="C" atomName2="HC"/>
      <Bond atomName1="C" atomName2="O"/>
      <Bond atomName1="O" atomName2="HO"/>
      <AllowPatch name="DEPROT"/>
    </Residue>
  </Residues>
  <Patches>
    <Patch name="DEPROT">
      <RemoveAtom name="HO"/>
      <ChangeAtom name="O" type="OH" charge="-1.22"/>
    </Patch>
  </Patches>
  <HarmonicBondForce>
    <Bond type1="CT" type2="HC" length="0.109" k="284512.0"/>
  </HarmonicBondForce>
  <NonbondedForce coulomb14scale="0.833333" lj14scale="0.5">
    <Atom type="HW" sigma="1" epsilon="0"/>
    <Atom type="OW" sigma="0.315" epsilon="0.636"/>
    <Atom type="CT" sigma="0.35" epsilon="0.276"/>
    <Atom type="HC" sigma="0.25" epsilon="0.1255"/>
    <Atom type="OH" sigma="0.312" epsilon="0.711"/>
    <Atom type="HO" sigma="0.1" epsilon="0"/>
  </NonbondedForce>
</ForceField>
"""


def _pdb(atoms, conect, cryst=(20.0, 20.0, 20.0)):
    lines = ["CRYST1%9.3f%9.3f%9.3f  90.00  90.00  90.00 P 1           1" % cryst]
    for serial, name, resname, resseq, (x, y, z), elem in atoms:
        lines.append("HETATM%5d %-4s %3s  %4d    %8.3f%8.3f%8.3f  1.00  0.00          %2s" % (serial, name, resname, resseq, x, y, z, elem))
    for a, b in conect:
        lines.append("CONECT%5d%5d" % (a, b))
    lines.append("END")
    return "\n".join(lines) + "\n"


@pytest.fixture()
def files(tmp_path):
    xml = tmp_path / "ff.xml"
    xml.write_text(XML)
    # residue 1: water; residue 2: methanol whose atoms are written in an order unlike the template's and INTERLEAVED
    # with residue 3 (a second water); residue 4: methoxide (the patched template)
    atoms = [
        (1, "OW", "HOH", 1, (1.0, 1.0, 1.0), "O"), (2, "HW1", "HOH", 1, (1.8, 1.0, 1.0), "H"), (3, "HW2", "HOH", 1, (0.8, 1.7, 1.0), "H"),
        (4, "HX", "MOH", 2, (5.0, 5.0, 5.9), "H"),        # the hydroxyl hydrogen first
        (5, "O1", "MOH", 2, (5.0, 5.0, 5.0), "O"),
        (6, "H1", "HOH", 3, (9.0, 9.0, 9.0), "H"), (7, "O", "HOH", 3, (9.5, 9.5, 9.0), "O"), (8, "H2", "HOH", 3, (10.0, 9.0, 9.0), "H"),
        (9, "C1", "MOH", 2, (6.2, 5.0, 4.4), "C"), (10, "H1'", "MOH", 2, (6.2, 5.9, 3.9), "H"), (11, "H2*", "MOH", 2, (7.0, 5.0, 5.1), "H"),
        (12, "H-3", "MOH", 2, (6.3, 4.1, 3.8), "H"),
        (13, "C", "MOX", 4, (15.0, 5.0, 4.4), "C"), (14, "O", "MOX", 4, (14.0, 5.0, 5.0), "O"), (15, "HA", "MOX", 4, (15.0, 5.9, 3.9), "H"),
        (16, "HB", "MOX", 4, (15.8, 5.0, 5.1), "H"), (17, "HC", "MOX", 4, (15.1, 4.1, 3.8), "H"),
    ]
    conect = [(1, 2), (1, 3), (5, 4), (5, 9), (9, 10), (9, 11), (9, 12), (7, 6), (7, 8), (13, 14), (13, 15), (13, 16), (13, 17)]
    pdb = tmp_path / "sys.pdb"
    pdb.write_text(_pdb(atoms, conect))
    return str(xml), str(pdb), tmp_path


def test_force_field_tables(em, files):
    from emdee_jl_b200 import modelling as md

    ff = md.ForceField(files[0])
    assert [t["name"] for t in ff.atom_types] == ["HW", "OW", "CT", "HC", "OH", "HO"]
    assert list(ff.templates) == ["HOH", "MOH", "MOH(DEPROT)"]
    assert (ff.lj14, ff.coulomb14) == (0.5, 0.833333)
    assert ff.lj_by_type()["OW"] == (0.315, 0.636) and len(ff.bond_types) == 1 and ff.angle_types == []
    hoh = ff.templates["HOH"]
    assert [a["name"] for a in hoh.atoms] == ["H1", "O", "H2"]
    assert hoh.adjacency.tolist() == [[False, True, False], [True, False, True], [False, True, False]]   # name and index bond forms
    pat = ff.templates["MOH(DEPROT)"]
    assert [a["name"] for a in pat.atoms] == ["C", "HA", "HB", "HC", "O"] and pat.atoms[4]["charge"] == -1.22
    assert pat.adjacency.sum() == 8                              # the removed atom took its bond along
    assert md.sanitized("H-3'*") == "H_3pa"


def test_color_ranks_and_isomorphism(em):
    from emdee_jl_b200 import modelling as md

    # src/molecular_graphs.jl:69-70: sorted masses, a new cell where neighbours differ by more than 0.1
    assert md.color_ranks([15.999, 1.008, 1.0079, 12.011, 16.0]).tolist() == [2, 0, 0, 1, 2]
    path = np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]], dtype=bool)
    assert md.isomorphism(path, [0, 1, 0], path, [0, 1, 0]) is not None
    assert md.isomorphism(path, [0, 1, 0], path, [1, 0, 0]) is None          # the heavy atom is not the centre
    tri = np.ones((3, 3), dtype=bool) & ~np.eye(3, dtype=bool)
    assert md.isomorphism(path, [0, 0, 0], tri, [0, 0, 0]) is None
    p = md.isomorphism(path[[1, 0, 2]][:, [1, 0, 2]], [1, 0, 0], path, [0, 1, 0])
    assert p == [1, 0, 2] or p == [1, 2, 0]


def test_isomorphism_against_brute_force(em):
    """Random small coloured graphs: the backtracking matcher finds a colour-preserving isomorphism exactly when a
    brute-force search over all permutations does, and what it returns is one."""
    import itertools

    from emdee_jl_b200 import modelling as md

    rng = np.random.default_rng(4)
    found = missed = 0
    for trial in range(300):
        n = int(rng.integers(2, 7))
        a = np.triu(rng.random((n, n)) < 0.45, 1)
        a = a | a.T
        ca = rng.integers(0, 2, n).tolist()
        if trial % 2:                                        # an isomorphic copy: relabel a
            perm = rng.permutation(n)
            b = a[np.ix_(perm, perm)]
            cb = [ca[k] for k in perm]
        else:                                                # an unrelated graph with the same colour multiset
            b = np.triu(rng.random((n, n)) < 0.45, 1)
            b = b | b.T
            cb = list(rng.permutation(ca))
        brute = any(all(ca[i] == cb[p[i]] for i in range(n)) and all(a[i, j] == b[p[i], p[j]] for i in range(n) for j in range(n))
                    for p in itertools.permutations(range(n)))
        got = md.isomorphism(a, ca, b, cb)
        assert (got is not None) == brute, (trial, a, b, ca, cb)
        if got is not None:
            found += 1
            assert sorted(got) == list(range(n))
            assert all(ca[i] == cb[got[i]] for i in range(n))
            assert all(a[i, j] == b[got[i], got[j]] for i in range(n) for j in range(n))
        else:
            missed += 1
    assert found > 100 and missed > 50


def test_system_matching_and_order(em, files):
    from emdee_jl_b200 import modelling as md

    ff = md.ForceField(files[0])
    s = md.System(files[1], ff)
    assert len(s) == 17 and s.count_residues() == 4
    assert s.matched == ["HOH", "MOH", "HOH", "MOH(DEPROT)"]
    # residue-contiguous order, residues in order of first appearance (src/modelling.jl:330-345)
    assert s.residue.tolist() == [0] * 3 + [1] * 6 + [2] * 3 + [3] * 5
    assert s.name[3:9] == ["HX", "O1", "C1", "H1p", "H2a", "H_3"]           # sanitised names
    # types by connectivity, not by file order: the hydrogen on the oxygen is HO, the three on the carbon are HC
    assert s.ff_type[:3] == ["OW", "HW", "HW"]
    assert s.ff_type[3:9] == ["HO", "OH", "CT", "HC", "HC", "HC"]
    assert s.ff_type[12:] == ["CT", "OH", "HC", "HC", "HC"] and s.ff_charge[13] == -1.22
    assert np.allclose(s.ff_charge[3:9], [0.38, -0.6, 0.1, 0.04, 0.04, 0.04])
    # bonds relocated with the atoms; positions follow
    assert s.bonds.shape == (13, 2) and [3, 4] in s.bonds.tolist() and [4, 5] in s.bonds.tolist()
    assert s.positions[5].tolist() == [6.2, 5.0, 4.4] and s.location[8] == 5
    # bridge to the nonbonded path
    assert s.box() == 20.0
    atoms = s.lj_atoms(ff, length_scale=10.0)
    assert atoms.shape == (17, 2) and np.allclose(atoms[0], [0.5 * 3.15, 2 * np.sqrt(0.636)]) and atoms[1, 1] == 0.0
    assert np.allclose(s.masses(ff)[:3], [15.9994, 1.008, 1.008])
    base, mask = s.exclusions()
    for i, j in ((0, 1), (1, 2), (3, 5), (6, 8)):            # 1-2 and 1-3 pairs
        assert (int(mask[i]) >> (j - base[i])) & 1 and (int(mask[j]) >> (i - base[j])) & 1
    assert not (int(mask[3]) >> (6 - base[3])) & 1           # HO...HC is 1-4: not excluded
    fx = s.fixture(ff)
    w = em.workloads.molecular_system(fx, reps=2)
    assert w["positions"].shape == (8 * 17, 3) and w["L"] == 40.0 and w["excl"][0][17] == base[0] + 17


def test_system_errors(em, files):
    from emdee_jl_b200 import modelling as md

    xml, pdb, tmp = files
    ff = md.ForceField(xml)
    bad = tmp / "bad.pdb"
    bad.write_text(_pdb([(1, "C", "XXX", 1, (0, 0, 0), "C"), (2, "O", "XXX", 1, (1, 0, 0), "O")], [(1, 2)]))
    with pytest.raises(ValueError, match=r"No force field templates matched residue 1 \(XXX\)"):
        md.System(str(bad), ff)
    # two templates with the same graph: ambiguous unless disambiguated (1-based residue number, as in the reference)
    xml2 = tmp / "ff2.xml"
    xml2.write_text(XML.replace('<Residue name="MOH">', '<Residue name="WAT"><Atom name="H1" type="HW"/><Atom name="O" type="OW"/>'
                                '<Atom name="H2" type="HW"/><Bond from="0" to="1"/><Bond from="1" to="2"/></Residue>\n    <Residue name="MOH">'))
    ff2 = md.ForceField(str(xml2))
    with pytest.raises(ValueError, match="Multiple force field templates"):
        md.System(pdb, ff2)
    with pytest.raises(ValueError, match="Provided disambiguation"):
        md.System(pdb, ff2, disambiguation={1: "MOH", 3: "WAT"})
    s = md.System(pdb, ff2, disambiguation={1: "WAT", 3: "HOH"})
    assert s.matched[0] == "WAT" and s.matched[2] == "HOH"
    tri = tmp / "tri.pdb"
    tri.write_text(_pdb([(1, "O", "HOH", 1, (0, 0, 0), "O")], [], cryst=(10.0, 12.0, 10.0)))
    with pytest.raises(ValueError):
        md.System(str(tri), ff)          # a lone oxygen matches no template
    

REF = "/root/reference/test/data/dibenzo-p-dioxin-in-water"


@pytest.mark.skipif(not os.path.exists(REF + ".pdb"), reason="the reference tree exists in the build container only")
def test_reference_fixture_matches_golden(em, dioxin_water):
    """test/runtests.jl:44-49 (1519 atoms, 500 residues) and the committed fixture tests/golden/dioxin_water.npz."""
    from emdee_jl_b200 import modelling as md

    ff = md.ForceField(REF + ".xml")
    s = md.System(REF + ".pdb", ff)
    assert len(s) == 1519 and s.count_residues() == 500
    assert s.matched[0] == "aaa" and set(s.matched[1:]) == {"HOH"}
    fx = s.fixture(ff)
    g = dioxin_water
    for key in ("positions", "bonds", "type_index", "type_sigma_nm", "type_epsilon", "type_mass", "residue"):
        assert np.array_equal(fx[key], g[key]), key
    assert fx["box"] == float(g["box"]) and fx["lj14scale"] == float(g["lj14scale"]) and list(fx["type_names"]) == list(g["type_names"])
