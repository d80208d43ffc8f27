"""modelling.py -- host-side mirror of `ForceField(xml)` and `System(file, force_field)` (src/modelling.jl:1,145,235),
the on-disk formats on the input side of the nonbonded path.

The reference leans on three native libraries for this: LightXML (OpenMM-style force-field XML), Chemfiles (PDB) and
nauty (canonical labelling, src/molecular_graphs.jl:66-81).  None of them is needed for what the path consumes, so
this is a from-scratch restatement on the Python standard library:

  * the XML reader keeps the tables the reference keeps (atom types, bond/angle/torsion types as attribute records,
    the nonbonded table, residue templates incl. patches, lj14scale / coulomb14scale: src/modelling.jl:145-203);
  * the PDB reader understands CRYST1, ATOM/HETATM and CONECT;
  * template matching is a coloured-graph isomorphism test (backtracking; residues are small) with the reference's
    colouring rule: vertices are sorted by mass and consecutive masses closer than 0.1 share a colour
    (src/molecular_graphs.jl:66-71), and -- as the reference compares canonical adjacency matrices only
    (src/modelling.jl:309) -- colours correspond by RANK, not by mass value;
  * atoms come out residue-contiguous in residue order (src/modelling.jl:330-348) with `ff_type` / `ff_charge`
    per atom (:321-324) and the bond list relocated (:346-348).

Additive, for the nonbonded path: `System.lj_atoms`, `.masses`, `.exclusions`, `.fixture` turn the matched system
into the arrays `NonbondedSystem` takes (per-type sigma/epsilon, src/modelling.jl:71-73,197; exclusions from the
intra-residue bond graph, :297-304 -- the reference builds the adjacency and never uses it).

Not restated (the reference's test system does not touch it): bonds inferred for standard PDB residues from
src/data/pdb_aliases.xml (src/modelling.jl:205-218,262-295) -- an ATOM-record residue without CONECT bonds raises.
"""
import xml.etree.ElementTree as ET
from collections import OrderedDict

import numpy as np

from .workloads import exclusion_masks

# element masses Chemfiles would assign to the atoms of a non-standard residue (only their ORDER and 0.1 gaps matter)
ELEMENT_MASS = {"H": 1.008, "D": 2.014, "He": 4.0026, "Li": 6.94, "B": 10.81, "C": 12.011, "N": 14.007, "O": 15.999,
                "F": 18.998, "Ne": 20.18, "Na": 22.99, "Mg": 24.305, "Al": 26.982, "Si": 28.085, "P": 30.974, "S": 32.06,
                "Cl": 35.45, "Ar": 39.948, "K": 39.098, "Ca": 40.078, "Mn": 54.938, "Fe": 55.845, "Co": 58.933,
                "Ni": 58.693, "Cu": 63.546, "Zn": 65.38, "Se": 78.971, "Br": 79.904, "I": 126.9}


def sanitized(name):
    """src/modelling.jl:86."""
    return name.replace("-", "_").replace("'", "p").replace("*", "a")


def color_ranks(masses, atol=0.1):
    """Colour of every vertex as the reference partitions them (src/molecular_graphs.jl:69-70): sort by mass, start
    a new cell wherever consecutive masses differ by more than atol.  Returns the cell index (rank) per vertex."""
    masses = np.asarray(masses, dtype=np.float64)
    order = np.argsort(masses, kind="stable")
    rank = np.zeros(len(masses), dtype=np.int64)
    r = 0
    for k in range(1, len(order)):
        if abs(masses[order[k]] - masses[order[k - 1]]) > atol:
            r += 1
        rank[order[k]] = r
    return rank


def isomorphism(adj_a, col_a, adj_b, col_b):
    """Colour-preserving isomorphism a -> b of two small graphs (boolean adjacency matrices), or None.
    Returns perm with perm[i] = vertex of b matched to vertex i of a."""
    n = len(col_a)
    if n != len(col_b) or sorted(col_a) != sorted(col_b):
        return None
    adj_a = np.asarray(adj_a, dtype=bool)
    adj_b = np.asarray(adj_b, dtype=bool)
    deg_a, deg_b = adj_a.sum(axis=1), adj_b.sum(axis=1)
    if sorted(zip(col_a, deg_a)) != sorted(zip(col_b, deg_b)):
        return None
    # most constrained first: walk a in BFS order from the highest-degree vertex so partial maps are connected
    order, seen = [], set()
    for start in sorted(range(n), key=lambda v: -deg_a[v]):
        if start in seen:
            continue
        queue = [start]
        seen.add(start)
        while queue:
            v = queue.pop(0)
            order.append(v)
            for u in np.nonzero(adj_a[v])[0]:
                if int(u) not in seen:
                    seen.add(int(u))
                    queue.append(int(u))
    perm = [-1] * n
    used = [False] * n

    def extend(k):
        if k == n:
            return True
        v = order[k]
        for cand in range(n):
            if used[cand] or col_b[cand] != col_a[v] or deg_b[cand] != deg_a[v]:
                continue
            if all(adj_b[cand, perm[u]] == adj_a[v, u] for u in order[:k]):
                perm[v] = cand
                used[cand] = True
                if extend(k + 1):
                    return True
                used[cand] = False
                perm[v] = -1
        return False

    return perm if extend(0) else None


class ResidueTemplate:
    """src/modelling.jl:11-28: the atoms (name, type, charge) of a residue and its symmetric adjacency matrix.
    The reference stores both in nauty's canonical order; here the XML order is kept and matching is done by
    isomorphism, which accepts exactly the same (residue, template) pairs."""

    def __init__(self, atoms, bonds, type_masses):
        self.atoms = list(atoms)                                    # dicts: name, type, charge
        index = {a["name"]: k for k, a in enumerate(self.atoms)}
        n = len(self.atoms)
        self.adjacency = np.zeros((n, n), dtype=bool)
        for a, b in bonds:
            i, j = index[a], index[b]
            self.adjacency[i, j] = self.adjacency[j, i] = True
        self.masses = np.array([type_masses[a["type"]] for a in self.atoms], dtype=np.float64)
        self.colors = color_ranks(self.masses)


class _Residue:
    """Mutable residue under construction (src/modelling.jl:76-84) with the patch actions of :88-132."""

    def __init__(self, atoms=None, bonds=None, external=None):
        self.atoms, self.bonds, self.external = atoms or [], bonds or [], external or []

    def copy(self):
        return _Residue([dict(a) for a in self.atoms], [frozenset(b) for b in self.bonds], list(self.external))

    def AddAtom(self, at):
        self.atoms.append(dict(name=sanitized(at["name"]), type=at["type"], charge=float(at.get("charge", 0))))

    def AddBond(self, at):
        self.bonds.append(frozenset(sanitized(at[k]) for k in ("atomName1", "atomName2")))

    def AddExternalBond(self, at):
        self.external.append(sanitized(at["atomName"]))

    def ChangeAtom(self, at):
        name = sanitized(at["name"])
        for a in self.atoms:
            if a["name"] == name:
                a["charge"] = float(at.get("charge", 0))
                a["type"] = at["type"]
                return

    def RemoveAtom(self, at):
        name = sanitized(at["name"])
        self.atoms = [a for a in self.atoms if a["name"] != name]

    def RemoveBond(self, at):
        bond = frozenset(sanitized(at[k]) for k in ("atomName1", "atomName2"))
        self.bonds = [b for b in self.bonds if b != bond]

    def RemoveExternalBond(self, at):
        name = sanitized(at["atomName"])
        self.external = [a for a in self.external if a != name]

    def template(self, type_masses):
        names = {a["name"] for a in self.atoms}
        bonds = [tuple(b) for b in self.bonds if len(b) == 2 and b <= names]     # a removed atom takes its bonds along
        return ResidueTemplate(self.atoms, bonds, type_masses)


def _records(root, section, item):
    return [dict(e.attrib) for sec in root.findall(section) for e in sec.findall(item)]


class ForceField:
    """ForceField(xml_file) -- src/modelling.jl:30-40,145-203.  Fields as in the reference: atom_types, bond_types,
    angle_types, dihedral_types, improper_types, nonbonded (lists of attribute records instead of DataFrames),
    templates (ordered name -> ResidueTemplate, patched variants as "name(patch)"), lj14 (lj₁₋₄), coulomb14."""

    def __init__(self, xml_file):
        root = ET.parse(xml_file).getroot()
        patches = {}
        for sec in root.findall("Patches"):
            for item in sec.findall("Patch"):
                patches[item.get("name")] = [(child.tag, dict(child.attrib)) for child in item]
        self.atom_types = _records(root, "AtomTypes", "Type")
        for t in self.atom_types:
            t["mass"] = float(t.get("mass", 0))
        self.type_masses = {t["name"]: t["mass"] for t in self.atom_types}
        self.type_index = {t["name"]: k for k, t in enumerate(self.atom_types)}          # 0-based (reference: 1-based)
        self.templates = OrderedDict()
        for sec in root.findall("Residues"):
            for item in sec.findall("Residue"):
                res = _Residue()
                names = []
                for at in item.findall("Atom"):
                    names.append(at.get("name"))
                    res.AddAtom(at.attrib)
                for bond in item.findall("Bond"):
                    a = dict(bond.attrib)
                    if "from" in a or "to" in a:                                          # index form (:169-172)
                        pair = [names[int(a["from"])], names[int(a["to"])]]
                    else:
                        pair = [a["atomName1"], a["atomName2"]]
                    res.AddBond(dict(atomName1=pair[0], atomName2=pair[1]))
                for bond in item.findall("ExternalBond"):
                    a = dict(bond.attrib)
                    if "from" in a:
                        a["atomName"] = names[int(a["from"])]
                    res.AddExternalBond(a)
                name = item.get("name")
                self.templates[name] = res.template(self.type_masses)
                for allow in item.findall("AllowPatch"):
                    patch = allow.get("name")
                    patched = res.copy()
                    for action, attributes in patches[patch]:
                        getattr(patched, action)(attributes)
                    self.templates["%s(%s)" % (name, patch)] = patched.template(self.type_masses)
        self.bond_types = _records(root, "HarmonicBondForce", "Bond")
        self.angle_types = _records(root, "HarmonicAngleForce", "Angle")
        self.dihedral_types = _records(root, "PeriodicTorsionForce", "Proper")
        self.improper_types = _records(root, "PeriodicTorsionForce", "Improper")
        self.nonbonded = _records(root, "NonbondedForce", "Atom")
        for r in self.nonbonded:
            for key in ("charge", "sigma", "epsilon"):
                r[key] = float(r.get(key, 0))
        nb = root.find("NonbondedForce")
        self.lj14 = float(nb.get("lj14scale", 1.0)) if nb is not None else 1.0
        self.coulomb14 = float(nb.get("coulomb14scale", 1.0)) if nb is not None else 1.0

    def lj_by_type(self):
        """type name -> (sigma, epsilon) in the file's units (nm, kJ/mol for OpenMM-style files)."""
        return {r["type"]: (r["sigma"], r["epsilon"]) for r in self.nonbonded}


def read_pdb(path):
    """CRYST1, ATOM/HETATM, CONECT of a PDB file.  Returns dict(cell (3,), angles (3,), serial, name, resname, chain,
    resseq, icode, element, hetero (bool), positions (N,3), bonds set of 0-based (i<j))."""
    out = dict(cell=None, angles=None, serial=[], name=[], resname=[], chain=[], resseq=[], icode=[], element=[],
               hetero=[], positions=[])
    conect = []
    with open(path) as fh:
        for ln in fh:
            rec = ln[:6]
            if rec == "CRYST1":
                out["cell"] = np.array([float(ln[6:15]), float(ln[15:24]), float(ln[24:33])])
                out["angles"] = np.array([float(ln[33:40]), float(ln[40:47]), float(ln[47:54])])
            elif rec in ("ATOM  ", "HETATM"):
                out["serial"].append(int(ln[6:11]))
                out["name"].append(sanitized(ln[12:16].strip()))
                out["resname"].append(ln[17:20].strip())
                out["chain"].append(ln[21:22])
                out["resseq"].append(int(ln[22:26]))
                out["icode"].append(ln[26:27])
                out["positions"].append([float(ln[30:38]), float(ln[38:46]), float(ln[46:54])])
                elem = ln[76:78].strip() if len(ln) >= 78 else ""
                if not elem:
                    elem = "".join(ch for ch in ln[12:14] if ch.isalpha())
                out["element"].append(elem[:1].upper() + elem[1:].lower())
                out["hetero"].append(rec == "HETATM")
            elif rec == "CONECT":
                body = ln[6:].rstrip("\n")
                fields = [body[k:k + 5] for k in range(0, len(body), 5)]
                ids = [int(f) for f in fields if f.strip()]
                conect.append(ids)
    index = {s: k for k, s in enumerate(out["serial"])}
    bonds = set()
    for ids in conect:
        for b in ids[1:]:
            if ids[0] in index and b in index and ids[0] != b:
                i, j = index[ids[0]], index[b]
                bonds.add((min(i, j), max(i, j)))
    out["positions"] = np.array(out["positions"], dtype=np.float64).reshape(-1, 3)
    out["bonds"] = bonds
    return out


class System:
    """System(file, force_field; disambiguation=Dict()) -- src/modelling.jl:235-349.

    Reads the structure, groups the atoms into residues, matches every residue against the force field's templates
    through its intra-residue bond graph, assigns `ff_type` / `ff_charge` from the matched template, and stores the
    atoms residue-contiguously in residue order with the bonds relocated.  Errors carry the reference's messages
    (:314-320).  `disambiguation` maps a 1-based residue number to a template name, as in the reference.

    Fields (all in the new, residue-contiguous atom order): positions (N,3), velocities (N,3; zeros: PDB holds none),
    cell (3,), name, element, ff_type, ff_charge, residue (0-based residue of every atom), residue_names, bonds
    (nb,2) int32 0-based, location (old index -> new index).  len(system) = atoms, count_residues() = residues
    (what test/runtests.jl:44-49 checks: 1519 and 500 on the reference's fixture)."""

    def __init__(self, file, force_field, disambiguation=None):
        disambiguation = dict(disambiguation or {})
        pdb = read_pdb(file)
        n = len(pdb["name"])
        # residues in order of first appearance, keyed like Chemfiles' PDB reader (chain, residue number, insertion code)
        keys, residue_atoms, residue_names = {}, [], []
        for i in range(n):
            key = (pdb["chain"][i], pdb["resseq"][i], pdb["icode"][i])
            if key not in keys:
                keys[key] = len(residue_atoms)
                residue_atoms.append([])
                residue_names.append(pdb["resname"][i])
            residue_atoms[keys[key]].append(i)
        atom_residue = np.empty(n, dtype=np.int64)
        internal = np.empty(n, dtype=np.int64)
        for r, lst in enumerate(residue_atoms):
            atom_residue[lst] = r
            internal[lst] = np.arange(len(lst))
        bonds = sorted(pdb["bonds"])
        for r, lst in enumerate(residue_atoms):
            if not all(pdb["hetero"][i] for i in lst) and not any(atom_residue[a] == r == atom_residue[b] for a, b in bonds):
                if len(lst) > 1:
                    raise NotImplementedError("residue %d (%s) is a standard PDB residue without CONECT records: bonds from "
                                              "pdb_aliases.xml (src/modelling.jl:262-295) are not restated" % (r + 1, residue_names[r]))
        adjacency = [np.zeros((len(lst), len(lst)), dtype=bool) for lst in residue_atoms]
        for a, b in bonds:                                                  # intra-residue bonds only (:297-304)
            r = atom_residue[a]
            if atom_residue[b] == r:
                i, j = internal[a], internal[b]
                adjacency[r][i, j] = adjacency[r][j, i] = True
        ff_type = [None] * n
        ff_charge = np.zeros(n)
        self.matched = []
        cache = {}
        for r, lst in enumerate(residue_atoms):
            masses = []
            for i in lst:
                if pdb["element"][i] not in ELEMENT_MASS:
                    raise ValueError("atom %d (%s): unknown element %r" % (i + 1, pdb["name"][i], pdb["element"][i]))
                masses.append(ELEMENT_MASS[pdb["element"][i]])
            colors = color_ranks(masses)
            sig = (adjacency[r].tobytes(), tuple(colors))
            if sig not in cache:
                found = []
                for tname, t in force_field.templates.items():
                    perm = isomorphism(adjacency[r], list(colors), t.adjacency, list(t.colors))
                    if perm is not None:
                        found.append((tname, perm))
                cache[sig] = found
            matches = cache[sig]
            resid, name = r + 1, residue_names[r]
            if not matches:
                raise ValueError("No force field templates matched residue %d (%s)" % (resid, name))
            if len(matches) > 1:
                names = [m[0] for m in matches]
                if resid not in disambiguation:
                    raise ValueError("Multiple force field templates %s matched residue %d (%s)" % (names, resid, name))
                if disambiguation[resid] not in names:
                    raise ValueError("Provided disambiguation for residue %d (%s) is not in %s" % (resid, name, names))
                matches = [m for m in matches if m[0] == disambiguation[resid]]
            tname, perm = matches[0]
            template = force_field.templates[tname]
            for k, i in enumerate(lst):
                ff_type[i] = template.atoms[perm[k]]["type"]
                ff_charge[i] = template.atoms[perm[k]]["charge"]
            self.matched.append(tname)
        # residue-contiguous order (:330-345)
        order = [i for lst in residue_atoms for i in lst]
        location = np.empty(n, dtype=np.int64)
        location[order] = np.arange(n)
        self.location = location
        self.positions = np.ascontiguousarray(pdb["positions"][order])
        self.velocities = np.zeros_like(self.positions)
        self.cell = pdb["cell"]
        self.cell_angles = pdb["angles"]
        self.name = [pdb["name"][i] for i in order]
        self.element = [pdb["element"][i] for i in order]
        self.ff_type = [ff_type[i] for i in order]
        self.ff_charge = ff_charge[order]
        self.residue = atom_residue[order].astype(np.int32)
        self.residue_names = residue_names
        self.bonds = np.array(sorted((min(location[a], location[b]), max(location[a], location[b])) for a, b in bonds),
                              dtype=np.int32).reshape(-1, 2)

    def __len__(self):
        return self.positions.shape[0]

    def count_residues(self):
        return len(self.residue_names)

    # ---- bridge to the nonbonded path (additive) ---------------------------------------------------------------
    def box(self):
        """Edge of the cubic box (the path has a scalar L everywhere, e.g. src/nonbonded.jl:60,70)."""
        if self.cell is None:
            raise ValueError("the structure file has no CRYST1 record")
        if not (np.allclose(self.cell, self.cell[0]) and np.allclose(self.cell_angles, 90.0)):
            raise ValueError("the nonbonded path needs a cubic box; CRYST1 gives %s / %s" % (self.cell, self.cell_angles))
        return float(self.cell[0])

    def lj_atoms(self, force_field, length_scale=1.0):
        """(N,2) array of LennardJonesAtom(epsilon, sigma*length_scale) = (sigma/2, 2 sqrt(epsilon)) per atom
        (src/lennard_jones.jl:13); length_scale = 10 converts an OpenMM-style file's nm to the PDB's Angstrom."""
        lj = force_field.lj_by_type()
        sig = np.array([lj[t][0] for t in self.ff_type]) * length_scale
        eps = np.array([lj[t][1] for t in self.ff_type])
        return np.stack([0.5 * sig, 2.0 * np.sqrt(eps)], axis=1)

    def masses(self, force_field):
        return np.array([force_field.type_masses[t] for t in self.ff_type])

    def exclusions(self, max_distance=2):
        """Exclusion bitmasks (base int32, mask uint64) from the bond graph: pairs at graph distance 1..max_distance."""
        return exclusion_masks(len(self), self.bonds, max_distance)

    def fixture(self, force_field):
        """The arrays workloads.molecular_system replicates (same keys as tests/golden/dioxin_water.npz)."""
        types = sorted(force_field.type_masses)
        lj = force_field.lj_by_type()
        return dict(positions=self.positions, box=self.box(), bonds=self.bonds, residue=self.residue + 1,
                    type_index=np.array([types.index(t) for t in self.ff_type], dtype=np.int32), type_names=np.array(types),
                    type_sigma_nm=np.array([lj[t][0] for t in types]), type_epsilon=np.array([lj[t][1] for t in types]),
                    type_mass=np.array([force_field.type_masses[t] for t in types]), lj14scale=force_field.lj14)
