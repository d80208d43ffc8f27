"""trajectory.py -- file formats on the output side of the path (host-side, additive).

The reference reads its LJ fixture through Chemfiles (`read(Chemfiles.Trajectory(xyz_file))`, test/runtests.jl:20-21)
and writes nothing; a stepping drop-in needs the two obvious pieces: XYZ frames (read and append) and a restart file.
Everything here is plain numpy on host arrays in atom-id order -- what `NonbondedSystem.positions()` returns."""
import numpy as np


def read_xyz(path, frame=0):
    """One frame of an XYZ file: (names list, positions (N,3) float64, comment line).  The layout of the reference's
    test/data/lj_sample.xyz: atom count, comment (may be empty), then `name x y z` per atom."""
    with open(path) as fh:
        k = 0
        while True:
            head = fh.readline()
            if not head:
                raise IndexError("%s holds %d frame(s), frame %d requested" % (path, k, frame))
            if not head.strip():
                continue
            n = int(head.split()[0])
            comment = fh.readline().rstrip("\n")
            if k == frame:
                names, pos = [], np.empty((n, 3))
                for i in range(n):
                    t = fh.readline().split()
                    if len(t) < 4:
                        raise ValueError("%s: atom %d of frame %d is malformed" % (path, i + 1, frame))
                    names.append(t[0])
                    pos[i] = (float(t[1]), float(t[2]), float(t[3]))
                return names, pos, comment
            for _ in range(n):
                fh.readline()
            k += 1


def count_xyz_frames(path):
    n_frames = 0
    with open(path) as fh:
        while True:
            head = fh.readline()
            if not head:
                return n_frames
            if not head.strip():
                continue
            n = int(head.split()[0])
            fh.readline()
            for _ in range(n):
                fh.readline()
            n_frames += 1


class XYZWriter:
    """Appends frames to an XYZ trajectory: w = XYZWriter(path, names); w.write(system.positions(), "step 100")."""

    def __init__(self, path, names=None, mode="w", precision=12):
        self.path, self.names, self.fmt = path, names, "%%s %%.%dE %%.%dE %%.%dE\n" % (precision, precision, precision)
        self.frames = 0
        open(path, mode).close()

    def write(self, positions, comment=""):
        p = np.asarray(positions, dtype=np.float64)
        if p.ndim != 2 or p.shape[1] != 3:
            raise ValueError("positions must be (N,3)")
        names = self.names if self.names is not None else [str(i + 1) for i in range(p.shape[0])]
        if len(names) != p.shape[0]:
            raise ValueError("%d names for %d atoms" % (len(names), p.shape[0]))
        with open(self.path, "a") as fh:
            fh.write("%d\n%s\n" % (p.shape[0], str(comment).replace("\n", " ")))
            fh.writelines(self.fmt % (nm, x, y, z) for nm, (x, y, z) in zip(names, p))
        self.frames += 1


def save_checkpoint(path, ckpt, **extra):
    """Restart file (npz) from NonbondedSystem.checkpoint(): N, L, positions, velocities (+ caller's extras, e.g. step, time)."""
    np.savez(path, N=np.int64(ckpt["N"]), L=np.float64(ckpt["L"]), positions=np.asarray(ckpt["positions"], dtype=np.float64),
             velocities=np.asarray(ckpt["velocities"], dtype=np.float64), **extra)


def load_checkpoint(path):
    """The dict NonbondedSystem.restore() takes (bit-identical positions and velocities), extras included."""
    with np.load(path) as z:
        out = {k: z[k] for k in z.files}
    out["N"], out["L"] = int(out["N"]), float(out["L"])
    if out["positions"].shape != (out["N"], 3) or out["velocities"].shape != (out["N"], 3):
        raise ValueError("%s: arrays do not match N=%d" % (path, out["N"]))
    return out
