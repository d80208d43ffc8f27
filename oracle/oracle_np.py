"""oracle_np.py -- numpy twin of the CPU oracle (TEST INFRASTRUCTURE ONLY).

An independent restatement of the same reference lines as oracle/emdee_oracle.c, written with
vectorised numpy instead of C loops, used to pin the C oracle (the reference itself is Julia and
cannot run here).  numpy elementwise arithmetic is plain IEEE (no
contraction), so in FP64 the two restatements must agree to the last bit wherever the summation
order is the same; where the C oracle calls fma() the twin evaluates the fused operation exactly
with rational arithmetic (fractions.Fraction -> float is correctly rounded).

Citations are relative to /root/reference/.  PARITY STATUS: "parity unpinned" against reference
binaries (see emdee_oracle.c header).
"""
from fractions import Fraction

import numpy as np

FORCES, ENERGIES, VIRIALS = 1, 2, 4  # src/nonbonded.jl:12-14


def lj_model(cutoff, switch, dtype=np.float64):
    """LennardJonesModel(cutoff, switch) -- src/lennard_jones.jl:6-11 (rc2, rs2, 1/(rc2-rs2))."""
    cutoff, switch = float(cutoff), float(switch)
    return np.array([cutoff * cutoff, switch * switch, 1.0 / (cutoff * cutoff - switch * switch)]).astype(dtype)


def lj_atom(eps, sigma, dtype=np.float64):
    """LennardJonesAtom(eps, sigma) = LJAtom(0.5 sigma, 2 sqrt(eps)) -- src/lennard_jones.jl:13."""
    return np.array([0.5 * float(sigma), 2.0 * np.sqrt(float(eps))]).astype(dtype)


def interaction(r2, model, hs_i, ts_i, hs_j, ts_j):
    """interaction() -- src/lennard_jones.jl:25-42, vectorised over pairs; dtype follows r2."""
    T = r2.dtype.type
    rs2, id2 = T(model[1]), T(model[2])
    sigma = hs_i + hs_j                                  # :29
    s2 = sigma * sigma / r2                              # :31
    s6 = s2 * s2 * s2                                    # :32
    e4s6 = ts_i * ts_j * s6                              # :33
    E = e4s6 * (s6 - T(1))                               # :34
    mEr = (T(6) * e4s6) * (T(2) * s6 - T(1))             # :35
    x = (r2 - rs2) * id2                                 # :36
    x = x * (T(0.5) * (np.sign(x) - np.sign(x - T(1))))  # :37
    x2 = x * x                                           # :38
    g = T(1) + (x * x2) * ((T(15) * x - T(6) * x2) - T(10))           # :39
    mgr = ((T(60) * x2) * ((T(1) - T(2) * x) + x2)) * id2 * r2        # :40
    return E * g, mEr * g + E * mgr                      # :41


def _min_image_vectors(pos, L, i, j):
    """rv = L*minimum_image.(s_i - s_j) -- src/nonbonded.jl:40,60-61,70; np.rint is ties-to-even."""
    T = pos.dtype.type
    s = pos / T(L)
    d = s[i] - s[j]
    d = d - np.rint(d)
    return T(L) * d


def allpairs(pos, L, model, atoms):
    """ALLPAIRS_REFERENCE: every i<j minimum-image pair through interaction() (F4: no cull).
    pos is (N,3); returns per-pair arrays (i, j, rv, r2, E, W) with r2 = (x*x + y*y) + z*z
    (src/nonbonded.jl:42,71)."""
    N = pos.shape[0]
    i, j = np.triu_indices(N, k=1)
    rv = _min_image_vectors(pos, L, i, j)
    r2 = (rv[:, 0] * rv[:, 0] + rv[:, 1] * rv[:, 1]) + rv[:, 2] * rv[:, 2]
    E, W = interaction(r2, model, atoms[i, 0], atoms[i, 1], atoms[j, 0], atoms[j, 1])
    return i, j, rv, r2, E, W


def naive_allpairs_f64(pos, L, model, atoms):
    """naively_compute_nonbonded! -- src/nonbonded.jl:122-155, all Float64.  The accumulation order
    is the reference's: atom i receives first the j-side contributions of rows i'<i (ascending i'),
    then the sum of its own row (ascending j) in one addition (:147-149)."""
    N = pos.shape[0]
    i, j, rv, r2, E, W = allpairs(pos, L, model, atoms)
    fij = (W / r2)[:, None] * rv
    f = np.zeros((N, 3)); e = np.zeros(N); w = np.zeros(N)
    # pairs from triu_indices are ordered by (i, j): rows are contiguous
    row_start = np.concatenate(([0], np.cumsum(np.arange(N - 1, 0, -1))))
    for a in range(N - 1):
        sl = slice(row_start[a], row_start[a + 1])
        fi = np.zeros(3); ei = 0.0; wi = 0.0
        for k in range(sl.start, sl.stop):          # sequential, to keep the reference's order
            fi += fij[k]; ei += E[k] / 2; wi += W[k] / 2
        jj = j[sl]
        f[jj] -= fij[sl]; e[jj] += E[sl] / 2; w[jj] += W[sl] / 2
        f[a] += fi; e[a] += ei; w[a] += wi
    return f, e, w


def cells_per_dimension(L, cutoff, ndiv):
    """floor(Int32, ndiv*L/cutoff) -- src/cells.jl:36."""
    return int(np.floor(ndiv * float(L) / float(cutoff)))


def cell_index(pos, L, M):
    """1-based cell index -- src/cells.jl:79-85,181, with the Q7 clamp v=min(v,M-1)."""
    T = pos.dtype.type
    s = pos / T(L)
    v = np.floor(T(M) * (s - np.floor(s))).astype(np.int32)
    v = np.minimum(v, M - 1)
    return (1 + v[:, 0] + (v[:, 1] + v[:, 2] * M) * M).astype(np.int32)


def _fma_exact(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def cutoff_r2(rv):
    """r2 = fma(vz,vz, fma(vy,vy, vx*vx)) evaluated exactly."""
    out = np.empty(rv.shape[0])
    for k in range(rv.shape[0]):
        x, y, z = (float(t) for t in rv[k])
        out[k] = _fma_exact(z, z, _fma_exact(y, y, x * x))
    return out


def cutoff_pairs(pos, L, rc2, excl_base=None, excl_mask=None):
    """CUTOFF pair set {i<j : r2 <= rc2} minus exclusions (src/cells.jl:241,246,260 restated on r2).
    Returns (i, j, rv, r2) sorted lexicographically.  Brute force; N up to a few thousand."""
    N = pos.shape[0]
    i, j = np.triu_indices(N, k=1)
    rv = _min_image_vectors(pos, L, i, j)
    r2_plain = (rv[:, 0] * rv[:, 0] + rv[:, 1] * rv[:, 1]) + rv[:, 2] * rv[:, 2]
    cand = np.nonzero(r2_plain <= rc2 * (1 + 1e-9))[0]      # superset; exact test below
    r2 = cutoff_r2(rv[cand])
    keep = r2 <= rc2
    cand, r2 = cand[keep], r2[keep]
    i, j, rv = i[cand], j[cand], rv[cand]
    if excl_base is not None:
        off = j.astype(np.int64) - excl_base[i].astype(np.int64)
        inwin = (off >= 0) & (off < 64)
        bit = (excl_mask[i] >> np.where(inwin, off, 0).astype(np.uint64)) & np.uint64(1)
        ok = ~(inwin & (bit == 1))
        i, j, rv, r2 = i[ok], j[ok], rv[ok], r2[ok]
    return i, j, rv, r2


def cutoff_compute(pos, L, cutoff, switch, atoms, excl_base=None, excl_mask=None):
    """CUTOFF-mode forces / per-atom energies / virials (np.add.at accumulation; the summation
    order differs from the C oracle's, so comparisons are to ~1e-13, not bitwise)."""
    model = lj_model(cutoff, switch)
    N = pos.shape[0]
    i, j, rv, r2 = cutoff_pairs(pos, L, model[0], excl_base, excl_mask)
    E, W = interaction(r2, model, atoms[i, 0], atoms[i, 1], atoms[j, 0], atoms[j, 1])
    fij = (W / r2)[:, None] * rv
    f = np.zeros((N, 3)); e = np.zeros(N); w = np.zeros(N)
    np.add.at(f, i, fij); np.add.at(f, j, -fij)
    np.add.at(e, i, E / 2); np.add.at(e, j, E / 2)
    np.add.at(w, i, W / 2); np.add.at(w, j, W / 2)
    return f, e, w, np.stack([i, j], axis=1).astype(np.int32)


def pairs14_correction(pos, L, cutoff, switch, atoms, ij, scale):
    """numpy twin of oracle_pairs14_correction (lj14scale, src/modelling.jl:199; never applied by the reference): listed pairs
    inside the cutoff add (scale-1) x (E, W, f_ij).  Returns (df (N,3), de (N,), dw (N,), n_inside)."""
    model = lj_model(cutoff, switch)
    N = pos.shape[0]
    ij = np.asarray(ij, dtype=np.int64).reshape(-1, 2)
    i, j = ij[:, 0], ij[:, 1]
    rv = _min_image_vectors(pos, L, i, j)
    r2 = cutoff_r2(rv)
    keep = r2 <= model[0]
    i, j, rv, r2 = i[keep], j[keep], rv[keep], r2[keep]
    E, W = interaction(r2, model, atoms[i, 0], atoms[i, 1], atoms[j, 0], atoms[j, 1])
    c = scale - 1.0
    fij = (c * (W / r2))[:, None] * rv
    f = np.zeros((N, 3)); e = np.zeros(N); w = np.zeros(N)
    np.add.at(f, i, fij); np.add.at(f, j, -fij)
    np.add.at(e, i, 0.5 * c * E); np.add.at(e, j, 0.5 * c * E)
    np.add.at(w, i, 0.5 * c * W); np.add.at(w, j, 0.5 * c * W)
    return f, e, w, int(keep.sum())


def mix64(z):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def pair_digest(ij):
    """(count, sum of hashes mod 2^64, xor of hashes) of a pair list with i<j, 0-based ids."""
    ij = np.asarray(ij, dtype=np.uint64).reshape(-1, 2)
    h = mix64((ij[:, 0] << np.uint64(32)) | ij[:, 1])
    with np.errstate(over="ignore"):
        s = np.add.reduce(h, dtype=np.uint64) if h.size else np.uint64(0)
    x = np.bitwise_xor.reduce(h) if h.size else np.uint64(0)
    return np.array([h.size, s, x], dtype=np.uint64)


def stencil_vectors_loops(rc, action):
    """src/cells.jl:22-34 as plain loops (test infrastructure: checks the vectorised host mirror in api.py)."""
    import math
    nmax = math.ceil(rc)
    M = 1 + 2 * nmax
    lo, hi = (1, M ** 3 // 2) if action else (M ** 3 // 2 + 2, M ** 3)
    out = []
    for index in range(lo, hi + 1):
        k, l = divmod(index - 1, M * M)
        j, i = divmod(l, M)
        v = (i - nmax, j - nmax, k - nmax)
        if sum((abs(c) - 1) ** 2 for c in v) < rc * rc:
            out.append(v)
    return out


def surrounding_cells_loops(L, cutoff, M, action):
    """src/cells.jl:38-44 as plain loops: table[v][cell-1] = 1-based neighbour cell."""
    vectors = stencil_vectors_loops(M * cutoff / L, action)

    def pbc(x):
        return x + M if x < 0 else (x - M if x >= M else x)

    table = []
    for v in vectors:
        row = []
        for index in range(1, M ** 3 + 1):
            k, l = divmod(index - 1, M * M)
            j, i = divmod(l, M)
            row.append(1 + pbc(i + v[0]) + M * pbc(j + v[1]) + M * M * pbc(k + v[2]))
        table.append(row)
    return table


def naive_allpairs_loops(pos, L, model, atoms):
    """naively_compute_nonbonded! (src/nonbonded.jl:122-155) as a literal scalar loop with the reference's mixed
    precision: pair math in pos.dtype (Float32 in the reference), i-side accumulators in Float64 (:132-134), j-side
    accumulated straight into the arrays (:141,143,145), i-side folded in after the j loop (:147-149).  Pure-Python
    loops -- small N only; third, independent restatement that the C oracle's f32 and f64 instantiations must match
    bit for bit."""
    T = pos.dtype.type
    N = pos.shape[0]
    Lt = T(L)
    sp = pos / Lt                                                         # :124
    f = np.zeros((N, 3), dtype=T); e = np.zeros(N, dtype=T); w = np.zeros(N, dtype=T)
    half = T(2)
    for i in range(N - 1):
        ei, wi, fi = np.float64(0), np.float64(0), np.zeros(3)             # :132-134
        for j in range(i + 1, N):
            d = sp[i] - sp[j]
            rv = Lt * (d - np.rint(d))                                    # :136
            r2 = (rv[0] * rv[0] + rv[1] * rv[1]) + rv[2] * rv[2]           # :137  sum(x.*y) is a left fold
            E, W = interaction(np.array([r2], dtype=T), model, atoms[i, 0], atoms[i, 1], atoms[j, 0], atoms[j, 1])
            E, W = E[0], W[0]
            fij = (W / r2) * rv                                           # :139
            fi += fij.astype(np.float64)                                  # :140
            f[j] -= fij                                                   # :141
            ei += np.float64(E / half); e[j] += E / half                  # :142-143
            wi += np.float64(W / half); w[j] += W / half                  # :144-145
        f[i] = (f[i].astype(np.float64) + fi).astype(T)                   # :147
        e[i] = T(np.float64(e[i]) + ei)                                   # :148
        w[i] = T(np.float64(w[i]) + wi)                                   # :149
    return f, e, w


def tiles_allpairs_loops(pos, L, tiles, model, atoms, bitmask=7):
    """compute_tile! + compute_nonbonded! (src/nonbonded.jl:44-120) lane by lane in plain loops: 1-based lane ids,
    partner lane j = (tid+m-1)%32+1 (:68), returning lane k = (tid+32-m-1)%32+1 (:69), shuffles modelled by reading
    the other lane's value of the same iteration, atomics applied in tile order / lane order, I side before J side.
    Lanes beyond N are masked (the reference would read out of bounds).  Small N only."""
    T = pos.dtype.type
    N = pos.shape[0]
    Lt = T(L)
    f = np.zeros((N, 3), dtype=T); e = np.zeros(N, dtype=T); w = np.zeros(N, dtype=T)
    for bI, bJ in np.asarray(tiles).reshape(-1, 2).tolist():
        diag = bI == bJ
        I = [(bI - 1) * 32 + tid for tid in range(1, 33)]                # 1-based atom ids, lane tid at index tid-1
        J = [(bJ - 1) * 32 + tid for tid in range(1, 33)]
        zero3 = np.zeros(3, dtype=T)
        si = [pos[a - 1] / Lt if a <= N else zero3 for a in I]           # :60
        sj = [pos[a - 1] / Lt if a <= N else zero3 for a in J]           # :61
        fi = [zero3.copy() for _ in range(32)]; fj = [zero3.copy() for _ in range(32)]
        ei = [T(0)] * 32; ej = [T(0)] * 32; wi = [T(0)] * 32; wj = [T(0)] * 32
        for m in range(1, 32 - int(diag) + 1):                           # :67
            fij, Eij, Wij = [zero3] * 32, [T(0)] * 32, [T(0)] * 32
            for tid in range(1, 33):
                j = (tid + m - 1) % 32 + 1                               # :68
                if I[tid - 1] > N or J[j - 1] > N:
                    continue
                d = si[tid - 1] - sj[j - 1]
                rv = Lt * (d - np.rint(d))                               # :70
                r2 = (rv[0] * rv[0] + rv[1] * rv[1]) + rv[2] * rv[2]      # :71
                ai, aj = atoms[I[tid - 1] - 1], atoms[J[j - 1] - 1]
                E, W = interaction(np.array([r2], dtype=T), model, ai[0], ai[1], aj[0], aj[1])   # :72
                Eij[tid - 1], Wij[tid - 1] = E[0], W[0]
                fij[tid - 1] = (W[0] / r2) * rv                          # :74
            for tid in range(1, 33):
                k = (tid + 32 - m - 1) % 32 + 1                          # :69
                fi[tid - 1] = fi[tid - 1] + fij[tid - 1]                 # :75
                fj[tid - 1] = fj[tid - 1] - fij[k - 1]                   # :76
                ei[tid - 1] = ei[tid - 1] + Eij[tid - 1]; ej[tid - 1] = ej[tid - 1] + Eij[k - 1]   # :79-80
                wi[tid - 1] = wi[tid - 1] + Wij[tid - 1]; wj[tid - 1] = wj[tid - 1] + Wij[k - 1]   # :83-84
        for tid in range(1, 33):                                         # :88-94
            a = I[tid - 1]
            if a > N:
                continue
            if bitmask & 1: f[a - 1] += fi[tid - 1]
            if bitmask & 2: e[a - 1] += T(0.5) * ei[tid - 1]
            if bitmask & 4: w[a - 1] += T(0.5) * wi[tid - 1]
        if not diag:                                                     # :96-104
            for tid in range(1, 33):
                a = J[tid - 1]
                if a > N:
                    continue
                if bitmask & 1: f[a - 1] += fj[tid - 1]
                if bitmask & 2: e[a - 1] += T(0.5) * ej[tid - 1]
                if bitmask & 4: w[a - 1] += T(0.5) * wj[tid - 1]
    return f, e, w


def vv_steps_twin(pos, vel, forces, mass, L, cutoff, switch, atoms, dt, nsteps, excl_base=None, excl_mask=None):
    """Velocity-Verlet as the oracle defines it (v += dt/2m f; r += dt v; f = F(r); v += dt/2m f) on the numpy twin's
    force evaluation.  No fused multiply-adds here, so agreement with the C oracle is to rounding, not bitwise."""
    pos, vel, forces = pos.copy(), vel.copy(), forces.copy()
    h = (0.5 * dt / mass)[:, None]
    for _ in range(nsteps):
        vel = vel + h * forces
        pos = pos + dt * vel
        forces = cutoff_compute(pos, L, cutoff, switch, atoms, excl_base, excl_mask)[0]
        vel = vel + h * forces
    return pos, vel, forces
