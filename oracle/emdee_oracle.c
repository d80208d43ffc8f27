/* emdee_oracle.c -- CPU restatement of EmDee.jl's nonbonded hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (emdee.jl_b200/, include/) may link,
 * import or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker or the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" against the reference's own binaries.  The reference is Julia
 * (not installed here), its tests hold ONE input fixture (test/data/lj_sample.xyz) and
 * ZERO stored numeric answers (test/runtests.jl:39-41 is a GPU-vs-CPU self-consistency check), so
 * the oracle is pinned by (1) following the cited reference lines operation for operation,
 * (2) bit-for-bit agreement in FP64 with an independent numpy restatement (oracle/oracle_np.py),
 * (3) its Float32 instantiation passing the reference's own <1e-4 criterion between the naive loop
 * and the tile-order loop, (4) analytic checks (FCC zero force, W = -r dE/dr, g(rs)=1, g(rc)->0).
 *
 * Citations are relative to /root/reference/.  Build: see oracle/Makefile (-ffp-contract=off, so the
 * only fused multiply-adds are the explicit fma() calls below).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "emdee_oracle.h"

/* ---- precision-generic pieces: Float64 (parity target) and Float32 (the reference's own type) ---- */
#define REAL double
#define SUF(x) x##_f64
#define RINT rint
#define FLOOR floor
#include "oracle_real.inc"
#undef REAL
#undef SUF
#undef RINT
#undef FLOOR

#define REAL float
#define SUF(x) x##_f32
#define RINT rintf
#define FLOOR floorf
#include "oracle_real.inc"
#undef REAL
#undef SUF
#undef RINT
#undef FLOOR

/* bench.py --impl reference under torchrun: the launcher exports OMP_NUM_THREADS=1 for every rank, but the reference
 * arm runs on rank 0 alone and must use all the host threads it can (n <= 0: leave the runtime's default alone). */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* nonbonded_computation_tiles(N) -- src/nonbonded.jl:18-26.  n = cld(N,32); for i=0:n-1, j=1:n-i
 * emit (j, j+i), 1-based.  Returns the tile count n(n+1)/2; out may be NULL to query the size. */
int64_t oracle_tiles(int64_t N, int32_t *out)
{
    int64_t n = (N + 31) / 32, k = 0;
    if (out)
        for (int64_t i = 0; i < n; i++)
            for (int64_t j = 1; j <= n - i; j++) { out[2 * k] = (int32_t)j; out[2 * k + 1] = (int32_t)(j + i); k++; }
    return n * (n + 1) / 2;
}

/* cells_per_dimension(L, cutoff, ndiv) = floor(Int32, ndiv*L/cutoff) -- src/cells.jl:36 */
int32_t oracle_cells_per_dimension(double L, double cutoff, int ndiv)
{
    return (int32_t)floor((double)ndiv * L / cutoff);
}

/* splitmix64 finaliser, the stateless generator of SURVEY section 8(d) and the pair-hash of the audit. */
static inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
uint64_t oracle_mix64(uint64_t z) { return mix64(z); }

static inline uint64_t pair_hash(int64_t i, int64_t j) /* i<j, 0-based global ids */
{
    return mix64(((uint64_t)i << 32) | (uint64_t)j);
}

/* Exclusion test: atom i carries a 64-bit mask over global ids [base_i, base_i+64). */
static inline int excluded(const int32_t *base, const uint64_t *mask, int64_t i, int64_t j)
{
    if (!base) return 0;
    int64_t o = j - (int64_t)base[i];
    return o >= 0 && o < 64 && ((mask[i] >> o) & 1ULL);
}

/* CUTOFF-mode distance: the minimum-image vector of src/nonbonded.jl:60-61,70 in FP64,
 *   s = r/L ; d = s_i - s_j ; d -= rint(d) ; v = L*d ; r2 = fma(vz,vz, fma(vy,vy, vx*vx)),
 * and the cull predicate of src/cells.jl:241,246,260 (`<=`, inclusive) restated on that r2. */
static inline double dist2(const double *si, const double *sj, double L, double v[3])
{
    for (int c = 0; c < 3; c++) {
        double d = si[c] - sj[c];
        d = d - rint(d);
        v[c] = L * d;
    }
    return fma(v[2], v[2], fma(v[1], v[1], v[0] * v[0]));
}

/* Brute-force pair set P = {(i,j): i<j, r2_ij <= rc2, not excluded}, lexicographic order.
 * Pair uniqueness i<j is src/cells.jl:246.  Returns |P|; writes at most cap pairs. */
int64_t oracle_pair_set_brute(int64_t N, const double *pos, double L, double rc2,
                              const int32_t *excl_base, const uint64_t *excl_mask,
                              int32_t *ij, int64_t cap, uint64_t digest[3])
{
    double *s = (double *)malloc(sizeof(double) * 3 * (size_t)N);
    for (int64_t k = 0; k < 3 * N; k++) s[k] = pos[k] / L;
    int64_t n = 0;
    uint64_t sum = 0, xr = 0;
    for (int64_t i = 0; i < N; i++)
        for (int64_t j = i + 1; j < N; j++) {
            double v[3];
            if (dist2(s + 3 * i, s + 3 * j, L, v) <= rc2 && !excluded(excl_base, excl_mask, i, j)) {
                if (ij && n < cap) { ij[2 * n] = (int32_t)i; ij[2 * n + 1] = (int32_t)j; }
                uint64_t h = pair_hash(i, j);
                sum += h; xr ^= h; n++;
            }
        }
    if (digest) { digest[0] = (uint64_t)n; digest[1] = sum; digest[2] = xr; }
    free(s);
    return n;
}

/* ---- cell list used by the O(N) oracle: counting sort by the reference's cell index ---- */
typedef struct {
    int32_t M;
    int64_t ncell;
    int64_t *start;   /* ncell+1 */
    int32_t *order;   /* atoms sorted by (cell, id) */
    int32_t *cell;    /* 0-based cell of each atom */
} cell_list;

static void cell_list_build(cell_list *cl, int64_t N, const double *pos, double L, int32_t M)
{
    cl->M = M;
    cl->ncell = (int64_t)M * M * M;
    cl->start = (int64_t *)calloc((size_t)cl->ncell + 1, sizeof(int64_t));
    cl->order = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
    cl->cell = (int32_t *)malloc(sizeof(int32_t) * (size_t)N);
    oracle_cell_index_f64(N, pos, L, M, cl->cell);
    for (int64_t i = 0; i < N; i++) { cl->cell[i] -= 1; cl->start[cl->cell[i] + 1]++; }
    for (int64_t c = 0; c < cl->ncell; c++) cl->start[c + 1] += cl->start[c];
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)cl->ncell);
    memcpy(fill, cl->start, sizeof(int64_t) * (size_t)cl->ncell);
    for (int64_t i = 0; i < N; i++) cl->order[fill[cl->cell[i]]++] = (int32_t)i;   /* stable: ascending id */
    free(fill);
}
static void cell_list_free(cell_list *cl) { free(cl->start); free(cl->order); free(cl->cell); }

/* Distinct neighbour cells of cell (cx,cy,cz) within `reach` cells, periodic, duplicates removed
 * (the reference's pbc() wraps one image only and double-counts for small M). */
static int neighbour_cells(int32_t M, int reach, int cx, int cy, int cz, int64_t *out)
{
    int n = 0;
    for (int dz = -reach; dz <= reach; dz++)
        for (int dy = -reach; dy <= reach; dy++)
            for (int dx = -reach; dx <= reach; dx++) {
                int x = ((cx + dx) % M + M) % M, y = ((cy + dy) % M + M) % M, z = ((cz + dz) % M + M) % M;
                int64_t c = x + ((int64_t)y + (int64_t)z * M) * M;
                int dup = 0;
                for (int k = 0; k < n; k++) if (out[k] == c) { dup = 1; break; }
                if (!dup) out[n++] = c;
            }
    return n;
}

/* CUTOFF-mode energy/force/virial through a cell list (O(N)), full-neighbour so that every atom's
 * sums are formed in a fixed order whatever the thread count:
 *   pair set  : r2 <= rc2 (dist2 above), minus exclusions
 *   pair math : interaction() verbatim on that set (src/lennard_jones.jl:25-42)
 *   force     : f_ij = (W/r2) * rv, f_i += f_ij           (src/nonbonded.jl:74-75)
 *   per atom  : e_i = 0.5*sum_j E, w_i = 0.5*sum_j W       (src/nonbonded.jl:93-94)
 * ndiv in {1,2,...}: M = floor(ndiv*L/rc) (src/cells.jl:36), neighbour reach = ndiv cells.
 * totals[0]=sum e_i, totals[1]=sum w_i (id order); *npairs = |P|; digest = (|P|, sum hash, xor hash). */
int oracle_cutoff_cells(int64_t N, const double *pos, double L, double cutoff, double sw, const double *atoms,
                        int ndiv, const int32_t *excl_base, const uint64_t *excl_mask, int bitmask,
                        double *forces, double *energies, double *virials,
                        double totals[2], int64_t *npairs, uint64_t digest[3])
{
    double model[3];
    oracle_lj_model_f64(cutoff, sw, model);
    const double rc2 = model[0];
    int32_t M = oracle_cells_per_dimension(L, cutoff, ndiv);
    if (M < 1) return 1;
    int reach = ndiv;
    if (2 * reach + 1 > M) reach = M / 2;          /* all cells are neighbours; dedupe handles it */
    if (reach < 1 && M > 1) reach = 1;
    cell_list cl;
    cell_list_build(&cl, N, pos, L, M);
    double *s = (double *)malloc(sizeof(double) * 3 * (size_t)N);
    for (int64_t k = 0; k < 3 * N; k++) s[k] = pos[k] / L;
    int nmax = (2 * reach + 1) * (2 * reach + 1) * (2 * reach + 1);
    int64_t np = 0;
    uint64_t hsum = 0, hxor = 0;
    double *ei = (double *)malloc(sizeof(double) * (size_t)N), *wi = (double *)malloc(sizeof(double) * (size_t)N);

#pragma omp parallel reduction(+ : np, hsum) reduction(^ : hxor)
    {
        int64_t *nb = (int64_t *)malloc(sizeof(int64_t) * (size_t)nmax);
#pragma omp for schedule(dynamic, 8)
        for (int64_t c = 0; c < cl.ncell; c++) {
            if (cl.start[c] == cl.start[c + 1]) continue;
            int cx = (int)(c % M), cy = (int)((c / M) % M), cz = (int)(c / ((int64_t)M * M));
            int nn = neighbour_cells(M, reach, cx, cy, cz, nb);
            for (int64_t a = cl.start[c]; a < cl.start[c + 1]; a++) {
                int64_t i = cl.order[a];
                double fx = 0, fy = 0, fz = 0, e = 0, w = 0;
                for (int k = 0; k < nn; k++)
                    for (int64_t b = cl.start[nb[k]]; b < cl.start[nb[k] + 1]; b++) {
                        int64_t j = cl.order[b];
                        if (j == i) continue;
                        double v[3];
                        double r2 = dist2(s + 3 * i, s + 3 * j, L, v);
                        if (!(r2 <= rc2) || excluded(excl_base, excl_mask, i, j)) continue;
                        double E, W;
                        interaction_f64(r2, model, atoms[2 * i], atoms[2 * i + 1], atoms[2 * j], atoms[2 * j + 1], &E, &W);
                        double q = W / r2;
                        fx += q * v[0]; fy += q * v[1]; fz += q * v[2];
                        e += E; w += W;
                        if (i < j) { uint64_t h = pair_hash(i, j); np++; hsum += h; hxor ^= h; }
                    }
                if (bitmask & 1) { forces[3 * i] = fx; forces[3 * i + 1] = fy; forces[3 * i + 2] = fz; }
                ei[i] = 0.5 * e; wi[i] = 0.5 * w;
            }
        }
        free(nb);
    }
    double Et = 0, Wt = 0;
    for (int64_t i = 0; i < N; i++) { Et += ei[i]; Wt += wi[i]; }
    if (bitmask & 2) memcpy(energies, ei, sizeof(double) * (size_t)N);
    if (bitmask & 4) memcpy(virials, wi, sizeof(double) * (size_t)N);
    if (totals) { totals[0] = Et; totals[1] = Wt; }
    if (npairs) *npairs = np;
    if (digest) { digest[0] = (uint64_t)np; digest[1] = hsum; digest[2] = hxor; }
    free(ei); free(wi); free(s);
    cell_list_free(&cl);
    return 0;
}

/* Sorted pair set through the cell list (for N too large for the brute-force loop).
 * Output is sorted lexicographically by (i,j), i<j, identical to oracle_pair_set_brute. */
static int cmp_pair(const void *a, const void *b)
{
    const int32_t *p = (const int32_t *)a, *q = (const int32_t *)b;
    if (p[0] != q[0]) return p[0] < q[0] ? -1 : 1;
    return (p[1] > q[1]) - (p[1] < q[1]);
}
int64_t oracle_pair_set_cells(int64_t N, const double *pos, double L, double cutoff, int ndiv,
                              const int32_t *excl_base, const uint64_t *excl_mask, int32_t *ij, int64_t cap)
{
    const double rc2 = cutoff * cutoff;
    int32_t M = oracle_cells_per_dimension(L, cutoff, ndiv);
    int reach = ndiv;
    if (2 * reach + 1 > M) reach = M / 2;
    if (reach < 1 && M > 1) reach = 1;
    cell_list cl;
    cell_list_build(&cl, N, pos, L, M);
    double *s = (double *)malloc(sizeof(double) * 3 * (size_t)N);
    for (int64_t k = 0; k < 3 * N; k++) s[k] = pos[k] / L;
    int nmax = (2 * reach + 1) * (2 * reach + 1) * (2 * reach + 1);
    int64_t *nb = (int64_t *)malloc(sizeof(int64_t) * (size_t)nmax);
    int64_t n = 0;
    for (int64_t c = 0; c < cl.ncell; c++) {
        int cx = (int)(c % M), cy = (int)((c / M) % M), cz = (int)(c / ((int64_t)M * M));
        int nn = neighbour_cells(M, reach, cx, cy, cz, nb);
        for (int64_t a = cl.start[c]; a < cl.start[c + 1]; a++) {
            int64_t i = cl.order[a];
            for (int k = 0; k < nn; k++)
                for (int64_t b = cl.start[nb[k]]; b < cl.start[nb[k] + 1]; b++) {
                    int64_t j = cl.order[b];
                    if (j <= i) continue;
                    double v[3];
                    if (dist2(s + 3 * i, s + 3 * j, L, v) <= rc2 && !excluded(excl_base, excl_mask, i, j)) {
                        if (ij && n < cap) { ij[2 * n] = (int32_t)i; ij[2 * n + 1] = (int32_t)j; }
                        n++;
                    }
                }
        }
    }
    if (ij) qsort(ij, (size_t)(n < cap ? n : cap), 2 * sizeof(int32_t), cmp_pair);
    free(nb); free(s);
    cell_list_free(&cl);
    return n;
}

/* 1-4 scaling (lj14scale: parsed at src/modelling.jl:199 from the force-field file, test/data/dibenzo-p-dioxin-in-water.xml:84,
 * and never applied by the reference -- "parity unpinned", this definition IS the pin): a pair of atoms three bonds apart
 * interacts with lj14scale times the Lennard-Jones energy / virial / force of an ordinary pair.  Stated as a correction to an
 * evaluation that treated those pairs at full strength: every listed pair (i<j) inside the cutoff (same pair-set predicate,
 * dist2 above) adds (scale-1)*(E, W) -- half to each atom, like every pair (src/nonbonded.jl:93-94) -- and (scale-1)*f_ij.
 * Serial, in list order.  Returns the number of listed pairs inside the cutoff; dtot[0..1] += correction of sum E, sum W. */
int64_t oracle_pairs14_correction(int64_t N, const double *pos, double L, double cutoff, double sw, const double *atoms,
                                  const int32_t *ij, int64_t n14, double scale, int bitmask,
                                  double *forces, double *energies, double *virials, double dtot[2])
{
    double model[3];
    oracle_lj_model_f64(cutoff, sw, model);
    const double rc2 = model[0], c = scale - 1.0;
    int64_t inside = 0;
    (void)N;
    for (int64_t k = 0; k < n14; k++) {
        const int64_t i = ij[2 * k], j = ij[2 * k + 1];
        double si[3], sj[3], v[3];
        for (int d = 0; d < 3; d++) { si[d] = pos[3 * i + d] / L; sj[d] = pos[3 * j + d] / L; }
        const double r2 = dist2(si, sj, L, v);
        if (!(r2 <= rc2)) continue;
        inside++;
        double E, W;
        interaction_f64(r2, model, atoms[2 * i], atoms[2 * i + 1], atoms[2 * j], atoms[2 * j + 1], &E, &W);
        const double q = c * (W / r2);
        if (bitmask & 1)
            for (int d = 0; d < 3; d++) { forces[3 * i + d] += q * v[d]; forces[3 * j + d] -= q * v[d]; }
        if (bitmask & 2) { energies[i] += 0.5 * c * E; energies[j] += 0.5 * c * E; }
        if (bitmask & 4) { virials[i] += 0.5 * c * W; virials[j] += 0.5 * c * W; }
        if (dtot) { dtot[0] += c * E; dtot[1] += c * W; }
    }
    return inside;
}

/* Velocity-Verlet (the reference has no integrator, F6 -- "parity unpinned", this
 * definition IS the pin):  v += (dt/2m) f ; r += dt v (no wrapping) ; f = F(r) ; v += (dt/2m) f,
 * written with explicit fma so that CPU and GPU round identically per step.
 * forces must hold F(r) on entry (call oracle_cutoff_cells first) and holds F(r) on exit. */
int oracle_vv_steps(int64_t N, double *pos, double *vel, double *forces, const double *mass, double L,
                    double cutoff, double sw, const double *atoms, int ndiv,
                    const int32_t *excl_base, const uint64_t *excl_mask, double dt, int64_t nsteps)
{
    for (int64_t st = 0; st < nsteps; st++) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; i++) {
            double h = 0.5 * dt / mass[i];
            for (int c = 0; c < 3; c++) {
                vel[3 * i + c] = fma(h, forces[3 * i + c], vel[3 * i + c]);
                pos[3 * i + c] = fma(dt, vel[3 * i + c], pos[3 * i + c]);
            }
        }
        int rc = oracle_cutoff_cells(N, pos, L, cutoff, sw, atoms, ndiv, excl_base, excl_mask, 1,
                                     forces, NULL, NULL, NULL, NULL, NULL);
        if (rc) return rc;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; i++) {
            double h = 0.5 * dt / mass[i];
            for (int c = 0; c < 3; c++) vel[3 * i + c] = fma(h, forces[3 * i + c], vel[3 * i + c]);
        }
    }
    return 0;
}
