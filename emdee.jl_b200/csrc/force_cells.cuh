// force_cells.cuh -- EMDEE_CUTOFF mode: cell-list Lennard-Jones energy/force/virial in FP64.
//
// The B200 replacement for the reference's pair loop (compute_tile!, src/nonbonded.jl:44-107, which is
// all-pairs) built on the cell grid the reference only sketches (src/cells.jl:36,79-85,224-297).
//
// One thread block owns a HOME BRICK of bx*by*bz cells.  It stages the brick plus a halo of R cells
// (27-cell neighbourhoods for R=1, 125 for R=2) in shared memory: because atoms are sorted by
// (cell, id) with x fastest, every staged row of cells is at most two contiguous slot ranges, so
// the loads are coalesced streams.  Per staged atom: FP64 scaled position (exact geometry), LJ
// parameters, and an FP32 position relative to the brick origin (periodic image resolved per cell).
//
// Each thread owns one home atom i (full-neighbour scheme: no atomics, no reaction scatter; e_i and
// w_i get half of every pair like src/nonbonded.jl:93-94).  Work is split by pipe:
//   scan  (FP32/INT pipes): walk the (2R+1)^2 candidate rows, conservative FP32 distance test,
//                           append survivors to a per-lane queue in shared memory;
//   drain (FP64 pipe)     : exact minimum-image r2 in the oracle's rounding sequence, exact cull
//                           r2 <= rc2 (bit-exact pair set), interaction(), accumulate f, E, W.
// The FP64 pipe (64 lanes/clk/SM) therefore only sees pairs that are inside the cutoff (plus a
// 1e-3 margin), at full lane occupancy, instead of the 6.5x larger candidate set.
#pragma once
#include "lj_pair.cuh"

#define FC_QCAP 64          // per-lane queue entries (uint16 staged indices)
#define FC_MAX_HOMEROWS 64  // by*bz of the largest supported brick

struct CellArgs {
    GridDesc g;
    const int32_t *cell_start;   // local cells + 1
    const double *sx, *sy, *sz, *hs, *ts;
    const int32_t *id, *xbase;
    const uint64_t *xmask;
    double *fx, *fy, *fz, *en, *vir;
    double *partial;                  // per block: {sum e_i, sum w_i}
    unsigned long long *partial_n;    // per block: pairs with id_i < id_j
    unsigned long long *digest;       // AUDIT: {count, sum hash, xor hash}
    int32_t *pairs;                   // AUDIT: optional pair list (2 x pair_cap)
    long long pair_cap;
    unsigned long long *pair_n;
    double L;
    double cell_edge;                 // L / M
    LJModel model;
    float rc2f;                       // FP32 pre-cull threshold: rc2 * (1 + margin)
    int cap;                          // staged-atom capacity of the shared-memory arrays
    int ncs_max;                      // staged-cell capacity
    int *err;                         // device error flag (capacity overflow)
    int block_first;                  // first brick of this launch (launches may cover a z-layer range)
};

__host__ __device__ inline size_t fc_smem_bytes(int cap, int ncs_max, int block)
{
    size_t b = (size_t)cap * (5 * sizeof(double) + sizeof(float4));
    b += (size_t)(ncs_max + 1) * sizeof(int) * 2;    // cs[], gbase[]
    b += (FC_MAX_HOMEROWS + 1) * sizeof(int);        // hstart[]
    b += 16;                                         // scalars
    b = (b + 15) & ~(size_t)15;
    b += (size_t)FC_QCAP * block * sizeof(uint16_t);
    return b;
}

__device__ __forceinline__ int wrap_mod(int a, int M)
{
    a %= M;
    return a < 0 ? a + M : a;
}

template <int BLOCK, bool F, bool EW, bool EXCL, bool AUDIT>
__global__ void __launch_bounds__(BLOCK) k_force_cells(CellArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GridDesc &g = a.g;
    const int cap = a.cap;
    double *sxs = reinterpret_cast<double *>(smem_raw);
    double *sys = sxs + cap, *szs = sys + cap, *hss = szs + cap, *tss = hss + cap;
    float4 *prel = reinterpret_cast<float4 *>(tss + cap);
    int *cs = reinterpret_cast<int *>(prel + cap);
    int *gbase = cs + (a.ncs_max + 1);
    int *hstart = gbase + (a.ncs_max + 1);
    int *scal = hstart + (FC_MAX_HOMEROWS + 1);
    uint16_t *queue = reinterpret_cast<uint16_t *>(
        smem_raw + ((reinterpret_cast<unsigned char *>(scal + 4) - smem_raw + 15) & ~(size_t)15));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = BLOCK / 32;
    const int R = g.R, M = g.M;

    // ---- brick geometry ---------------------------------------------------------------------
    const int bid = blockIdx.x + a.block_first;
    int b = bid;
    const int bxi = b % g.nbx; b /= g.nbx;
    const int byi = b % g.nby;
    const int bzi = b / g.nby;
    const int hx0 = bxi * g.bx, hy0 = byi * g.by, hz0 = g.zhome0 + bzi * g.bz;   // first home cell (local z)
    const int nhx = min(g.bx, M - hx0), nhy = min(g.by, M - hy0), nhz = min(g.bz, g.zhome0 + g.nzhome - hz0);
    const int sxn = nhx + 2 * R, syn = nhy + 2 * R, szn = nhz + 2 * R;
    const int nrows = syn * szn, ncs = nrows * sxn;

    // ---- phase A: staged-cell table (count and first global slot of every staged cell) --------
    for (int t = tid; t < ncs; t += BLOCK) {
        const int cx = t % sxn, row = t / sxn, cy = row % syn, cz = row / syn;
        const int gx = wrap_mod(hx0 - R + cx, M), gy = wrap_mod(hy0 - R + cy, M);
        int lz = hz0 - R + cz;
        if (g.zwrap) lz = wrap_mod(lz, M);
        const int lc = gx + M * (gy + M * lz);
        const int s0 = a.cell_start[lc];
        gbase[t] = s0;
        cs[t] = a.cell_start[lc + 1] - s0;
    }
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the counts, 32 at a time
        int run = 0;
        for (int base = 0; base < ncs; base += 32) {
            const int t = base + lane;
            const int c = t < ncs ? cs[t] : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            if (t < ncs) cs[t] = run + inc - c;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) cs[ncs] = run;
        __syncwarp();
        if (lane == 0) {
            // home rows: the atoms of the home cells of row (hy,hz) are one contiguous staged range
            int h = 0;
            for (int hz = 0; hz < nhz; hz++)
                for (int hy = 0; hy < nhy; hy++) {
                    hstart[hz * nhy + hy] = h;
                    const int row = (hz + R) * syn + (hy + R);
                    h += cs[row * sxn + R + nhx] - cs[row * sxn + R];
                }
            hstart[nhy * nhz] = h;
            scal[0] = h;      // nhome
            scal[1] = run;    // nstaged
            if (run > cap) atomicExch(a.err, 2);
        }
    }
    __syncthreads();
    const int nhome = scal[0];
    const int nstaged = min(scal[1], cap);

    // ---- phase B: stage atoms row by row (each row = up to two contiguous slot ranges) --------
    for (int piece = warp; piece < 2 * nrows; piece += NW) {
        const int row = piece >> 1, second = piece & 1;
        const int cy = row % syn, cz = row / syn;
        const int ux0 = hx0 - R;                       // unwrapped x of staged cx = 0
        const int gx0 = wrap_mod(ux0, M);
        const int len1 = min(sxn, M - gx0);            // cells before the periodic wrap
        const int cfirst = second ? len1 : 0, clast = second ? sxn : len1;
        if (cfirst >= clast) continue;
        const int ibeg = cs[row * sxn + cfirst], iend = cs[row * sxn + clast];
        const int sbeg = gbase[row * sxn + cfirst];
        const double ccy = (double)(cy - R) + 0.5, ccz = (double)(cz - R) + 0.5;   // cell centre, brick-relative, cell units
        const double csy = ((double)(hy0 - R + cy) + 0.5) / M;                      // unwrapped scaled centre
        const int uz = (g.zwrap ? hz0 : g.zglob0 + hz0) - R + cz;
        const double csz = ((double)uz + 0.5) / M;
        for (int idx = ibeg + lane; idx < iend && idx < cap; idx += 32) {
            const int slot = sbeg + (idx - ibeg);
            int cx = cfirst;
            while (cx + 1 < clast && cs[row * sxn + cx + 1] <= idx) cx++;
            const double x = a.sx[slot], y = a.sy[slot], z = a.sz[slot];
            sxs[idx] = x; sys[idx] = y; szs[idx] = z;
            hss[idx] = a.hs[slot]; tss[idx] = a.ts[slot];
            const double csx = ((double)(ux0 + cx) + 0.5) / M;
            double dx = x - csx, dy = y - csy, dz = z - csz;
            dx -= rint(dx); dy -= rint(dy); dz -= rint(dz);      // image nearest to the staged cell
            float4 p;
            p.x = (float)(a.L * dx + ((double)(cx - R) + 0.5) * a.cell_edge);
            p.y = (float)(a.L * dy + ccy * a.cell_edge);
            p.z = (float)(a.L * dz + ccz * a.cell_edge);
            p.w = __int_as_float(slot);
            prel[idx] = p;
        }
    }
    __syncthreads();

    // ---- phases C+D: one home atom per lane, scan (FP32) / drain (FP64) -------------------------
    const double c60id2 = 60.0 * a.model.id2;
    const float rc2f = a.rc2f;
    double esum = 0, wsum = 0;
    unsigned long long npair = 0, hsum = 0, hxor = 0;

    for (int hbase = warp * 32; hbase < nhome; hbase += NW * 32) {
        const int h = hbase + lane;
        const bool active = h < nhome;
        int self = 0, cxi = R, cyi = R, czi = R;
        if (active) {
            int hr = 0;
            while (hstart[hr + 1] <= h) hr++;
            cyi = hr % nhy + R; czi = hr / nhy + R;
            const int row = czi * syn + cyi;
            self = cs[row * sxn + R] + (h - hstart[hr]);
            while (cs[row * sxn + cxi + 1] <= self) cxi++;
        }
        const float4 pi = prel[active ? self : 0];
        const double six = sxs[self], siy = sys[self], siz = szs[self], hsi = hss[self], tsi = tss[self];
        const int slot_i = __float_as_int(pi.w);
        int32_t idi = 0, xb = 0; uint64_t xm = 0;
        if (EXCL || AUDIT) idi = a.id[slot_i];
        if (EXCL) { xb = a.xbase[slot_i]; xm = a.xmask[slot_i]; }
        double fx = 0, fy = 0, fz = 0, e = 0, w = 0;

        // candidate iterator: (2R+1)^2 rows, each a contiguous staged range
        const int nwin = 2 * R + 1;
        int rw = active ? 0 : nwin * nwin;   // next row-window index
        int p = 0, pend = 0;
        for (;;) {
            int cnt = 0;
            // -------- scan --------
            while (cnt < FC_QCAP) {
                if (p == pend) {
                    if (rw == nwin * nwin) break;
                    const int dy = rw % nwin - R, dz = rw / nwin - R;
                    const int row = (czi + dz) * syn + (cyi + dy);
                    p = cs[row * sxn + cxi - R];
                    pend = min(cs[row * sxn + cxi + R + 1], nstaged);
                    rw++;
                    continue;
                }
                const float4 c = prel[p];
                const float dx = c.x - pi.x, dy = c.y - pi.y, dz = c.z - pi.z;
                const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (r2 <= rc2f && p != self) {
                    queue[cnt * BLOCK + tid] = (uint16_t)p;
                    cnt++;
                }
                p++;
            }
            // -------- drain --------
            const int maxcnt = __reduce_max_sync(0xffffffffu, cnt);
            if (maxcnt == 0) break;
            for (int q = 0; q < maxcnt; q++) {
                if (q < cnt) {
                    const int j = queue[q * BLOCK + tid];
                    double vx, vy, vz;
                    const double r2 = min_image_r2(six, siy, siz, sxs[j], sys[j], szs[j], a.L, vx, vy, vz);
                    bool ok = r2 <= a.model.rc2;
                    int32_t idj = 0;
                    if (EXCL || AUDIT) {
                        if (ok) idj = a.id[__float_as_int(prel[j].w)];
                        if (EXCL && ok && pair_excluded(xb, xm, idj)) ok = false;
                    }
                    if (ok) {
                        const double inv = rcp_fast(r2);
                        double Eg, Wg;
                        lj_interaction(r2, inv, hsi + hss[j], tsi * tss[j], a.model, c60id2, Eg, Wg);
                        if (F) {
                            const double qf = Wg * inv;
                            fx = fma(qf, vx, fx); fy = fma(qf, vy, fy); fz = fma(qf, vz, fz);
                        }
                        if (EW) { e += Eg; w += Wg; }
                        if (AUDIT && idi < idj) {
                            const uint64_t hh = pair_hash(idi, idj);
                            npair++; hsum += hh; hxor ^= hh;
                            if (a.pairs) {
                                const unsigned long long k = atomicAdd(a.pair_n, 1ull);
                                if ((long long)k < a.pair_cap) { a.pairs[2 * k] = idi; a.pairs[2 * k + 1] = idj; }
                            }
                        }
                    }
                }
            }
            if (__all_sync(0xffffffffu, rw == nwin * nwin && p == pend)) break;
        }
        if (active) {
            if (F) { a.fx[slot_i] = fx; a.fy[slot_i] = fy; a.fz[slot_i] = fz; }
            if (EW) {
                e *= 0.5; w *= 0.5;                         // src/nonbonded.jl:93-94
                a.en[slot_i] = e; a.vir[slot_i] = w;
                esum += e; wsum += w;
            }
        }
    }

    // ---- block totals in a fixed order (deterministic) ----------------------------------------
    __syncthreads();
    if (EW || AUDIT) {
        double *red = sxs;    // staged data no longer needed
        unsigned long long *redn = reinterpret_cast<unsigned long long *>(red + 2 * NW);
        esum = warp_sum_f64(esum); wsum = warp_sum_f64(wsum);
        for (int o = 16; o > 0; o >>= 1) {
            npair += __shfl_xor_sync(0xffffffffu, npair, o);
            hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
            hxor ^= __shfl_xor_sync(0xffffffffu, hxor, o);
        }
        if (lane == 0) { red[warp] = esum; red[NW + warp] = wsum; redn[warp] = npair; redn[NW + warp] = hsum; redn[2 * NW + warp] = hxor; }
        __syncthreads();
        if (tid == 0) {
            double E = 0, W = 0; unsigned long long n = 0, hs_ = 0, hx_ = 0;
            for (int k = 0; k < NW; k++) { E += red[k]; W += red[NW + k]; n += redn[k]; hs_ += redn[NW + k]; hx_ ^= redn[2 * NW + k]; }
            if (EW) { a.partial[2 * bid] = E; a.partial[2 * bid + 1] = W; }
            if (AUDIT) {
                a.partial_n[bid] = n;
                atomicAdd(a.digest, n); atomicAdd(a.digest + 1, hs_); atomicXor(a.digest + 2, hx_);
            }
        }
    }
}

// Sum the per-block partials in index order: deterministic totals.
__global__ void k_reduce_partials(int nblocks, const double *__restrict__ partial, double *__restrict__ totals)
{
    __shared__ double sE[256], sW[256];
    double E = 0, W = 0;
    // fixed assignment of blocks to threads, fixed tree afterwards
    for (int b = threadIdx.x; b < nblocks; b += 256) { E += partial[2 * b]; W += partial[2 * b + 1]; }
    sE[threadIdx.x] = E; sW[threadIdx.x] = W;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sE[threadIdx.x] += sE[threadIdx.x + o]; sW[threadIdx.x] += sW[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = sE[0]; totals[1] = sW[0]; }
}
