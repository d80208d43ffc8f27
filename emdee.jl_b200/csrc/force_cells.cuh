// force_cells.cuh -- EMDEE_CUTOFF mode: cell-list Lennard-Jones energy/force/virial in FP64.
//
// The B200 replacement for the reference's pair loop (compute_tile!, src/nonbonded.jl:44-107, which is
// all-pairs) built on the cell grid the reference only sketches (src/cells.jl:36,79-85,224-297).
//
// One thread block owns a HOME BRICK of bx*by*bz cells.  It stages the brick plus a halo of R cells
// (27-cell neighbourhoods for R=1, 125 for R=2) in shared memory: because atoms are sorted by
// (cell, id) with x fastest, every staged row of cells is at most two contiguous slot ranges, so
// the loads are coalesced streams.  Per staged atom: two 16-byte FP64 records {s_x, s_y} and
// {s_z, id|type} (exact geometry, two LDS.128 per neighbour) and a 16-byte FP32 record {x, y, z, |c|^2} relative to
// the brick centre with the periodic image resolved per cell.
//
// Work unit = WARP TASK: 32 consecutive home atoms of one home row (cells contiguous in x), one atom
// per lane, handed out dynamically (no trailing block barrier).  Full-neighbour scheme: every lane
// accumulates its own atom, no atomics, no reaction scatter; e_i and w_i get half of every pair like
// src/nonbonded.jl:93-94.  Each task alternates two phases that use different pipes:
//   scan  (FP32/INT pipes): all lanes walk the SAME candidate window -- the (2R+1)^2 staged rows
//          around the task's row, cells [first-R, last+R] -- so every LDS.128 is a one-address
//          broadcast and control flow is warp-uniform; a conservative FP32 distance test
//          (|c|^2 - 2 c.p + |p|^2: one add and three FMAs per candidate) pushes survivors on a
//          per-lane stack of staged indices in shared memory;
//   drain (FP64 pipe): two stack entries per iteration, branch-free: exact minimum-image r2 in the
//          oracle's rounding sequence, exact cull r2 <= rc2 (bit-exact pair set), interaction(),
//          predicated accumulation of f, E, W.  When a stack nears capacity only the excess is
//          popped (from every lane), so lanes stay evenly loaded until the final drain.
// Pair-list reuse (MODE): while the binning is valid (atoms moved < skin/2) the set of pairs within
// rc + skin cannot grow, so the scan result is kept in global memory as rows of 32 staged indices:
//   MODE 1 (build, the evaluation right after a re-binning): the scan accepts r <= rc + skin and every
//          entry popped by the drain is also stored (one coalesced 64-byte row per drain iteration);
//   MODE 2 (use, the following steps): the window scan is replaced by a walk over the stored rows
//          (~90 entries per atom instead of ~750 candidates), re-tested in FP32 against rc;
//   MODE 0: plain window scan (single-point evaluations, audits).
// This is the Verlet-list step SURVEY section 8(f) ranks first (the direction find_action_partners1!
// was heading, src/cells.jl:224-297).
// The FP64 pipe (64 lanes/clk/SM) therefore only sees pairs that are inside the cutoff (plus a
// 1e-3 margin) at high lane occupancy, instead of the ~10x larger candidate set.
#pragma once
#include "lj_pair.cuh"

#define FC_QCAP 64          // per-lane queue entries (uint16 staged indices)
#define FC_QCHUNK 16        // candidates scanned between queue-capacity checks
#define FC_MINPOP 8         // a partial drain pops at least this many entries per lane
#define FC_MAX_HOMEROWS 64  // by*bz of the largest supported brick
#define FC_MAX_TYPES 16     // LJ parameter classes held as a pair table in shared memory

#define FC_MAX_BLOCK 384     // launch bound: 12 warps, up to 170 registers per thread

struct CellArgs {
    GridDesc g;
    const int32_t *cell_start;   // local cells + 1
    const double *sx, *sy, *sz, *hs, *ts;
    const int32_t *id, *type, *xbase;
    const uint64_t *xmask;
    const double2 *ljtab;             // ntypes^2 entries {half_sigma_a + half_sigma_b, twice_sqrt_eps_a * twice_sqrt_eps_b}
    int ntypes;                       // 0: more than FC_MAX_TYPES classes, per-atom parameters are gathered from global
    double *fx, *fy, *fz, *en, *vir;
    double *partial;                  // per warp: {sum e_i, sum w_i}
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
    uint16_t *list;                   // pair-list rows: list[((brick*tmax + task)*lcap + row)*32 + lane]
    int32_t *list_rows;               // rows stored per task
    int tmax, lcap;                   // task slots per brick, row capacity per task
    float rl2f;                       // FP32 threshold of the list: (rc + skin)^2 * (1 + margin)
};

__host__ __device__ inline size_t fc_smem_bytes(int cap, int ncs_max, int block, bool typed)
{
    size_t b = (size_t)cap * (2 * sizeof(double2) + sizeof(float4));
    if (!typed) b += (size_t)cap * sizeof(int);      // slot of every staged atom (per-atom parameter gathers)
    b += (size_t)FC_MAX_TYPES * FC_MAX_TYPES * sizeof(double2);
    b += (size_t)(ncs_max + 1) * sizeof(int) * 2;    // cs[], gbase[]
    b += 2 * (FC_MAX_HOMEROWS + 1) * sizeof(int);    // hstart[], tstart[]
    b += 8 * sizeof(int);                            // scalars
    b = (b + 15) & ~(size_t)15;
    b += (size_t)FC_QCAP * block * sizeof(uint16_t);
    return b;
}

__device__ __forceinline__ int wrap_mod(int a, int M)
{
    a %= M;
    return a < 0 ? a + M : a;
}

template <bool F, bool EW, bool EXCL, bool AUDIT, bool TYPED, int MODE>
__global__ void __launch_bounds__(FC_MAX_BLOCK, 1) k_force_cells(CellArgs a)
{
    const int BLOCK = blockDim.x;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GridDesc &g = a.g;
    const int cap = a.cap;
    double2 *pxy = reinterpret_cast<double2 *>(smem_raw);      // {s_x, s_y}, s = r/L (src/nonbonded.jl:60-61)
    double2 *pzm = pxy + cap;                                  // {s_z, (type << 32) | id}
    float4 *prel = reinterpret_cast<float4 *>(pzm + cap);
    constexpr bool typed = TYPED;    // LJ classes through the shared-memory pair table; else per-atom gathers
    int *sslot = reinterpret_cast<int *>(prel + cap);
    double2 *ljt = reinterpret_cast<double2 *>(sslot + (typed ? 0 : cap));
    int *cs = reinterpret_cast<int *>(ljt + FC_MAX_TYPES * FC_MAX_TYPES);
    int *gbase = cs + (a.ncs_max + 1);
    int *hstart = gbase + (a.ncs_max + 1);
    int *tstart = hstart + (FC_MAX_HOMEROWS + 1);
    int *scal = tstart + (FC_MAX_HOMEROWS + 1);
    uint16_t *queue = reinterpret_cast<uint16_t *>(
        smem_raw + ((reinterpret_cast<unsigned char *>(scal + 8) - smem_raw + 15) & ~(size_t)15));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NW = BLOCK / 32;
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
    if (typed)
        for (int t = tid; t < a.ntypes * a.ntypes; t += BLOCK) ljt[t] = a.ljtab[t];
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
            // home rows: the atoms of the home cells of row (hy,hz) are one contiguous staged range;
            // warp tasks: chunks of 32 consecutive atoms of one home row
            int h = 0, t = 0;
            for (int hz = 0; hz < nhz; hz++)
                for (int hy = 0; hy < nhy; hy++) {
                    hstart[hz * nhy + hy] = h;
                    tstart[hz * nhy + hy] = t;
                    const int row = (hz + R) * syn + (hy + R);
                    const int n = cs[row * sxn + R + nhx] - cs[row * sxn + R];
                    h += n;
                    t += (n + 31) >> 5;
                }
            hstart[nhy * nhz] = h;
            tstart[nhy * nhz] = t;
            scal[0] = h;      // home atoms
            scal[1] = run;    // staged atoms
            scal[2] = t;      // warp tasks
            scal[3] = 0;      // task cursor
            if (run > cap) atomicCAS(a.err, 0, 2);
        }
    }
    __syncthreads();
    const int nstaged = min(scal[1], cap);
    const int ntasks = scal[2];

    // ---- phase B: stage atoms row by row (each row = up to two contiguous slot ranges) --------
    for (int piece = warp; piece < 2 * nrows; piece += NW) {
        const int row = piece >> 1, second = piece & 1;
        const int cy = row % syn, cz = row / syn;
        const int ux0 = hx0 - R;                       // unwrapped x of staged cx = 0
        const int gx0 = wrap_mod(ux0, M);
        const int len1 = min(sxn, M - gx0);            // cells before the periodic wrap
        const int cfirst = second ? len1 : 0, clast = second ? sxn : len1;
        if (cfirst >= clast) continue;
        const int ibeg = cs[row * sxn + cfirst], iend = min(cs[row * sxn + clast], cap);
        const int sbeg = gbase[row * sxn + cfirst];
        // cell centre relative to the brick centre, in cell units
        const double ccy = (double)(cy - R) + 0.5 - 0.5 * nhy, ccz = (double)(cz - R) + 0.5 - 0.5 * nhz;
        const double csy = ((double)(hy0 - R + cy) + 0.5) / M;                      // unwrapped scaled centre
        const int uz = (g.zwrap ? hz0 : g.zglob0 + hz0) - R + cz;
        const double csz = ((double)uz + 0.5) / M;
        for (int idx = ibeg + lane; idx < iend; idx += 32) {
            const int slot = sbeg + (idx - ibeg);
            int cx = cfirst;
            while (cx + 1 < clast && cs[row * sxn + cx + 1] <= idx) cx++;
            const double sxv = a.sx[slot], syv = a.sy[slot], szv = a.sz[slot];
            pxy[idx] = make_double2(sxv, syv);
            pzm[idx] = make_double2(szv, __hiloint2double(typed ? a.type[slot] : 0, a.id[slot]));
            if (!typed) sslot[idx] = slot;
            const double csx = ((double)(ux0 + cx) + 0.5) / M;
            double dx = sxv - csx, dy = syv - csy, dz = szv - csz;
            dx -= rint(dx); dy -= rint(dy); dz -= rint(dz);      // image nearest to the staged cell
            float4 p;
            p.x = (float)(a.L * dx + ((double)(cx - R) + 0.5 - 0.5 * nhx) * a.cell_edge);
            p.y = (float)(a.L * dy + ccy * a.cell_edge);
            p.z = (float)(a.L * dz + ccz * a.cell_edge);
            p.w = fmaf(p.z, p.z, fmaf(p.y, p.y, p.x * p.x));
            prel[idx] = p;
        }
    }
    __syncthreads();

    // ---- phase C: warp tasks -----------------------------------------------------------------
    const double c60id2 = 60.0 * a.model.id2;
    const double rc2 = a.model.rc2, L = a.L;
    const float rc2f = a.rc2f;
    const float scan2f = MODE == 1 ? a.rl2f : a.rc2f;     // what the window scan accepts
    const int nwin = 2 * R + 1;
    double esum = 0, wsum = 0;
    unsigned long long npair = 0, hsum = 0, hxor = 0;

    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&scal[3], 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= ntasks) break;
        int hr = 0;
        while (tstart[hr + 1] <= t) hr++;
        const int cyi = hr % nhy + R, czi = hr / nhy + R;
        const int hrow = czi * syn + cyi;
        const int a0 = cs[hrow * sxn + R] + ((t - tstart[hr]) << 5);
        const int a1 = min(a0 + 32, cs[hrow * sxn + R + nhx]);
        int cxa = R, cxb = R;
        while (cs[hrow * sxn + cxa + 1] <= a0) cxa++;
        while (cs[hrow * sxn + cxb + 1] <= a1 - 1) cxb++;
        const int self = a0 + lane;
        const bool active = self < a1;
        const int me = active ? self : a0;

        const float4 pi = prel[me];
        // r2 = |c|^2 + (|p|^2 - 2 c.p); an inactive lane gets |p|^2 = 1e30 and never accepts anything
        const float m2x = -2.0f * pi.x, m2y = -2.0f * pi.y, m2z = -2.0f * pi.z, pp = active ? pi.w : 1e30f;
        int cxi = cxa;
        while (cs[hrow * sxn + cxi + 1] <= me) cxi++;
        const int slot_i = gbase[hrow * sxn + cxi] + (me - cs[hrow * sxn + cxi]);
        const double2 q0 = pxy[me], q1 = pzm[me];
        const double six = q0.x, siy = q0.y, siz = q1.x;
        const int32_t idi = __double2loint(q1.y), typi = __double2hiint(q1.y);
        double hsi = 0, tsi = 0;
        if (!typed) { hsi = a.hs[slot_i]; tsi = a.ts[slot_i]; }
        const double2 *ljrow = ljt + typi * a.ntypes;
        int32_t xb = 0; uint64_t xm = 0;
        if (EXCL) { xb = a.xbase[slot_i]; xm = a.xmask[slot_i]; }
        double fx = 0, fy = 0, fz = 0, e = 0, w = 0;
        uint16_t *lrow = nullptr;               // this task's pair-list rows, lane's column
        int nrow = 0;                           // rows stored so far (MODE 1)
        if (MODE != 0) {
            if (t >= a.tmax) { atomicCAS(a.err, 0, 5); break; }
            lrow = a.list + ((size_t)(bid * a.tmax + t) * a.lcap) * 32 + lane;
        }
        int cnt = 0;                            // entries on this lane's stack: queue[k*BLOCK + tid], k < cnt
        uint16_t *qp = queue + tid;             // next free entry

        // one stack entry, branch-free: invalid or culled entries contribute nothing
        auto pair_eval = [&](int idx, bool valid, int row) {
            const int j = valid ? (int)queue[idx * BLOCK + tid] : me;
            if (MODE == 1) {
                if (row < a.lcap) lrow[(size_t)row * 32] = valid ? (uint16_t)j : (uint16_t)0xFFFF;
                else atomicCAS(a.err, 0, 5);
            }
            const double2 j0 = pxy[j], j1 = pzm[j];
            double vx, vy, vz;
            const double r2 = min_image_r2(six, siy, siz, j0.x, j0.y, j1.x, L, vx, vy, vz);
            bool ok = valid && (j != me) && (r2 <= rc2);
            const int32_t idj = __double2loint(j1.y);
            if (EXCL) ok = ok && !pair_excluded(xb, xm, idj);
            double sig, tt;
            if (!typed) {
                const int slot_j = sslot[j];
                sig = hsi + a.hs[slot_j];
                tt = tsi * a.ts[slot_j];
            } else {
                const double2 pr = ljrow[__double2hiint(j1.y)];   // one class: every lane reads entry 0 (broadcast)
                sig = pr.x; tt = pr.y;
            }
            const double r2s = ok ? r2 : 1.0;
            const double inv = rcp_fast(r2s);
            double Eg, Wg;
            lj_interaction(r2s, inv, sig, tt, a.model, c60id2, Eg, Wg);
            if (F) {
                const double qf = ok ? Wg * inv : 0.0;
                fx = fma(qf, vx, fx); fy = fma(qf, vy, fy); fz = fma(qf, vz, fz);
            }
            if (EW) { e += ok ? Eg : 0.0; w += ok ? Wg : 0.0; }
            if (AUDIT && ok && idi < idj) {
                const uint64_t hh = pair_hash(idi, idj);
                npair++; hsum += hh; hxor ^= hh;
                if (a.pairs) {
                    const unsigned long long k = atomicAdd(a.pair_n, 1ull);
                    if ((long long)k < a.pair_cap) { a.pairs[2 * k] = idi; a.pairs[2 * k + 1] = idj; }
                }
            }
        };
        // pop the newest `depth` entries of every lane (all of them when depth >= the fullest stack)
        auto drain = [&](int depth) {
            for (int k = 0; k < depth; k += 4) {      // four independent pair evaluations in flight
                pair_eval(cnt - 1 - k, k < cnt, nrow);
                pair_eval(cnt - 2 - k, k + 1 < cnt && k + 1 < depth, nrow + 1);
                pair_eval(cnt - 3 - k, k + 2 < cnt && k + 2 < depth, nrow + 2);
                pair_eval(cnt - 4 - k, k + 3 < cnt && k + 3 < depth, nrow + 3);
                nrow += 4;
            }
            cnt = max(cnt - depth, 0);
            qp = queue + cnt * BLOCK + tid;
        };

        // -------- scan: (2R+1)^2 rows, one shared window of cells [cxa-R, cxb+R] per row --------
        for (int rw = 0; MODE != 2 && rw < nwin * nwin; rw++) {
            const int row = (czi + rw / nwin - R) * syn + (cyi + rw % nwin - R);
            const int p0 = cs[row * sxn + cxa - R];
            const int p1 = min(cs[row * sxn + cxb + R + 1], nstaged);
            for (int pb = p0; pb < p1; pb += FC_QCHUNK) {
                const int pe = min(pb + FC_QCHUNK, p1);
                // four candidates per iteration: the loads are issued together (one-address broadcasts)
                int p = pb;
                for (; p + 4 <= pe; p += 4) {
                    const float4 c0 = prel[p], c1 = prel[p + 1], c2 = prel[p + 2], c3 = prel[p + 3];
                    const float r0 = fmaf(c0.x, m2x, fmaf(c0.y, m2y, fmaf(c0.z, m2z, c0.w + pp)));
                    const float r1 = fmaf(c1.x, m2x, fmaf(c1.y, m2y, fmaf(c1.z, m2z, c1.w + pp)));
                    const float r2 = fmaf(c2.x, m2x, fmaf(c2.y, m2y, fmaf(c2.z, m2z, c2.w + pp)));
                    const float r3 = fmaf(c3.x, m2x, fmaf(c3.y, m2y, fmaf(c3.z, m2z, c3.w + pp)));
                    if (r0 <= scan2f) { *qp = (uint16_t)p; qp += BLOCK; cnt++; }
                    if (r1 <= scan2f) { *qp = (uint16_t)(p + 1); qp += BLOCK; cnt++; }
                    if (r2 <= scan2f) { *qp = (uint16_t)(p + 2); qp += BLOCK; cnt++; }
                    if (r3 <= scan2f) { *qp = (uint16_t)(p + 3); qp += BLOCK; cnt++; }
                }
                for (; p < pe; p++) {
                    const float4 c = prel[p];
                    const float r = fmaf(c.x, m2x, fmaf(c.y, m2y, fmaf(c.z, m2z, c.w + pp)));
                    if (r <= scan2f) { *qp = (uint16_t)p; qp += BLOCK; cnt++; }
                }
                const int over = __reduce_max_sync(0xffffffffu, cnt) - (FC_QCAP - FC_QCHUNK);
                if (over > 0) drain(max(over, FC_MINPOP));
            }
        }
        // -------- MODE 2: walk the stored pair-list rows instead of the window --------
        if (MODE == 2) {
            const int nrows = min(a.list_rows[bid * a.tmax + t], a.lcap);
            uint16_t ent[FC_QCHUNK], nxt[FC_QCHUNK];
#pragma unroll
            for (int k = 0; k < FC_QCHUNK; k++) ent[k] = (k < nrows) ? lrow[(size_t)k * 32] : (uint16_t)0xFFFF;
            for (int k0 = 0; k0 < nrows; k0 += FC_QCHUNK) {
                // the next chunk's rows are requested before this chunk is processed (hides the L2/HBM latency)
#pragma unroll
                for (int k = 0; k < FC_QCHUNK; k++)
                    nxt[k] = (k0 + FC_QCHUNK + k < nrows) ? lrow[(size_t)(k0 + FC_QCHUNK + k) * 32] : (uint16_t)0xFFFF;
#pragma unroll
                for (int k = 0; k < FC_QCHUNK; k++) {
                    const int j = ent[k];
                    const float4 c = prel[j == 0xFFFF ? me : j];
                    const float r = fmaf(c.x, m2x, fmaf(c.y, m2y, fmaf(c.z, m2z, c.w + pp)));
                    if (j != 0xFFFF && r <= rc2f) { *qp = (uint16_t)j; qp += BLOCK; cnt++; }
                }
#pragma unroll
                for (int k = 0; k < FC_QCHUNK; k++) ent[k] = nxt[k];
                const int over = __reduce_max_sync(0xffffffffu, cnt) - (FC_QCAP - FC_QCHUNK);
                if (over > 0) drain(max(over, FC_MINPOP));
            }
        }
        drain(__reduce_max_sync(0xffffffffu, cnt));
        if (MODE == 1 && lane == 0) a.list_rows[bid * a.tmax + t] = nrow;

        if (active) {
            if (F) { a.fx[slot_i] = fx; a.fy[slot_i] = fy; a.fz[slot_i] = fz; }
            if (EW) {
                e *= 0.5; w *= 0.5;                         // src/nonbonded.jl:93-94
                a.en[slot_i] = e; a.vir[slot_i] = w;
                esum += e; wsum += w;
            }
        }
    }

    // ---- per-warp totals (summed later in a fixed order: deterministic) ---------------------------
    if (EW) {
        esum = warp_sum_f64(esum); wsum = warp_sum_f64(wsum);
        if (lane == 0) { a.partial[2 * (bid * NW + warp)] = esum; a.partial[2 * (bid * NW + warp) + 1] = wsum; }
    }
    if (AUDIT) {
        for (int o = 16; o > 0; o >>= 1) {
            npair += __shfl_xor_sync(0xffffffffu, npair, o);
            hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
            hxor ^= __shfl_xor_sync(0xffffffffu, hxor, o);
        }
        if (lane == 0 && npair) { atomicAdd(a.digest, npair); atomicAdd(a.digest + 1, hsum); atomicXor(a.digest + 2, hxor); }
    }
}

// Sum the per-warp partials in index order: deterministic totals.
__global__ void k_reduce_partials(int n, const double *__restrict__ partial, double *__restrict__ totals)
{
    __shared__ double sE[256], sW[256];
    double E = 0, W = 0;
    for (int b = threadIdx.x; b < n; b += 256) { E += partial[2 * b]; W += partial[2 * b + 1]; }
    sE[threadIdx.x] = E; sW[threadIdx.x] = W;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sE[threadIdx.x] += sE[threadIdx.x + o]; sW[threadIdx.x] += sW[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = sE[0]; totals[1] = sW[0]; }
}
