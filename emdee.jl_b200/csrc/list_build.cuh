// list_build.cuh -- builds the pair list (Verlet list with skin) that k_force_list walks.
//
// Runs once per re-binning.  The direction find_action_partners1! was heading (src/cells.jl:224-297: for every
// atom, the partners inside its 27 / 125 neighbouring cells), restated for the GPU: one block per home brick,
// the brick and its halo staged in shared memory in exactly the order k_force_list stages them, so a stored
// entry (staged index + 1, 16 bits) means the same atom in both kernels.
//
// FP32 only, no force evaluation: the kernel is a filter.  Warp task = 32 consecutive home atoms of one home row,
// one per lane; all lanes walk the same candidate window (the (2R+1)^2 staged rows around the task's row, cells
// [first-R, last+R]), so every LDS.128 is a one-address broadcast.  Per candidate: |c|^2 - 2 c.p + |p|^2 against
// (rc + skin)^2 plus a bound on the FP32 rounding (conservative: a listed pair may be outside, a pair inside is
// never missed); accepted candidates (exclusions removed here, once, instead of on every step) are shifted into a
// 16-byte register buffer and stored as one chunk per eight entries:
//   list8[((brick*gmax + h/32)*lcap8 + chunk)*32 + h%32],  h = index of the home atom in the brick's home list.
#pragma once
#include "force_cells.cuh"

#define LB_MAX_BLOCK 256

__host__ __device__ inline size_t lb_smem_bytes(int cap, int ncs_max, bool excl)
{
    size_t b = (size_t)cap * sizeof(float4);
    if (excl) b += (size_t)cap * sizeof(int);          // global id of every staged atom
    b += FC_DIMTAB * sizeof(double);
    b += (size_t)(ncs_max + 1) * sizeof(int) * 3;      // cs[], gbase[], ccoord[]
    b += 2 * (FC_MAX_HOMEROWS + 1) * sizeof(int);      // hstart[], tstart[]
    b += 8 * sizeof(int);
    return (b + 15) & ~(size_t)15;
}

template <bool EXCL>
__global__ void __launch_bounds__(LB_MAX_BLOCK, 3) k_list_build(CellArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GridDesc &g = a.g;
    const int cap = a.cap;
    float4 *prel = reinterpret_cast<float4 *>(smem_raw);       // {x, y, z, |p|^2} in the brick's frame
    int *pid = reinterpret_cast<int *>(prel + cap);
    double *ctab = reinterpret_cast<double *>(pid + (EXCL ? cap : 0));
    int *cs = reinterpret_cast<int *>(ctab + FC_DIMTAB);
    int *gbase = cs + (a.ncs_max + 1);
    int *ccoord = gbase + (a.ncs_max + 1);
    int *hstart = ccoord + (a.ncs_max + 1);
    int *tstart = hstart + (FC_MAX_HOMEROWS + 1);
    int *scal = tstart + (FC_MAX_HOMEROWS + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = g.R;
    const int bid = blockIdx.x + a.block_first;
    const BrickGeom bg = brick_geom(g, bid);
    const int nhx = bg.nhx, nhy = bg.nhy, nhz = bg.nhz;
    const int sxn = bg.sxn, syn = bg.syn, ncs = bg.ncs;

    stage_cell_table(a, bg, cs, gbase, ccoord, ctab);
    __syncthreads();
    if (warp == 0) {
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
            scal[1] = run;    // staged atoms
            scal[2] = t;      // warp tasks
            scal[3] = 0;      // task cursor
            if (run > cap) atomicCAS(a.err, 0, 2);
        }
    }
    __syncthreads();
    const int nstaged = min(scal[1], cap);
    const int ntasks = scal[2];

    stage_atoms(a, bg, cs, gbase, ccoord, ctab, 0, nstaged, [&](int idx, int slot, double px, double py, double pz) {
        float4 p;
        p.x = (float)px; p.y = (float)py; p.z = (float)pz;
        p.w = fmaf(p.z, p.z, fmaf(p.y, p.y, p.x * p.x));
        prel[idx] = p;
        if (EXCL) pid[idx] = a.id[slot];
    });
    __syncthreads();

    const float rl2f = a.rl2f;
    const int nwin = 2 * R + 1;
    const int lmax = a.lcap8 * 8;
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
        int32_t xb = 0; uint64_t xm = 0;
        if (EXCL) {
            int cxi = cxa;
            while (cs[hrow * sxn + cxi + 1] <= me) cxi++;
            const int slot_i = gbase[hrow * sxn + cxi] + (me - cs[hrow * sxn + cxi]);
            xb = a.xbase[slot_i]; xm = a.xmask[slot_i];
        }
        const int h = hstart[hr] + (me - cs[hrow * sxn + R]);
        if ((h >> 5) >= a.gmax) { atomicCAS(a.err, 0, 5); break; }
        const size_t gs = (size_t)bid * a.gmax + (h >> 5);
        uint4 *lp = a.list8 + gs * a.lcap8 * 32 + (h & 31);
        uint4 lb = make_uint4(0u, 0u, 0u, 0u);      // the chunk being filled: new entries enter at the top, zeros (dummy) below
        int nlist = 0;

        auto accept = [&](int p) {
            bool take = p != me;
            if (EXCL) take = take && !pair_excluded(xb, xm, pid[p]);
            if (take) {
                lb.x = __funnelshift_r(lb.x, lb.y, 16); lb.y = __funnelshift_r(lb.y, lb.z, 16);
                lb.z = __funnelshift_r(lb.z, lb.w, 16); lb.w = (lb.w >> 16) | ((unsigned)(p + 1) << 16);
                nlist++;
                if ((nlist & 7) == 0) {           // one 16-byte store per eight entries
                    if (nlist <= lmax) lp[(size_t)((nlist >> 3) - 1) * 32] = lb;
                    lb = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        };

        for (int rw = 0; rw < nwin * nwin; rw++) {
            const int row = (czi + rw / nwin - R) * syn + (cyi + rw % nwin - R);
            const int p0 = cs[row * sxn + cxa - R];
            const int p1 = min(cs[row * sxn + cxb + R + 1], nstaged);
            int p = p0;
            for (; p + 4 <= p1; p += 4) {      // four candidates per iteration: the broadcast loads are issued together
                const float4 c0 = prel[p], c1 = prel[p + 1], c2 = prel[p + 2], c3 = prel[p + 3];
                const float r0 = fmaf(c0.x, m2x, fmaf(c0.y, m2y, fmaf(c0.z, m2z, c0.w + pp)));
                const float r1 = fmaf(c1.x, m2x, fmaf(c1.y, m2y, fmaf(c1.z, m2z, c1.w + pp)));
                const float r2 = fmaf(c2.x, m2x, fmaf(c2.y, m2y, fmaf(c2.z, m2z, c2.w + pp)));
                const float r3 = fmaf(c3.x, m2x, fmaf(c3.y, m2y, fmaf(c3.z, m2z, c3.w + pp)));
                if (r0 <= rl2f) accept(p);
                if (r1 <= rl2f) accept(p + 1);
                if (r2 <= rl2f) accept(p + 2);
                if (r3 <= rl2f) accept(p + 3);
            }
            for (; p < p1; p++) {
                const float4 c = prel[p];
                const float r = fmaf(c.x, m2x, fmaf(c.y, m2y, fmaf(c.z, m2z, c.w + pp)));
                if (r <= rl2f) accept(p);
            }
        }
        if (active) {
            if (nlist > lmax) { atomicCAS(a.err, 0, 5); nlist = lmax; }
            a.list_n[gs * 32 + (h & 31)] = (uint16_t)nlist;
            if (nlist & 7) lp[(size_t)(nlist >> 3) * 32] = lb;     // last chunk: the unused entries point at the dummy atom
        }
    }
}
