// force_list.cuh -- the stepping kernel: Lennard-Jones forces from the stored pair list, FP64.
//
// Block-per-brick form; the default is the persistent form in force_list_p.cuh (same walk / drain), this one runs
// when two staging buffers do not fit in shared memory (EMDEE_PERSIST=0 forces it).  k_list_build (list_build.cuh)
// builds the list on a re-binning step.  Same decomposition as k_force_cells -- one block per home brick, the brick and its
// halo staged in shared memory in the same order, so a list entry (staged index + 1) means the same atom --
// but organised around what ncu showed to bound the kernel (profiles/): the FP64 pipe, the shared-memory
// crossbar (a random 16-byte gather costs ~10 wavefronts per warp) and the issue slots, in that order.
//
//   staged per atom: {x, y} (16 B) + z (8 B) in FP64 in the brick's frame, and an FP16 copy {x, y, z, 0}
//     (8 B) for the pre-cull; index 0 is a dummy atom far away, so padding entries need no validity test;
//   warp task: 32 home atoms of the brick's flattened home list (one per lane, full-neighbour, no atomics),
//     handed out dynamically;
//   walk:  each lane streams its own chunks (LDG.128 = 8 entries, two chunks in flight), tests every entry
//     in packed FP16 against a conservative threshold (7 FMA-pipe/ALU instructions and one 8-byte gather per
//     entry) and pushes the survivors (~64 % of the list for skin = 0.4) on its stack in shared memory;
//   drain: four stack entries per iteration: 24-byte FP64 gather, three subtractions, r2, cutoff decision
//     by integer compares (exact rounding sequence only inside a 3e-6 band around rc2, so the pair set is
//     bit-exact), lj_pair_q (22 FP64 instructions), predicated accumulation.  Partial drains pop only the
//     excess, so lanes stay evenly loaded.
// Exclusions were applied when the list was built; LJ classes come from the shared-memory pair table.
#pragma once
#include <cuda_fp16.h>

#include "force_cells.cuh"

#ifndef FL_QCAP
#define FL_QCAP 48          // per-lane stack entries
#endif
#ifndef FL_MINPOP
#define FL_MINPOP 8          // a partial drain pops at least this many entries per lane
#endif
#define FL_MAX_BLOCK 256
#ifndef FL_AHEAD
#define FL_AHEAD 4          // list chunks requested into L2 ahead of the register loads
#endif

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__host__ __device__ inline size_t fl_smem_bytes(int cap, int ncs_max, int block, int ntypes)
{
    const size_t cap1 = (size_t)cap + 1;                       // index 0 is the dummy atom
    size_t b = cap1 * (sizeof(double2) + sizeof(double) + sizeof(uint2));
    if (ntypes > 1) b += (cap1 + 15) & ~(size_t)15;            // uint8 LJ class per staged atom
    b = (b + 15) & ~(size_t)15;
    b += (size_t)ntypes * ntypes * sizeof(double2);
    b += FC_DIMTAB * sizeof(double);
    b += (size_t)(ncs_max + 1) * sizeof(int) * 3;              // cs[], gbase[], ccoord[]
    b += (FC_MAX_HOMEROWS + 1) * sizeof(int);                  // hstart[]
    b += 8 * sizeof(int);
    b = (b + 15) & ~(size_t)15;
    b += (size_t)(FL_QCAP + 1) * block * sizeof(uint16_t);     // + one guard row of dummy entries below the stacks
    return b;
}

// Everything a lane needs to redo its whole list the careful way (see careful_lane).
struct LaneRedo {
    const double2 *pxy;
    const double *pz;
    const uint8_t *ptyp;
    const double2 *ljt;
    const int *cs, *gbase;       // staged-cell table (k_force_list) ...
    const int2 *recipe;          // ... or the staging recipe (k_force_list_p): global slot of a staged atom
    int ncs, ntypes, me, slot_i, nent;
    const uint16_t *entries;     // this lane's first chunk, viewed as uint16 (chunk stride 256)
    const double *sx, *sy, *sz;  // scaled coordinates in slot order (the oracle's inputs)
    double L;
    LJModel model;
    LJFast fast;
    int rc2hi;
};

// The hot loop decides "inside the cutoff" with integer compares on a local-frame r2 and counts a pair whose r2
// is within 3*2^-20 of rc2 as outside, only remembering that it met one.  A lane that met one (about one lane
// in a hundred warp tasks for a fluid; every lane for a lattice with a shell exactly at rc) discards its result
// and re-evaluates its list here, taking the oracle's exact decision (and the oracle's clamped x) for those pairs.
template <bool MULTI, bool EW = false, bool CG = false>
__device__ __noinline__ void careful_lane(const LaneRedo &w, double *f, unsigned long long *npair)
{
    const double2 q0 = w.pxy[w.me];
    const double pix = q0.x, piy = q0.y, piz = w.pz[w.me];
    double fx = 0, fy = 0, fz = 0, e = 0, vw = 0;
    unsigned long long n = 0;
    const double2 *ljrow = w.ljt;
    if (MULTI) ljrow = w.ljt + (int)w.ptyp[w.me] * w.ntypes;
    const int nslots = ((w.nent + 7) >> 3) << 3;       // the last chunk is padded with zeros in any position
    for (int k = 0; k < nslots; k++) {
        const int j = w.entries[((k >> 3) << 8) + (k & 7)];
        if (j == 0) continue;
        const double2 j0 = w.pxy[j];
        const double vx = pix - j0.x, vy = piy - j0.y, vz = piz - w.pz[j];
        const double r2 = fma(vz, vz, fma(vy, vy, vx * vx));
        const int where = pair_in_range(r2, w.rc2hi);
        if (where > 0) continue;
        bool xover = false;
        double xval = 0.0;
        if (where == 0) {
            const int slot_j = w.recipe ? w.recipe[j].x : staged_slot(j, w.cs, w.gbase, w.ncs);
            if (!exact_in_range<CG>(w.sx, w.sy, w.sz, w.slot_i, slot_j, w.L, w.model, &xval)) continue;
            xover = true;
        }
        double2 pr = ljrow[0];
        if (MULTI) pr = ljrow[w.ptyp[j]];
        double Eg = 0, Wg = 0;
        const double qf = lj_pair_q<EW>(r2, pr.x, pr.y, w.fast, xover, xval, Eg, Wg);
        fx = fma(qf, vx, fx); fy = fma(qf, vy, fy); fz = fma(qf, vz, fz);
        if (EW) { e += Eg; vw += Wg; }
        n++;
    }
    f[0] = fx; f[1] = fy; f[2] = fz;
    if (EW) { f[3] = e; f[4] = vw; }
    *npair = n;
}

// ILP = stack entries evaluated per drain iteration.  The drain is a chain of ~20 dependent FP64 operations per
// pair, so what keeps the FP64 pipe busy is the number of independent chains per scheduler: registers per
// scheduler / (registers per chain).  ILP 8 with 192-thread blocks (<= 170 registers) gives 3 warps x 8 chains,
// ILP 4 with 256-thread blocks (<= 128 registers) 4 warps x 4.
template <bool MULTI, bool COUNT, int ILP>
__global__ void __launch_bounds__(ILP == 8 ? 192 : FL_MAX_BLOCK, 2) k_force_list(CellArgs a)
{
    const int BLOCK = blockDim.x;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GridDesc &g = a.g;
    const int cap1 = a.cap + 1;
    double2 *pxy = reinterpret_cast<double2 *>(smem_raw);
    double *pz = reinterpret_cast<double *>(pxy + cap1);
    uint2 *ph = reinterpret_cast<uint2 *>(pz + cap1);
    uint8_t *ptyp = reinterpret_cast<uint8_t *>(ph + cap1);
    size_t off = (size_t)(reinterpret_cast<unsigned char *>(ptyp) - smem_raw);
    if (MULTI) off += ((size_t)cap1 + 15) & ~(size_t)15;
    off = (off + 15) & ~(size_t)15;
    double2 *ljt = reinterpret_cast<double2 *>(smem_raw + off);
    double *ctab = reinterpret_cast<double *>(ljt + a.ntypes * a.ntypes);
    int *cs = reinterpret_cast<int *>(ctab + FC_DIMTAB);
    int *gbase = cs + (a.ncs_max + 1);
    int *ccoord = gbase + (a.ncs_max + 1);
    int *hstart = ccoord + (a.ncs_max + 1);
    int *scal = hstart + (FC_MAX_HOMEROWS + 1);
    uint16_t *qguard = reinterpret_cast<uint16_t *>(
        smem_raw + ((reinterpret_cast<unsigned char *>(scal + 8) - smem_raw + 15) & ~(size_t)15));
    uint16_t *queue = qguard + BLOCK;          // row -1 holds dummy entries: popping an empty stack yields the dummy atom

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NW = BLOCK / 32;
    const int R = g.R;
    const int bid = FC_BRICK_OF(a, (int)blockIdx.x);
    const BrickGeom bg = brick_geom(g, bid);
    qguard[tid] = 0;

    // The first task of a warp is its own index, later ones come from the block's counter.  A task's entry count
    // and first two chunks are requested one task ahead (here: before the staging), so the L2/HBM latency of the
    // list is off the critical path.  list_n is zero for lanes without an atom (cleared before every build).
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    int grp = warp;
    int pre_n = 0;
    uint4 pre0 = zero4, pre1 = zero4;
    if (grp < a.gmax) {
        const size_t gs = (size_t)bid * a.gmax + grp;
        pre_n = a.list_n[gs * 32 + lane];
        pre0 = a.list8[gs * a.lcap8 * 32 + lane];
        pre1 = a.list8[(gs * a.lcap8 + 1) * 32 + lane];
#pragma unroll
        for (int k = 2; k < 2 + FL_AHEAD; k++) prefetch_l2(a.list8 + (gs * a.lcap8 + k) * 32 + lane);
    }
    const int nhx = bg.nhx, nhy = bg.nhy, nhz = bg.nhz;
    const int sxn = bg.sxn, syn = bg.syn, ncs = bg.ncs;

    // ---- phase A: staged-cell table; staged indices start at 1 ---------------------------------
    stage_cell_table(a, bg, cs, gbase, ccoord, ctab, tid, (int)blockDim.x);
    for (int t = tid; t < a.ntypes * a.ntypes; t += BLOCK) ljt[t] = a.ljtab[t];
    if (tid == 0) {
        pxy[0] = make_double2(1e30, 1e30);
        pz[0] = 1e30;
        const __half2 far = __floats2half2_rn(60000.0f, 60000.0f), farz = __floats2half2_rn(60000.0f, 0.0f);
        ph[0] = make_uint2(*reinterpret_cast<const unsigned *>(&far), *reinterpret_cast<const unsigned *>(&farz));
        if (MULTI) ptyp[0] = 0;
    }
    __syncthreads();
    if (warp == 0) {
        int run = 1;
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
            int h = 0;
            for (int hz = 0; hz < nhz; hz++)
                for (int hy = 0; hy < nhy; hy++) {
                    hstart[hz * nhy + hy] = h;
                    const int row = (hz + R) * syn + (hy + R);
                    h += cs[row * sxn + R + nhx] - cs[row * sxn + R];
                }
            hstart[nhy * nhz] = h;
            scal[0] = h;      // home atoms
            scal[1] = run;    // staged atoms + 1
            scal[3] = NW;     // task cursor (tasks 0..NW-1 are taken)
            if (run > cap1) atomicCAS(a.err, 0, 2);
        }
    }
    __syncthreads();
    const int nh = scal[0];
    const int ngroups = (nh + 31) >> 5;

    // ---- phase B: stage atoms (same order as k_force_cells, shifted by the dummy at index 0) -------
    stage_atoms(a, bg, cs, gbase, ccoord, ctab, 1, min(scal[1], cap1), tid, BLOCK, [&](int idx, int slot, int, double px, double py, double pzv) {
        pxy[idx] = make_double2(px, py);
        pz[idx] = pzv;
        const __half2 hxy = __floats2half2_rn((float)px, (float)py), hz0h = __floats2half2_rn((float)pzv, 0.0f);
        ph[idx] = make_uint2(*reinterpret_cast<const unsigned *>(&hxy), *reinterpret_cast<const unsigned *>(&hz0h));
        if (MULTI) ptyp[idx] = (uint8_t)a.type[slot];
    });
    __syncthreads();

    // ---- phase C: warp tasks = groups of 32 home atoms -------------------------------------------
    const __half thr = __float2half_ru(a.rc2h);
    unsigned long long npair = 0;
    double sig2_0 = 0, tt_0 = 0;
    if (!MULTI) { const double2 pr = ljt[0]; sig2_0 = pr.x; tt_0 = pr.y; }

    for (;;) {
        if (grp >= ngroups) break;
        if (grp >= a.gmax) { atomicCAS(a.err, 0, 5); break; }
        const int h = (grp << 5) + lane;
        const bool active = h < nh;
        const int hh = active ? h : (grp << 5);
        int hr = 0;
        while (hstart[hr + 1] <= hh) hr++;
        const int hrow = (hr / nhy + R) * syn + (hr % nhy + R);
        const int me = cs[hrow * sxn + R] + (hh - hstart[hr]);
        int cxi = R;
        while (cs[hrow * sxn + cxi + 1] <= me) cxi++;
        const int slot_i = gbase[hrow * sxn + cxi] + (me - cs[hrow * sxn + cxi]);
        EMDEE_CHECK(me >= 1 && me < scal[1] && hr < nhy * nhz && cxi < sxn, a.err);
        const double2 q0 = pxy[me];
        const double pix = q0.x, piy = q0.y, piz = pz[me];
        const uint2 hme = ph[me];
        const __half2 ixy = *reinterpret_cast<const __half2 *>(&hme.x), izw = *reinterpret_cast<const __half2 *>(&hme.y);
        const double2 *ljrow = ljt;
        if (MULTI) ljrow = ljt + (int)ptyp[me] * a.ntypes;
        double fx = 0, fy = 0, fz = 0;

        const int nent = pre_n;
        const int nch = (nent + 7) >> 3;
        const int nchmax = __reduce_max_sync(0xffffffffu, nch);
        const uint4 *lp = a.list8 + ((size_t)bid * a.gmax + grp) * a.lcap8 * 32 + lane;
        uint4 e0 = 0 < nch ? pre0 : zero4;
        uint4 e1 = 1 < nch ? pre1 : zero4;
        // claim the next task and request its head (registers) and its next chunks (L2 only: no register is tied up
        // and no scoreboard waits for a prefetch)
        int ngrp = 0;
        if (lane == 0) ngrp = atomicAdd(&scal[3], 1);
        ngrp = __shfl_sync(0xffffffffu, ngrp, 0);
        pre_n = 0;
        if (ngrp < ngroups && ngrp < a.gmax) {
            const size_t gs = (size_t)bid * a.gmax + ngrp;
            pre_n = a.list_n[gs * 32 + lane];
            pre0 = a.list8[gs * a.lcap8 * 32 + lane];
            pre1 = a.list8[(gs * a.lcap8 + 1) * 32 + lane];
#pragma unroll
            for (int k = 2; k < 2 + FL_AHEAD; k++) prefetch_l2(a.list8 + (gs * a.lcap8 + k) * 32 + lane);
        }
        int cnt = 0;
        uint16_t *qp = queue + tid;
        unsigned tmin = 0xffffffffu;            // smallest (unsigned)(hi(r2) - hi(rc2) + 1) seen: <= 2 means a borderline pair
        unsigned long long np = 0;

        // one stack entry, branch-free; an empty stack yields the dummy atom (far away, contributes nothing)
        auto pair_eval = [&](int idx) {
            const int j = queue[max(idx, -1) * BLOCK + tid];
            const double2 j0 = pxy[j];
            const double jz = pz[j];
            const double vx = pix - j0.x, vy = piy - j0.y, vz = piz - jz;
            const double r2 = fma(vz, vz, fma(vy, vy, vx * vx));
            const int t = __double2hiint(r2) - (a.rc2hi - 1);     // pair_in_range: t < 0 inside, t <= 2 borderline
            tmin = min(tmin, (unsigned)t);
            double sig2 = sig2_0, tt = tt_0;
            if (MULTI) { const double2 pr = ljrow[ptyp[j]]; sig2 = pr.x; tt = pr.y; }
            double Eg, Wg;
            double qf = lj_pair_q<false>(r2, sig2, tt, a.fast, false, 0.0, Eg, Wg);
            qf = t < 0 ? qf : 0.0;
            fx = fma(qf, vx, fx); fy = fma(qf, vy, fy); fz = fma(qf, vz, fz);
            if (COUNT) np += t < 0 ? 1 : 0;
        };
        // pop the newest `depth` (a multiple of ILP) entries of every lane
        auto drain = [&](int depth) {
            for (int k = 0; k < depth; k += ILP) {    // ILP independent pair evaluations in flight
#pragma unroll
                for (int u = 1; u <= ILP; u++) pair_eval(cnt - u - k);
            }
            cnt = max(cnt - depth, 0);
            qp = queue + cnt * BLOCK + tid;
        };
        auto test = [&](unsigned j, uint2 hj) {
            const __half2 dxy = __hsub2(*reinterpret_cast<const __half2 *>(&hj.x), ixy);
            const __half2 dzw = __hsub2(*reinterpret_cast<const __half2 *>(&hj.y), izw);
            const __half2 s = __hfma2(dzw, dzw, __hmul2(dxy, dxy));
            if (__hle(__hadd(__low2half(s), __high2half(s)), thr)) {
                EMDEE_CHECK(cnt < FL_QCAP, a.err);
                *qp = (uint16_t)j; qp += BLOCK; cnt++;
            }
        };

        for (int c = 0; c < nchmax; c++) {
            // chunk c+2 into registers (an L2 hit by now), chunk c+2+FL_AHEAD from HBM into L2
            const uint4 e2 = c + 2 < nch ? lp[(size_t)(c + 2) * 32] : zero4;
            if (c + 2 + FL_AHEAD < nch) prefetch_l2(lp + (size_t)(c + 2 + FL_AHEAD) * 32);
            const unsigned j0 = e0.x & 0xffffu, j1 = e0.x >> 16, j2 = e0.y & 0xffffu, j3 = e0.y >> 16;
            const unsigned j4 = e0.z & 0xffffu, j5 = e0.z >> 16, j6 = e0.w & 0xffffu, j7 = e0.w >> 16;
            EMDEE_CHECK((int)max(max(max(j0, j1), max(j2, j3)), max(max(j4, j5), max(j6, j7))) < scal[1], a.err);
            // the eight gathers are issued together, ahead of the stack stores they must not be reordered with
            const uint2 h0 = ph[j0], h1 = ph[j1], h2 = ph[j2], h3 = ph[j3], h4 = ph[j4], h5 = ph[j5], h6 = ph[j6], h7 = ph[j7];
            test(j0, h0); test(j1, h1); test(j2, h2); test(j3, h3);
            test(j4, h4); test(j5, h5); test(j6, h6); test(j7, h7);
            e0 = e1; e1 = e2;
            const int over = __reduce_max_sync(0xffffffffu, cnt) - (FL_QCAP - 8);
            if (over > 0) drain((max(over, FL_MINPOP) + ILP - 1) & ~(ILP - 1));
        }
        drain((__reduce_max_sync(0xffffffffu, cnt) + ILP - 1) & ~(ILP - 1));

        if (tmin <= 2u) {     // this lane met a pair within 3e-6 of rc2: redo its list with the oracle's decision
            LaneRedo w;
            w.pxy = pxy; w.pz = pz; w.ptyp = ptyp; w.ljt = ljt; w.cs = cs; w.gbase = gbase; w.recipe = nullptr;
            w.ncs = ncs; w.ntypes = a.ntypes; w.me = me; w.slot_i = slot_i; w.nent = nent;
            w.entries = reinterpret_cast<const uint16_t *>(lp);
            w.sx = a.sx; w.sy = a.sy; w.sz = a.sz; w.L = a.L; w.model = a.model; w.fast = a.fast; w.rc2hi = a.rc2hi;
            double f3[3];
            careful_lane<MULTI>(w, f3, &np);
            fx = f3[0]; fy = f3[1]; fz = f3[2];
        }
        if (COUNT) npair += np;

        if (active && !COUNT) { a.fx[slot_i] = fx; a.fy[slot_i] = fy; a.fz[slot_i] = fz; }     // the counting pass leaves the forces alone
        grp = ngrp;
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) npair += __shfl_xor_sync(0xffffffffu, npair, o);
        if (lane == 0 && npair) atomicAdd(a.digest, npair);
    }
}
