// list_build.cuh -- builds the pair list (Verlet list with skin) that k_force_list walks.
//
// Runs once per re-binning.  The direction find_action_partners1! was heading (src/cells.jl:224-297: for every
// atom, the partners inside its 27 / 125 neighbouring cells), restated for the GPU: one block per home brick,
// the brick and its halo staged in shared memory in exactly the order k_force_list stages them, so a stored
// entry (staged index + 1, 16 bits) means the same atom in both kernels.
//
// The kernel is a filter (no force evaluation) and it is issue-bound, so everything is organised around the
// instruction count per candidate (ncu: 28 per candidate in the first version, 7 here):
//   * candidates are staged in FP16, two per 16-byte record {x0,x1, y0,y1, z0,z1, -,-}: one broadcast LDS.128
//     and seven packed-half instructions test TWO candidates against (rc + skin)^2 plus a bound on the FP16
//     rounding (conservative: a listed pair may be outside, a pair inside is never missed);
//   * warp task = 32 consecutive home atoms of one home row, one per lane; all lanes walk the same window (the
//     (2R+1)^2 staged rows around the task's row, cells [first-R, last+R]), control flow is warp-uniform;
//   * an accepted candidate costs two predicated instructions: a 2-byte store to the lane's row in shared
//     memory and the pointer increment; rows are flushed to global memory as 16-byte chunks (LDS.128 + STG.128
//     per eight entries) whenever one of them is half full.  Exclusions are applied during the flush.
//   list8[((brick*gmax + h/32)*lcap8 + chunk)*32 + h%32],  h = index of the home atom in the brick's home list.
#pragma once
#include <cuda_fp16.h>

#include <type_traits>

#include "force_cells.cuh"

#define LB_MAX_BLOCK 256
#ifndef LB_MIN_BLOCKS
#define LB_MIN_BLOCKS 4
#endif
#define LB_ROW_ENTRIES 48                         // capacity of a lane's row
#define LB_ROW_BYTES (LB_ROW_ENTRIES * 2 + 16)    // 112 B = 28 words: eight distinct banks over the lanes
#define LB_FLUSH_AT 28                            // flush when a lane holds this many entries (up to 18 may arrive before the next check)

__host__ __device__ inline size_t lb_smem_bytes(int cap, int ncs_max, int block, bool excl)
{
    size_t b = (size_t)(cap / 2 + 20) * sizeof(uint4);     // FP16 pair records (+ slack: the unrolled scan reads up to 16 pairs past the window)
    if (excl) b += (size_t)(cap + (cap & 1) + 2) * sizeof(int);   // global id of every staged atom
    b += FC_DIMTAB * sizeof(double);
    b += (size_t)(ncs_max + 1) * sizeof(int) * 3;          // cs[], gbase[], ccoord[]
    b += 2 * (FC_MAX_HOMEROWS + 1) * sizeof(int);          // hstart[], tstart[]
    b += 8 * sizeof(int);
    b += (size_t)(ncs_max + 1) * sizeof(int);              // cfull[]: atoms of every staged cell before compaction
    b = (b + 15) & ~(size_t)15;
    b += (size_t)block * LB_ROW_BYTES;
    return b;
}

// ---- compacted staging (CellArgs::compact) ----------------------------------------------------------------------------------
// Position of the atom in global slot `slot` of staged cell t in the brick's frame (the arithmetic of stage_atoms), and whether
// it lies within rc + skin of the brick's home box: only such atoms can be a partner of a home atom.
struct BrickFrame {
    double bcx, bcy, bcz;       // scaled centre of the home box
    double hx, hy, hz;          // half extents of the home box (length units)
};
__device__ __forceinline__ BrickFrame brick_frame(const CellArgs &a, const BrickGeom &bg)
{
    const GridDesc &g = a.g;
    const int uz0 = (g.zwrap ? bg.hz0 : g.zglob0 + bg.hz0);
    BrickFrame f;
    f.bcx = ((double)bg.hx0 + 0.5 * bg.nhx) / g.M; f.bcy = ((double)bg.hy0 + 0.5 * bg.nhy) / g.M; f.bcz = ((double)uz0 + 0.5 * bg.nhz) / g.M;
    f.hx = 0.5 * bg.nhx * a.cell_edge; f.hy = 0.5 * bg.nhy * a.cell_edge; f.hz = 0.5 * bg.nhz * a.cell_edge;
    return f;
}
__device__ __forceinline__ bool staged_keep(const CellArgs &a, const BrickFrame &f, const double *ctab, int cc, int slot, double &px, double &py, double &pz)
{
    const double cx = ctab[cc & 255], cy = ctab[32 + ((cc >> 8) & 255)], cz = ctab[64 + (cc >> 16)];
    double dx = a.sx[slot] - cx, dy = a.sy[slot] - cy, dz = a.sz[slot] - cz;
    dx -= rint_magic(dx); dy -= rint_magic(dy); dz -= rint_magic(dz);
    px = a.L * (dx + (cx - f.bcx)); py = a.L * (dy + (cy - f.bcy)); pz = a.L * (dz + (cz - f.bcz));
    const double ex = fmax(fabs(px) - f.hx, 0.0), ey = fmax(fabs(py) - f.hy, 0.0), ez = fmax(fabs(pz) - f.hz, 0.0);
    return fma(ez, ez, fma(ey, ey, ex * ex)) <= a.keep2;
}
// atoms of staged cell t that are kept (one warp; every lane returns the count)
__device__ __forceinline__ int count_kept(const CellArgs &a, const BrickFrame &f, const double *ctab, int cc, int slot0, int n, int lane)
{
    int cnt = 0;
    for (int k0 = 0; k0 < n; k0 += 32) {
        double px, py, pz;
        const bool keep = k0 + lane < n && staged_keep(a, f, ctab, cc, slot0 + k0 + lane, px, py, pz);
        cnt += __popc(__ballot_sync(0xffffffffu, keep));
    }
    return cnt;
}
// largest number of kept atoms over all bricks (one block per brick): the staged-atom capacity of the compacted list kernels
__global__ void __launch_bounds__(256) k_brick_keep_max(CellArgs a, int *out)
{
    __shared__ int cs[FC_MAX_NCS_SMALL], gbase[FC_MAX_NCS_SMALL], ccoord[FC_MAX_NCS_SMALL];
    __shared__ double ctab[FC_DIMTAB];
    __shared__ int total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const BrickGeom bg = brick_geom(a.g, (int)blockIdx.x);
    if (bg.ncs > FC_MAX_NCS_SMALL) { if (tid == 0) atomicMax(out, 1 << 30); return; }      // (not a one-cell-class brick: no compaction)
    if (tid == 0) total = 0;
    stage_cell_table(a, bg, cs, gbase, ccoord, ctab, tid, (int)blockDim.x);
    __syncthreads();
    const BrickFrame f = brick_frame(a, bg);
    int mine = 0;
    for (int t = warp; t < bg.ncs; t += (int)blockDim.x >> 5) mine += count_kept(a, f, ctab, ccoord[t], gbase[t], cs[t], lane);
    if (lane == 0) atomicAdd(&total, mine);
    __syncthreads();
    if (tid == 0) atomicMax(out, total);
}

// N3 (Newton's third law inside the brick): a pair of two HOME atoms of the brick is listed once, by the atom with the smaller
// staged index (the force kernel adds the reaction to the partner's accumulator in shared memory); pairs with a halo atom stay
// listed from the home side as before (the brick that owns the halo atom evaluates its side itself).  The recipe then also
// carries each staged atom's home index + 1 (bits 21..31 of the cell code; 0: a halo atom).
// DENSE: the instantiations that know about compacted staging and split lists (CellArgs::compact, ::split); the others keep
// exactly the code (and the registers) of the plain kernel.
template <bool EXCL, bool N3 = false, bool DENSE = false>
__global__ void __launch_bounds__(LB_MAX_BLOCK, LB_MIN_BLOCKS) k_list_build(CellArgs a)
{
    static_assert(!(DENSE && N3), "dense-cell lists are full-neighbour lists");
    const int SPL = DENSE ? a.split : 0;
    const bool CMP = DENSE && a.compact != 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GridDesc &g = a.g;
    const int cap = a.cap;
    const int npairrec = cap / 2 + 20;
    uint4 *hp = reinterpret_cast<uint4 *>(smem_raw);           // FP16 coordinates of staged atoms 2q, 2q+1 in the brick's frame
    int *pid = reinterpret_cast<int *>(hp + npairrec);
    double *ctab = reinterpret_cast<double *>(pid + (EXCL ? cap + (cap & 1) + 2 : 0));
    int *cs = reinterpret_cast<int *>(ctab + FC_DIMTAB);
    int *gbase = cs + (a.ncs_max + 1);
    int *ccoord = gbase + (a.ncs_max + 1);
    int *hstart = ccoord + (a.ncs_max + 1);
    int *tstart = hstart + (FC_MAX_HOMEROWS + 1);
    int *scal = tstart + (FC_MAX_HOMEROWS + 1);
    int *cfull = scal + 8;
    unsigned char *rows = smem_raw + ((reinterpret_cast<unsigned char *>(cfull + (a.ncs_max + 1)) - smem_raw + 15) & ~(size_t)15);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = g.R;
    const int bid = FC_BRICK_OF(a, (int)blockIdx.x);
    const BrickGeom bg = brick_geom(g, bid);
    const int nhx = bg.nhx, nhy = bg.nhy, nhz = bg.nhz;
    const int sxn = bg.sxn, syn = bg.syn, ncs = bg.ncs;

    stage_cell_table(a, bg, cs, gbase, ccoord, ctab, tid, (int)blockDim.x);
    {   // records past the last staged atom are read by the unrolled scan: make them far away
        const __half2 far = __floats2half2_rn(60000.0f, 60000.0f);
        const unsigned fu = *reinterpret_cast<const unsigned *>(&far);
        for (int q = tid; q < npairrec; q += blockDim.x) hp[q] = make_uint4(fu, fu, fu, 0u);
    }
    __syncthreads();
    if (CMP) {
        const BrickFrame frame = brick_frame(a, bg);
        // compaction, pass A: cs[t] = atoms of staged cell t within rc + skin of the home box (cfull[t] keeps the cell's population)
        for (int t = warp; t < ncs; t += (int)blockDim.x >> 5) {
            const int n = cs[t];
            const int kept = count_kept(a, frame, ctab, ccoord[t], gbase[t], n, lane);
            if (lane == 0) { cfull[t] = n; cs[t] = kept; }
        }
        __syncthreads();
    }
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

    __half *hph = reinterpret_cast<__half *>(hp);
    int2 *recipe = a.recipe + (size_t)bid * a.rcap;
    auto store_atom = [&](int idx, int slot, int ccode, double px, double py, double pz) {
        __half *rec = hph + (idx >> 1) * 8 + (idx & 1);
        rec[0] = __float2half_rn((float)px);
        rec[2] = __float2half_rn((float)py);
        rec[4] = __float2half_rn((float)pz);
        if (EXCL) pid[idx] = a.id[slot];
        // the staging recipe of this brick: slot and staged-cell coordinates of every staged atom
        int code = ccode;
        if (N3) {
            const int cx = ccode & 255, cy = (ccode >> 8) & 255, cz = ccode >> 16;
            if (cx >= R && cx < R + nhx && cy >= R && cy < R + nhy && cz >= R && cz < R + nhz) {
                const int hrow_ = cz * syn + cy;
                code |= (hstart[(cz - R) * nhy + (cy - R)] + (idx - cs[hrow_ * sxn + R]) + 1) << 21;
            }
        }
        if (idx + 1 < a.rcap) recipe[idx + 1] = make_int2(slot, code);
    };
    if (CMP) {
        const BrickFrame frame = brick_frame(a, bg);
        // compaction, pass B: the kept atoms of every staged cell in their (cell, id) order, numbered from the cell's new prefix
        for (int t = warp; t < ncs; t += (int)blockDim.x >> 5) {
            const int n = cfull[t], cc = ccoord[t];
            int run = cs[t];
            for (int k0 = 0; k0 < n; k0 += 32) {
                double px = 0, py = 0, pz = 0;
                const int slot = gbase[t] + k0 + lane;
                const bool keep = k0 + lane < n && staged_keep(a, frame, ctab, cc, slot, px, py, pz);
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                const int idx = run + __popc(m & ((1u << lane) - 1u));
                if (keep && idx < nstaged) store_atom(idx, slot, cc, px, py, pz);
                run += __popc(m);
            }
        }
    } else
        stage_atoms(a, bg, cs, gbase, ccoord, ctab, 0, nstaged, tid, (int)blockDim.x, store_atom);
    if (a.seg) {
        // segment table for the TMA staging of k_force_list_p: every staged (y, z) row is one contiguous slot range, two at the
        // periodic seam in x; bulk copies need 16-byte alignment, so a segment starts at the even slot at or below its first
        // atom and covers an even number of slots.  One thread per raw group lays the group's segments out in the raw arrays.
        const int nrows = bg.nrows, ngroups = (nrows + a.raw_rows - 1) / a.raw_rows;
        int4 *seg = a.seg + (size_t)bid * a.segcap;
        for (int gi = tid; gi < ngroups; gi += blockDim.x) {
            int off = 0;
            for (int row = gi * a.raw_rows; row < min(nrows, (gi + 1) * a.raw_rows); row++) {
                int seam = sxn;                                   // first staged cell behind the seam (global x = 0), if any
                for (int cx = 1; cx < sxn; cx++)
                    if (wrap_mod(bg.hx0 - R + cx, g.M) == 0) { seam = cx; break; }
                const int cy = row % syn, cz = row / syn;
                for (int part = 0; part < 2; part++) {
                    const int c0 = part ? seam : 0, c1 = part ? sxn : seam;
                    int4 e = make_int4(0, 0, 0, 0);
                    if (c0 < c1) {
                        const int t0 = row * sxn + c0;
                        const int n = cs[row * sxn + c1] - cs[t0], g0 = gbase[t0], mis = g0 & 1;
                        if (n > 0 && 2 * row + part < a.segcap) {
                            e = make_int4(g0 - mis, n | (mis << 16), (cs[t0] + 1) | (off << 16), c0 | ((c1 - c0) << 8) | (cy << 16) | (cz << 24));
                            off += (mis + n + 1) & ~1;
                        }
                    }
                    if (2 * row + part < a.segcap) seg[2 * row + part] = e;
                }
            }
            if (off > a.rawlen) atomicCAS(a.err, 0, 2);           // (the host sized the ring from the longest row of any brick)
        }
    }
    {   // header and home list (home atom h -> staged index + 1)
        const int nh = hstart[nhy * nhz];
        if (tid == 0) { a.brickhdr[2 * bid] = nstaged + 1; a.brickhdr[2 * bid + 1] = nh; }
        // (split = 1: every home atom is two consecutive "virtual" home atoms 2h, 2h + 1 -- two lanes of the stepping kernel --
        // that share the atom's list chunk by chunk)
        for (int h = tid; h < nh && (((h << SPL) + SPL) >> 5) < a.gmax; h += blockDim.x) {
            int hr = 0;
            while (hstart[hr + 1] <= h) hr++;
            const int hrow = (hr / nhy + R) * syn + (hr % nhy + R);
            const uint16_t st = (uint16_t)(cs[hrow * sxn + R] + (h - hstart[hr]) + 1);
            for (int par = 0; par <= SPL; par++) {
                const int hv = (h << SPL) + par;
                a.homeidx[((size_t)bid * a.gmax + (hv >> 5)) * 32 + (hv & 31)] = st;
            }
        }
    }
    __syncthreads();

    const __half thr1 = __float2half_ru(a.rl2h);
    const __half2 thr = __halves2half2(thr1, thr1);
    const __half2 thr_lo_off = __halves2half2(__float2half(-1.0f), thr1);   // low candidate of the pair is outside the window
    const __half2 thr_hi_off = __halves2half2(thr1, __float2half(-1.0f));   // high candidate is outside
    const int nwin = 2 * R + 1;
    unsigned char *myrow = rows + (size_t)tid * LB_ROW_BYTES;

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
        const int h0own = cs[hrow * sxn + R];          // first home atom of the task's own row (N3)
        // own coordinates, duplicated in both halves; an inactive lane sits far away and accepts nothing
        const __half *mrec = hph + (me >> 1) * 8 + (me & 1);
        const __half hfar = __float2half(-60000.0f);
        const __half2 ix = __half2half2(active ? mrec[0] : hfar), iy = __half2half2(mrec[2]), iz = __half2half2(mrec[4]);
        int32_t xb = 0; uint64_t xm = 0;
        if (EXCL) {
            int cxi = cxa;
            while (cs[hrow * sxn + cxi + 1] <= me) cxi++;
            const int slot_i = gbase[hrow * sxn + cxi] + (me - cs[hrow * sxn + cxi]);
            xb = a.xbase[slot_i]; xm = a.xmask[slot_i];
        }
        const int h = hstart[hr] + (me - cs[hrow * sxn + R]);
        EMDEE_CHECK(hr < nhy * nhz && me < nstaged && a0 >= cs[hrow * sxn + R] && cxb < sxn, a.err);
        const int hv0 = h << SPL, hv1 = hv0 + SPL;       // the lane's virtual home atom(s)
        if ((hv1 >> 5) >= a.gmax) { atomicCAS(a.err, 0, 5); break; }
        const size_t gs = (size_t)bid * a.gmax + (hv0 >> 5), gs1 = (size_t)bid * a.gmax + (hv1 >> 5);
        uint4 *lp = a.list8 + gs * a.lcap8 * 32 + (hv0 & 31);
        uint4 *lp1 = a.list8 + gs1 * a.lcap8 * 32 + (hv1 & 31);  // split: odd chunks of the atom's list go to the second half
        int nchunks = 0;                    // chunks already in global memory
        // the lane's row is addressed by its 32-bit shared-memory address: a push is one st.shared.u16 and one add.
        // The compiler does not see these stores as memory accesses, which lets it batch the candidate loads of an
        // unrolled block; the flush (the only reader of the row) is fenced explicitly.
        const unsigned row_sh = (unsigned)__cvta_generic_to_shared(myrow);
        unsigned wsh = row_sh;              // next free entry

        // entries in the row -> global chunks; keep = false also writes the last partial chunk (unused entries = dummy)
        auto flush = [&](bool keep) {
            asm volatile("" ::: "memory");
            __syncwarp();
            const int n = (int)(wsh - row_sh) >> 1;
            const int nfull = n >> 3, rem = n & 7;
            const int nout = keep ? nfull : nfull + (rem ? 1 : 0);
            const int nmax = __reduce_max_sync(0xffffffffu, nout);
            for (int c = 0; c < nmax; c++) {
                if (c < nout) {
                    uint4 v = *reinterpret_cast<const uint4 *>(myrow + c * 16);
                    if (EXCL || c >= nfull) {
                        unsigned e[4] = {v.x, v.y, v.z, v.w};
                        const int valid = c < nfull ? 8 : rem;       // entries of this chunk that are real
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const unsigned ent = (e[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                            bool drop = k >= valid;
                            if (EXCL && !drop) drop = pair_excluded(xb, xm, pid[ent - 1]);
                            if (drop) e[k >> 1] &= (k & 1) ? 0x0000ffffu : 0xffff0000u;
                        }
                        v = make_uint4(e[0], e[1], e[2], e[3]);
                    }
                    const int cg = nchunks + c, ch = cg >> SPL;       // chunk of the atom's list, chunk of its half
                    if (ch < a.lcap8) ((cg & SPL) ? lp1 : lp)[(size_t)ch * 32] = v;
                    else atomicCAS(a.err, 0, 5);
                }
            }
            if (keep) {
                if (rem && nfull) *reinterpret_cast<uint4 *>(myrow) = *reinterpret_cast<const uint4 *>(myrow + nfull * 16);
                wsh = row_sh + 2 * rem;
            }
            nchunks += nout;
            __syncwarp();
            asm volatile("" ::: "memory");
        };
        auto push = [&](bool take, unsigned value) {
            EMDEE_CHECK(!take || ((int)(wsh - row_sh) < 2 * LB_ROW_ENTRIES && (int)value >= 1 && (int)value <= nstaged), a.err);
            asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.u16 [%0], %1; }" ::"r"(wsh), "h"((unsigned short)value), "r"((unsigned)take));
            wsh += take ? 2u : 0u;
        };
        // two candidates (staged indices 2q, 2q+1) against this lane's atom; SELF: the row holds this lane's own atom
        auto test2 = [&](uint4 c, int q, __half2 th, auto SELF) {
            const __half2 dx = __hsub2(*reinterpret_cast<const __half2 *>(&c.x), ix);
            const __half2 dy = __hsub2(*reinterpret_cast<const __half2 *>(&c.y), iy);
            const __half2 dz = __hsub2(*reinterpret_cast<const __half2 *>(&c.z), iz);
            const __half2 r2 = __hfma2(dz, dz, __hfma2(dy, dy, __hmul2(dx, dx)));
            bool lo = __hle(__low2half(r2), __low2half(th)), hi = __hle(__high2half(r2), __high2half(th));
            if (decltype(SELF)::value) {
                if (N3) { lo = lo && (2 * q < h0own || 2 * q > me); hi = hi && (2 * q + 1 < h0own || 2 * q + 1 > me); }
                else { lo = lo && (2 * q != me); hi = hi && (2 * q + 1 != me); }
            }
            push(lo, 2 * q + 1);
            push(hi, 2 * q + 2);
        };
        auto row_full = [&]() { return __any_sync(0xffffffffu, (int)(wsh - row_sh) >= 2 * LB_FLUSH_AT); };
        // one window row [p0, p1): the first and the last record may hold a candidate of a neighbouring window
        auto scan_row = [&](int p0, int p1, auto SELF) {
            int q = p0 >> 1;
            const int qlast = (p1 - 1) >> 1;
            if (q == qlast) {
                const bool lo_ok = 2 * q >= p0, hi_ok = 2 * q + 1 < p1;
                test2(hp[q], q, lo_ok ? (hi_ok ? thr : thr_hi_off) : thr_lo_off, SELF);
                return;
            }
            test2(hp[q], q, (p0 & 1) ? thr_lo_off : thr, SELF);
            q++;
            for (; q + 8 <= qlast; q += 8) {              // interior records, 16 candidates per iteration
                uint4 c[8];
#pragma unroll
                for (int u = 0; u < 8; u++) c[u] = hp[q + u];     // broadcast loads, issued together
#pragma unroll
                for (int u = 0; u < 8; u++) test2(c[u], q + u, thr, SELF);
                if (row_full()) flush(true);
            }
            for (; q < qlast; q++) test2(hp[q], q, thr, SELF);
            test2(hp[qlast], qlast, (p1 & 1) ? thr_hi_off : thr, SELF);
        };

        for (int rw = 0; rw < nwin * nwin; rw++) {
            const int row = (czi + rw / nwin - R) * syn + (cyi + rw % nwin - R);
            const int p0 = cs[row * sxn + cxa - R];
            const int p1 = min(cs[row * sxn + cxb + R + 1], nstaged);
            if (N3 && row < hrow) {
                // a home row in front of mine: its home atoms list me, not the other way round -- scan what lies beside them
                const int ry = row % syn, rz = row / syn;
                const bool homerow = ry >= R && ry < R + nhy && rz >= R && rz < R + nhz;
                const int h0 = homerow ? cs[row * sxn + R] : p1, h1 = homerow ? cs[row * sxn + R + nhx] : p1;
                if (p0 < min(p1, h0)) scan_row(p0, min(p1, h0), std::false_type());
                if (max(p0, h1) < p1) scan_row(max(p0, h1), p1, std::false_type());
            } else if (p0 < p1) {
                if (row == hrow) scan_row(p0, p1, std::true_type());
                else scan_row(p0, p1, std::false_type());
            }
            if (row_full()) flush(true);
        }
        {
            const int n = nchunks * 8 + ((int)(wsh - row_sh) >> 1);
            flush(false);
            if (active && !SPL) a.list_n[gs * 32 + (h & 31)] = (uint16_t)min(n, a.lcap8 * 8);
            if (active && SPL) {      // whole chunks per half (dropped and padding entries are zeros = the dummy atom)
                a.list_n[gs * 32 + (hv0 & 31)] = (uint16_t)min(((nchunks + 1) >> 1) * 8, a.lcap8 * 8);
                a.list_n[gs1 * 32 + (hv1 & 31)] = (uint16_t)min((nchunks >> 1) * 8, a.lcap8 * 8);
            }
        }
    }
}

// ---- the plain kernel ----------------------------------------------------------------------------------------------------------
// The list build of every configuration that needs neither compaction nor split lists, kept as its own function (the text of
// k_list_build before the dense-cell code went in): the filter loop is issue-bound and its register allocation decides its speed
// -- with the dense-cell branches compiled in, even switched off, the same loop ran 16 % slower (2.10 against 1.81 ms at config 3).
template <bool EXCL, bool N3 = false>
__global__ void __launch_bounds__(LB_MAX_BLOCK, LB_MIN_BLOCKS) k_list_build_plain(CellArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GridDesc &g = a.g;
    const int cap = a.cap;
    const int npairrec = cap / 2 + 20;
    uint4 *hp = reinterpret_cast<uint4 *>(smem_raw);           // FP16 coordinates of staged atoms 2q, 2q+1 in the brick's frame
    int *pid = reinterpret_cast<int *>(hp + npairrec);
    double *ctab = reinterpret_cast<double *>(pid + (EXCL ? cap + (cap & 1) + 2 : 0));
    int *cs = reinterpret_cast<int *>(ctab + FC_DIMTAB);
    int *gbase = cs + (a.ncs_max + 1);
    int *ccoord = gbase + (a.ncs_max + 1);
    int *hstart = ccoord + (a.ncs_max + 1);
    int *tstart = hstart + (FC_MAX_HOMEROWS + 1);
    int *scal = tstart + (FC_MAX_HOMEROWS + 1);
    unsigned char *rows = smem_raw + ((reinterpret_cast<unsigned char *>(scal + 8) - smem_raw + 15) & ~(size_t)15);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = g.R;
    const int bid = FC_BRICK_OF(a, (int)blockIdx.x);
    const BrickGeom bg = brick_geom(g, bid);
    const int nhx = bg.nhx, nhy = bg.nhy, nhz = bg.nhz;
    const int sxn = bg.sxn, syn = bg.syn, ncs = bg.ncs;

    stage_cell_table(a, bg, cs, gbase, ccoord, ctab, tid, (int)blockDim.x);
    {   // records past the last staged atom are read by the unrolled scan: make them far away
        const __half2 far = __floats2half2_rn(60000.0f, 60000.0f);
        const unsigned fu = *reinterpret_cast<const unsigned *>(&far);
        for (int q = tid; q < npairrec; q += blockDim.x) hp[q] = make_uint4(fu, fu, fu, 0u);
    }
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

    __half *hph = reinterpret_cast<__half *>(hp);
    int2 *recipe = a.recipe + (size_t)bid * a.rcap;
    stage_atoms(a, bg, cs, gbase, ccoord, ctab, 0, nstaged, tid, (int)blockDim.x, [&](int idx, int slot, int ccode, double px, double py, double pz) {
        __half *rec = hph + (idx >> 1) * 8 + (idx & 1);
        rec[0] = __float2half_rn((float)px);
        rec[2] = __float2half_rn((float)py);
        rec[4] = __float2half_rn((float)pz);
        if (EXCL) pid[idx] = a.id[slot];
        // the staging recipe of this brick: slot and staged-cell coordinates of every staged atom
        int code = ccode;
        if (N3) {
            const int cx = ccode & 255, cy = (ccode >> 8) & 255, cz = ccode >> 16;
            if (cx >= R && cx < R + nhx && cy >= R && cy < R + nhy && cz >= R && cz < R + nhz) {
                const int hrow_ = cz * syn + cy;
                code |= (hstart[(cz - R) * nhy + (cy - R)] + (idx - cs[hrow_ * sxn + R]) + 1) << 21;
            }
        }
        if (idx + 1 < a.rcap) recipe[idx + 1] = make_int2(slot, code);
    });
    if (a.seg) {
        // segment table for the TMA staging of k_force_list_p: every staged (y, z) row is one contiguous slot range, two at the
        // periodic seam in x; bulk copies need 16-byte alignment, so a segment starts at the even slot at or below its first
        // atom and covers an even number of slots.  One thread per raw group lays the group's segments out in the raw arrays.
        const int nrows = bg.nrows, ngroups = (nrows + a.raw_rows - 1) / a.raw_rows;
        int4 *seg = a.seg + (size_t)bid * a.segcap;
        for (int gi = tid; gi < ngroups; gi += blockDim.x) {
            int off = 0;
            for (int row = gi * a.raw_rows; row < min(nrows, (gi + 1) * a.raw_rows); row++) {
                int seam = sxn;                                   // first staged cell behind the seam (global x = 0), if any
                for (int cx = 1; cx < sxn; cx++)
                    if (wrap_mod(bg.hx0 - R + cx, g.M) == 0) { seam = cx; break; }
                const int cy = row % syn, cz = row / syn;
                for (int part = 0; part < 2; part++) {
                    const int c0 = part ? seam : 0, c1 = part ? sxn : seam;
                    int4 e = make_int4(0, 0, 0, 0);
                    if (c0 < c1) {
                        const int t0 = row * sxn + c0;
                        const int n = cs[row * sxn + c1] - cs[t0], g0 = gbase[t0], mis = g0 & 1;
                        if (n > 0 && 2 * row + part < a.segcap) {
                            e = make_int4(g0 - mis, n | (mis << 16), (cs[t0] + 1) | (off << 16), c0 | ((c1 - c0) << 8) | (cy << 16) | (cz << 24));
                            off += (mis + n + 1) & ~1;
                        }
                    }
                    if (2 * row + part < a.segcap) seg[2 * row + part] = e;
                }
            }
            if (off > a.rawlen) atomicCAS(a.err, 0, 2);           // (the host sized the ring from the longest row of any brick)
        }
    }
    {   // header and home list (home atom h -> staged index + 1)
        const int nh = hstart[nhy * nhz];
        if (tid == 0) { a.brickhdr[2 * bid] = nstaged + 1; a.brickhdr[2 * bid + 1] = nh; }
        for (int h = tid; h < nh && (h >> 5) < a.gmax; h += blockDim.x) {
            int hr = 0;
            while (hstart[hr + 1] <= h) hr++;
            const int hrow = (hr / nhy + R) * syn + (hr % nhy + R);
            a.homeidx[((size_t)bid * a.gmax + (h >> 5)) * 32 + (h & 31)] = (uint16_t)(cs[hrow * sxn + R] + (h - hstart[hr]) + 1);
        }
    }
    __syncthreads();

    const __half thr1 = __float2half_ru(a.rl2h);
    const __half2 thr = __halves2half2(thr1, thr1);
    const __half2 thr_lo_off = __halves2half2(__float2half(-1.0f), thr1);   // low candidate of the pair is outside the window
    const __half2 thr_hi_off = __halves2half2(thr1, __float2half(-1.0f));   // high candidate is outside
    const int nwin = 2 * R + 1;
    unsigned char *myrow = rows + (size_t)tid * LB_ROW_BYTES;

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
        const int h0own = cs[hrow * sxn + R];          // first home atom of the task's own row (N3)
        // own coordinates, duplicated in both halves; an inactive lane sits far away and accepts nothing
        const __half *mrec = hph + (me >> 1) * 8 + (me & 1);
        const __half hfar = __float2half(-60000.0f);
        const __half2 ix = __half2half2(active ? mrec[0] : hfar), iy = __half2half2(mrec[2]), iz = __half2half2(mrec[4]);
        int32_t xb = 0; uint64_t xm = 0;
        if (EXCL) {
            int cxi = cxa;
            while (cs[hrow * sxn + cxi + 1] <= me) cxi++;
            const int slot_i = gbase[hrow * sxn + cxi] + (me - cs[hrow * sxn + cxi]);
            xb = a.xbase[slot_i]; xm = a.xmask[slot_i];
        }
        const int h = hstart[hr] + (me - cs[hrow * sxn + R]);
        EMDEE_CHECK(hr < nhy * nhz && me < nstaged && a0 >= cs[hrow * sxn + R] && cxb < sxn, a.err);
        if ((h >> 5) >= a.gmax) { atomicCAS(a.err, 0, 5); break; }
        const size_t gs = (size_t)bid * a.gmax + (h >> 5);
        uint4 *lp = a.list8 + gs * a.lcap8 * 32 + (h & 31);
        int nchunks = 0;                    // chunks already in global memory
        // the lane's row is addressed by its 32-bit shared-memory address: a push is one st.shared.u16 and one add.
        // The compiler does not see these stores as memory accesses, which lets it batch the candidate loads of an
        // unrolled block; the flush (the only reader of the row) is fenced explicitly.
        const unsigned row_sh = (unsigned)__cvta_generic_to_shared(myrow);
        unsigned wsh = row_sh;              // next free entry

        // entries in the row -> global chunks; keep = false also writes the last partial chunk (unused entries = dummy)
        auto flush = [&](bool keep) {
            asm volatile("" ::: "memory");
            __syncwarp();
            const int n = (int)(wsh - row_sh) >> 1;
            const int nfull = n >> 3, rem = n & 7;
            const int nout = keep ? nfull : nfull + (rem ? 1 : 0);
            const int nmax = __reduce_max_sync(0xffffffffu, nout);
            for (int c = 0; c < nmax; c++) {
                if (c < nout) {
                    uint4 v = *reinterpret_cast<const uint4 *>(myrow + c * 16);
                    if (EXCL || c >= nfull) {
                        unsigned e[4] = {v.x, v.y, v.z, v.w};
                        const int valid = c < nfull ? 8 : rem;       // entries of this chunk that are real
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const unsigned ent = (e[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                            bool drop = k >= valid;
                            if (EXCL && !drop) drop = pair_excluded(xb, xm, pid[ent - 1]);
                            if (drop) e[k >> 1] &= (k & 1) ? 0x0000ffffu : 0xffff0000u;
                        }
                        v = make_uint4(e[0], e[1], e[2], e[3]);
                    }
                    if (nchunks + c < a.lcap8) lp[(size_t)(nchunks + c) * 32] = v;
                    else atomicCAS(a.err, 0, 5);
                }
            }
            if (keep) {
                if (rem && nfull) *reinterpret_cast<uint4 *>(myrow) = *reinterpret_cast<const uint4 *>(myrow + nfull * 16);
                wsh = row_sh + 2 * rem;
            }
            nchunks += nout;
            __syncwarp();
            asm volatile("" ::: "memory");
        };
        auto push = [&](bool take, unsigned value) {
            EMDEE_CHECK(!take || ((int)(wsh - row_sh) < 2 * LB_ROW_ENTRIES && (int)value >= 1 && (int)value <= nstaged), a.err);
            asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.shared.u16 [%0], %1; }" ::"r"(wsh), "h"((unsigned short)value), "r"((unsigned)take));
            wsh += take ? 2u : 0u;
        };
        // two candidates (staged indices 2q, 2q+1) against this lane's atom; SELF: the row holds this lane's own atom
        auto test2 = [&](uint4 c, int q, __half2 th, auto SELF) {
            const __half2 dx = __hsub2(*reinterpret_cast<const __half2 *>(&c.x), ix);
            const __half2 dy = __hsub2(*reinterpret_cast<const __half2 *>(&c.y), iy);
            const __half2 dz = __hsub2(*reinterpret_cast<const __half2 *>(&c.z), iz);
            const __half2 r2 = __hfma2(dz, dz, __hfma2(dy, dy, __hmul2(dx, dx)));
            bool lo = __hle(__low2half(r2), __low2half(th)), hi = __hle(__high2half(r2), __high2half(th));
            if (decltype(SELF)::value) {
                if (N3) { lo = lo && (2 * q < h0own || 2 * q > me); hi = hi && (2 * q + 1 < h0own || 2 * q + 1 > me); }
                else { lo = lo && (2 * q != me); hi = hi && (2 * q + 1 != me); }
            }
            push(lo, 2 * q + 1);
            push(hi, 2 * q + 2);
        };
        auto row_full = [&]() { return __any_sync(0xffffffffu, (int)(wsh - row_sh) >= 2 * LB_FLUSH_AT); };
        // one window row [p0, p1): the first and the last record may hold a candidate of a neighbouring window
        auto scan_row = [&](int p0, int p1, auto SELF) {
            int q = p0 >> 1;
            const int qlast = (p1 - 1) >> 1;
            if (q == qlast) {
                const bool lo_ok = 2 * q >= p0, hi_ok = 2 * q + 1 < p1;
                test2(hp[q], q, lo_ok ? (hi_ok ? thr : thr_hi_off) : thr_lo_off, SELF);
                return;
            }
            test2(hp[q], q, (p0 & 1) ? thr_lo_off : thr, SELF);
            q++;
            for (; q + 8 <= qlast; q += 8) {              // interior records, 16 candidates per iteration
                uint4 c[8];
#pragma unroll
                for (int u = 0; u < 8; u++) c[u] = hp[q + u];     // broadcast loads, issued together
#pragma unroll
                for (int u = 0; u < 8; u++) test2(c[u], q + u, thr, SELF);
                if (row_full()) flush(true);
            }
            for (; q < qlast; q++) test2(hp[q], q, thr, SELF);
            test2(hp[qlast], qlast, (p1 & 1) ? thr_hi_off : thr, SELF);
        };

        for (int rw = 0; rw < nwin * nwin; rw++) {
            const int row = (czi + rw / nwin - R) * syn + (cyi + rw % nwin - R);
            const int p0 = cs[row * sxn + cxa - R];
            const int p1 = min(cs[row * sxn + cxb + R + 1], nstaged);
            if (N3 && row < hrow) {
                // a home row in front of mine: its home atoms list me, not the other way round -- scan what lies beside them
                const int ry = row % syn, rz = row / syn;
                const bool homerow = ry >= R && ry < R + nhy && rz >= R && rz < R + nhz;
                const int h0 = homerow ? cs[row * sxn + R] : p1, h1 = homerow ? cs[row * sxn + R + nhx] : p1;
                if (p0 < min(p1, h0)) scan_row(p0, min(p1, h0), std::false_type());
                if (max(p0, h1) < p1) scan_row(max(p0, h1), p1, std::false_type());
            } else if (p0 < p1) {
                if (row == hrow) scan_row(p0, p1, std::true_type());
                else scan_row(p0, p1, std::false_type());
            }
            if (row_full()) flush(true);
        }
        {
            const int n = nchunks * 8 + ((int)(wsh - row_sh) >> 1);
            flush(false);
            if (active) a.list_n[gs * 32 + (h & 31)] = (uint16_t)min(n, a.lcap8 * 8);
        }
    }
}
