// force_list_p.cuh -- persistent, warp-specialised form of the stepping kernel (k_force_list).
//
// ncu on k_force_list (profiles/): a third of all warp time went into staging a brick (three dependent global
// round trips and two block barriers per block) during which the block's FP64 work stands still, and shared
// memory allowed only two blocks per SM, so staging was overlapped by at most one other block.  Here one
// 512-thread block per SM stays resident and walks over bricks (the first one is blockIdx.x, later ones are claimed
// from a global counter, so the SMs finish within one brick of each other):
//   * 4 producer warps stage brick k+1 into one of two shared-memory buffers from the per-brick staging recipe that
//     k_list_build wrote at the last re-binning (slot and staged-cell coordinates of every staged atom: one coalesced
//     load, three gathers and ~30 instructions per atom, no cell table / prefix scan / search on the step path; the
//     first batch of loads is issued before the buffer is handed over), while
//   * 12 consumer warps walk the pair list of brick k (same walk / drain as k_force_list, ILP 4).
// Hand-over by named barriers (bar.arrive / bar.sync): full[b] producers -> consumers, empty[b] consumers ->
// producers; no block-wide barrier after the prologue.  A consumer warp that runs out of tasks in brick k moves
// on to brick k+1 as soon as that buffer is full; the list head of its next task (also across bricks) is requested
// one task ahead.
#pragma once
#include "force_list.cuh"

#ifndef FLP_THREADS
#define FLP_THREADS 512
#endif
#ifndef FLP_ILP
#define FLP_ILP 4            // stack entries evaluated per drain iteration (independent FP64 chains per lane)
#endif
#ifndef FLP_STAGE_U
#define FLP_STAGE_U 9        // atoms per producer thread in flight: ~2300 staged atoms / 128 threads = two batches
#endif
#ifndef FLP_NPROD
#define FLP_NPROD 4
#endif
#ifndef FLP_ADV_W
#define FLP_ADV_W 3          // atoms per producer thread in flight in the fused integrator
#endif
#ifndef FLP_MBAR
#define FLP_MBAR 1           // hand-over of the staging buffers by mbarriers (0: named barriers, every consumer warp waits for the slowest)
#endif
#ifndef FLP_CARRY
#define FLP_CARRY 0          // a warp's first task of the NEXT brick is claimed (and its list head requested) before the final drain
#endif                       // of its last task in this one, when that buffer is already full
#define FLP_NCONS (FLP_THREADS / 32 - FLP_NPROD)
#define FLP_QS (FLP_NCONS * 32)          // row stride of the consumers' stacks

// n3_groups > 0 (Newton's third law inside the brick): per buffer also the home index + 1 of every staged atom (uint16) and
// the force accumulators of the home atoms (3 doubles per lane of every 32-atom group)
__host__ __device__ inline size_t flp_buf_bytes(int cap, int ncs_max, int ntypes, int n3_groups = 0)
{
    const size_t cap1 = (size_t)cap + 1;
    size_t b = cap1 * (sizeof(double2) + sizeof(double) + sizeof(uint2));
    if (ntypes > 1) b += (cap1 + 15) & ~(size_t)15;
    b = (b + 15) & ~(size_t)15;
    b += 8 * sizeof(int);                         // scal[]
    b = (b + 15) & ~(size_t)15;
    if (n3_groups > 0) {
        b += (cap1 * sizeof(uint16_t) + 15) & ~(size_t)15;
        b += (size_t)n3_groups * 32 * 3 * sizeof(double);
    }
    return b;
}
// nbuf staging buffers; the per-lane stacks shrink to 24 entries when three buffers are wanted (a brick period is
// set by its slowest warp task, a third buffer lets the other warps run ahead instead of waiting for it)
#define FLP_QCAP_SMALL 24     // stack depth the host falls back to when two buffers do not fit next to FL_QCAP entries per lane
__host__ __device__ inline size_t flp_smem_bytes(int cap, int ncs_max, int ntypes, int nbuf, int n3_groups = 0, int qcap = FL_QCAP)
{
    return nbuf * flp_buf_bytes(cap, ncs_max, ntypes, n3_groups) + (size_t)ntypes * ntypes * sizeof(double2) +
           (size_t)(qcap + 1) * FLP_QS * sizeof(uint16_t) +
           128;     // (head-room for the kernel's static shared memory -- barriers, brick claims -- which counts against the same limit)
}

// -DFLP_TIMING=1: per-role cycle counters (clock64) accumulated into a.timing[8]: producers' wait for an empty buffer,
// staging, integrator, rest; consumers' wait for a full buffer and total time (summed over consumer warps), bricks
#ifndef FLP_TIMING
#define FLP_TIMING 0
#endif
__device__ __forceinline__ long long flp_clock() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory"); return t; }
#if FLP_TIMING
// BAR.SYNC does not block at issue on sm_100 (the block is deferred to the next instruction that touches barrier-protected
// state), so a clock read right behind a barrier measures nothing: FLP_TB reads a word of shared memory first
__device__ __forceinline__ long long flp_clock_after(const volatile int *protected_word)
{
    const int v = *protected_word;
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(v) : "memory");
    return t;
}
#define FLP_T(var) const long long var = flp_clock()
#define FLP_TB(var, word) const long long var = flp_clock_after(word)
#define FLP_TACC(slot, expr) do { if (lane == 0) tacc[slot] += (expr); } while (0)
#else
#define FLP_TB(var, word)
#define FLP_T(var)
#define FLP_TACC(slot, expr)
#endif
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- bulk asynchronous copies (TMA, cp.async.bulk) completing on an mbarrier: the staging path of the TMA variant ----------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// (bounded: a copy that never completes -- a fault, a byte count that does not add up -- raises device flag 9 instead of hanging)
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity, int *err)
{
    unsigned done;
    int spins = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && ++spins > (1 << 20)) { atomicCAS(err, 0, 9); break; }
    } while (!done);
}
// ---- hand-over of the staging buffers by mbarriers: full[b] counts the producer threads, empty[b] the consumer threads -------
// (named barriers made every consumer warp wait for the slowest one at each brick boundary: bar.sync needs all of them to arrive;
// with an mbarrier a warp that is done with brick k goes on to brick k+1 as soon as that buffer is full, one brick of slack)
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded by time (~4 s: boundary bricks of a slab may wait ~2 s for a neighbour's flag before they are staged): device flag 9
__device__ __forceinline__ bool mbar_wait_long(unsigned long long *bar, unsigned parity, int *err)
{
    unsigned done;
    int spins = 0;
    long long t0 = 0;
    for (;;) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
        if ((++spins & 1023) == 0) {
            long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
            if (t0 == 0) t0 = t;
            else if (t - t0 > (1ll << 33)) { atomicCAS(err, 0, 9); return false; }
        }
    }
}
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned parity)      // has that phase completed? (no waiting)
{
    unsigned done;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
// `bytes` (a multiple of 16) from 16-byte aligned global memory to 16-byte aligned shared memory; completion is counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// shared memory of the TMA variant behind the stacks: two segment tables (this brick's, the next one's) and a ring of two raw
// row groups (three coordinate arrays each) that the bulk copies fill
__host__ __device__ inline size_t flp_tma_bytes(int segcap, int rawlen)
{
    return 2 * (size_t)segcap * sizeof(int4) + 2 * 3 * (size_t)rawlen * sizeof(double) + 32;
}

struct BrickBuf {
    double2 *pxy;
    double *pz;
    uint2 *ph;
    uint8_t *ptyp;
    int *scal;            // [0] home atoms, [1] staged atoms + 1, [3] task cursor, [4] brick (-1: done)
    uint16_t *phome;      // N3: home index + 1 of a staged atom (0: halo atom)
    double *acc;          // N3: force accumulators of the home atoms, acc[3 * home index + c]
};
__device__ __forceinline__ BrickBuf brick_buf(unsigned char *base, int cap, int ncs_max, bool multi)
{
    const size_t cap1 = (size_t)cap + 1;
    BrickBuf b;
    b.pxy = reinterpret_cast<double2 *>(base);
    b.pz = reinterpret_cast<double *>(b.pxy + cap1);
    b.ph = reinterpret_cast<uint2 *>(b.pz + cap1);
    b.ptyp = reinterpret_cast<uint8_t *>(b.ph + cap1);
    size_t off = (size_t)(reinterpret_cast<unsigned char *>(b.ptyp) - base);
    if (multi) off += (cap1 + 15) & ~(size_t)15;
    off = (off + 15) & ~(size_t)15;
    b.scal = reinterpret_cast<int *>(base + off);
    off = (off + 8 * sizeof(int) + 15) & ~(size_t)15;
    b.phome = reinterpret_cast<uint16_t *>(base + off);
    b.acc = reinterpret_cast<double *>(base + off + ((cap1 * sizeof(uint16_t) + 15) & ~(size_t)15));
    return b;
}

// EW: also per-atom energies and virials (half of every pair to each atom, src/nonbonded.jl:93-94) -- the single-point
// evaluation behind emdee_compute_nonbonded(EMDEE_CUTOFF); store_f: write the forces (bitmask without FORCES: false).
// FUSE: from the second chunk on, every walk iteration also pops and evaluates ILP stack entries in the same basic
// block, so the scheduler fills the FP64 dependency stalls of the drain with the FP16 tests and stack pushes of the
// walk (ncu: `wait` was the top stall of the consumers with walk and drain as separate loops).
// VV: the epilogue also advances the atom (k_vv's arithmetic, one fma per line, same rounding): v += h f [step n done];
// v += h f; r += dt v; s' = r/L [step n+1 started], so a step is ONE kernel instead of k_vv + force kernel.
#ifdef FLP_MAXNREG      // block sizes that are not a multiple of 128: ptxas rounds the launch bound up, this states the budget
#define FLP_BOUNDS __maxnreg__(FLP_MAXNREG)
#else
#define FLP_BOUNDS __launch_bounds__(FLP_THREADS, 1)
#endif
// P2P: slab decomposition with peer-mapped halos (see CellArgs): ghost positions are written by the neighbouring GPUs during
// the launch, so staged coordinates are read with L2-only loads, boundary bricks wait for the neighbours' flags, and the
// integrator pushes my boundary atoms to the neighbours.
// N3: the list holds every pair of two home atoms once (k_list_build<.., N3>); the evaluating lane adds the reaction to the
// partner's accumulator in shared memory (FP64 compare-and-swap loops), home atoms' own sums go to the same accumulators at the
// end of a warp task, and the producers write a brick's forces out once its consumers have released the buffer.
// TMA: the producers stage a brick from bulk asynchronous copies instead of per-atom gathers.  After the (cell, id) sort every
// (y, z) row of the brick's cell block is one contiguous slot range of the coordinate arrays (two at the periodic seam), so a
// row is three cp.async.bulk copies (x, y, z) that one lane issues and an mbarrier counts; k_list_build leaves the segment
// table of every brick behind.  Row groups are copied into a two-slot ring of raw scaled coordinates ahead of time -- also
// across bricks, before the consumers have released the buffer -- and what remains on the hand-over's critical path is the
// shared-to-shared pass into the brick's frame (FP64 + FP16 copies).  No staging recipe is read.
// LM (two-level list, GROMACS calls it dynamic pruning): the list built at a re-binning holds every pair inside rc + skin (~85
// entries per atom at skin 0.45, 54 of them inside rc), and walking it costs an FP16 test per entry on every step.  LM = 1 (a
// "prune" step) walks it with the FP16 threshold of rc + skin2 (skin2 << skin) and logs every stack entry it drains -- the
// survivors, in drain order, padded with dummies where a lane's stack ran dry -- as the INNER list of the warp task: the same
// number of rows for every lane, so it is written as coalesced 16-byte chunks.  LM = 2 replays that log on the following steps:
// no FP16 test, no stack, one gather and the pair arithmetic per row.  The host switches back to LM = 1 before an atom can have
// moved skin2 / 2 since the prune step (emdee_vv_step), so the evaluated pair set stays the oracle's on every step.
// DENSE (dense cells, see choose_bricks): the per-lane stack depth and the split of an atom's list over two lanes are run-time
// arguments (a.qcap, a.split); as compile-time constants in the other instantiations they cost the hot loops no registers.
template <bool MULTI, bool COUNT, int NBUF, bool EW, bool FUSE, bool VV = false, bool P2P = false, bool N3 = false, bool TMA = false, int LM = 0,
          bool DENSE = false>
__global__ void FLP_BOUNDS k_force_list_p(CellArgs a, int nbricks, int store_f)
{
    constexpr int ILP = FLP_ILP;
    static_assert(!DENSE || (!P2P && !N3 && !TMA && LM == 0), "the dense-cell variants exist for the plain single-GPU kernels");
    const int SPL = DENSE ? a.split : 0;
    static_assert(LM == 0 || (ILP == 4 && FUSE && !N3 && !EW && (LM == 2 || !COUNT)), "the two-level list is a variant of the fused stepping kernel");
    const int QCAP = DENSE ? a.qcap : FL_QCAP;      // per-lane stack entries (FLP_QCAP_SMALL where shared memory is short)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GridDesc &g = a.g;
    const int cap1 = a.cap + 1;
    const size_t bufsz = flp_buf_bytes(a.cap, a.ncs_max, a.ntypes, N3 ? a.gmax : 0);
    double2 *ljt = reinterpret_cast<double2 *>(smem_raw + NBUF * bufsz);
    uint16_t *qguard = reinterpret_cast<uint16_t *>(ljt + a.ntypes * a.ntypes);
    uint16_t *queue = qguard + FLP_QS;
    __shared__ int claimed[2];                 // the producers' next brick (double-buffered over the brick parity)
    __shared__ __align__(8) unsigned long long tma_bar[2];      // one mbarrier per slot of the raw ring (TMA variant)
    __shared__ __align__(8) unsigned long long hand_bar[2 * NBUF];   // full[b] = hand_bar[b], empty[b] = hand_bar[NBUF + b]
    // TMA variant: segment tables and raw ring behind the stacks
    int4 *segs = reinterpret_cast<int4 *>(
        smem_raw + ((reinterpret_cast<unsigned char *>(queue + (size_t)QCAP * FLP_QS) - smem_raw + 15) & ~(size_t)15));
    double *rawring = reinterpret_cast<double *>(segs + 2 * (TMA ? a.segcap : 0));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        if (TMA) { mbar_init(&tma_bar[0], FLP_NPROD); mbar_init(&tma_bar[1], FLP_NPROD); }
        for (int q = 0; q < NBUF; q++) { mbar_init(&hand_bar[q], FLP_NPROD * 32); mbar_init(&hand_bar[NBUF + q], FLP_NCONS * 32); }
        mbar_fence_init();
    }
    // fill number n of a buffer (n = k / NBUF): the consumers wait for phase n of full[b], the producers for phase n of empty[b]
    // before fill n + 1; every thread of the arriving side arrives (release), every waiting thread polls (acquire)
    auto full_arrive = [&](int b_) { if (FLP_MBAR) mbar_arrive(&hand_bar[b_]); else bar_arrive(1 + b_, FLP_THREADS); };
    auto full_wait = [&](int b_, int fill) { return FLP_MBAR ? mbar_wait_long(&hand_bar[b_], fill & 1, a.err) : (bar_sync(1 + b_, FLP_THREADS), true); };
    auto empty_arrive = [&](int b_) { if (FLP_MBAR) mbar_arrive(&hand_bar[NBUF + b_]); else bar_arrive(1 + NBUF + b_, FLP_THREADS); };
    auto empty_wait = [&](int b_, int fill) { return FLP_MBAR ? mbar_wait_long(&hand_bar[NBUF + b_], fill & 1, a.err) : (bar_sync(1 + NBUF + b_, FLP_THREADS), true); };
    for (int t = tid; t < a.ntypes * a.ntypes; t += FLP_THREADS) ljt[t] = a.ljtab[t];
    if (tid < FLP_QS) qguard[tid] = 0;
    if (N3)
        for (int q = 0; q < NBUF; q++) {
            double *acc0 = brick_buf(smem_raw + q * bufsz, a.cap, a.ncs_max, MULTI).acc;
            for (int t = tid; t < a.gmax * 96; t += FLP_THREADS) acc0[t] = 0.0;
        }
    __syncthreads();
#if FLP_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = flp_clock();
#endif

    if (warp < FLP_NPROD) {
        // ================================ producers ================================
        constexpr int PN = FLP_NPROD * 32;
        const int R = g.R;
        int hdr_n1 = 0, hdr_nh = 0;
        if ((int)blockIdx.x < nbricks) {
            hdr_n1 = a.brickhdr[2 * FC_BRICK_OF(a, (int)blockIdx.x)];
            hdr_nh = a.brickhdr[2 * FC_BRICK_OF(a, (int)blockIdx.x) + 1];
        }
        int brick = blockIdx.x;
        // VV: once the consumers have released a buffer, the forces of that brick's home atoms are final; the producers
        // (idle ~40 % of the time) then complete the step for those atoms (second half-kick) and start the next one
        // (first half-kick, drift, s = r/L into the second position buffer) -- k_vv's arithmetic, one fma per line.
        int held_bid[NBUF], held_nh[NBUF];           // brick in each buffer whose atoms still have to be advanced
#pragma unroll
        for (int q = 0; q < NBUF; q++) { held_bid[q] = -1; held_nh[q] = 0; }
        auto advance_atoms = [&](int vbid, int vnh) {
            const int2 *vrecipe = a.recipe + (size_t)vbid * a.rcap;
            unsigned dmax = 0, smax = 0;
            const bool push = VV && P2P && a.publish_epoch != 0 && a.vv_mode == 2;
            const long long peer_lo_first = push ? a.peer_info[0] : 0;       // the lower neighbour's first upper-ghost slot
            constexpr int W = FLP_ADV_W;         // atoms per thread in flight: the chain home index -> slot -> data is pure latency
            for (int h00 = 0; h00 < vnh; h00 += W * PN) {
                int slot[W];
                bool ok[W];
                double mass[W], f3[W][3], v3[W][3], r3[W][3], rb3[W][3];
#pragma unroll
                for (int u = 0; u < W; u++) {
                    const int h = h00 + u * PN + tid, hv = h << SPL;      // (split lists: home atom h is the virtual atoms 2h, 2h + 1)
                    ok[u] = h < vnh && (hv >> 5) < a.gmax;
                    slot[u] = ok[u] ? (int)a.homeidx[((size_t)vbid * a.gmax + (hv >> 5)) * 32 + (hv & 31)] : 1;
                }
#pragma unroll
                for (int u = 0; u < W; u++) slot[u] = vrecipe[slot[u]].x;
#pragma unroll
                for (int u = 0; u < W; u++) {
                    mass[u] = a.vv_mass[slot[u]];
                    f3[u][0] = a.fx[slot[u]]; f3[u][1] = a.fy[slot[u]]; f3[u][2] = a.fz[slot[u]];
#pragma unroll
                    for (int c3 = 0; c3 < 3; c3++) { v3[u][c3] = a.vv_v[c3][slot[u]]; r3[u][c3] = a.vv_r[c3][slot[u]]; rb3[u][c3] = a.vv_rb[c3][slot[u]]; }
                }
#pragma unroll
                for (int u = 0; u < W; u++) {
                    if (!ok[u]) continue;
                    const double hk = __ddiv_rn(0.5 * a.vv_dt, mass[u]);
                    double d2 = 0, v2 = 0;
#pragma unroll
                    for (int c3 = 0; c3 < 3; c3++) {
                        double v = __fma_rn(hk, f3[u][c3], v3[u][c3]);                // second half-kick of this step
                        if (a.vv_mode == 2) {
                            v = __fma_rn(hk, f3[u][c3], v);                           // first half-kick of the next step
                            const double r = __fma_rn(a.vv_dt, v, r3[u][c3]);         // drift
                            a.vv_r[c3][slot[u]] = r;
                            const double sn = __ddiv_rn(r, a.L);
                            a.vv_snew[c3][slot[u]] = sn;
                            if (push) {      // my boundary planes are the neighbours' ghosts: same order, one contiguous range per side
                                const unsigned ol = (unsigned)(slot[u] - a.lo_send_a), oh = (unsigned)(slot[u] - a.hi_send_a);
                                if (ol < (unsigned)a.lo_send_n) a.peer_lo[c3][peer_lo_first + ol] = sn;
                                if (oh < (unsigned)a.hi_send_n) a.peer_hi[c3][oh] = sn;
                            }
                            const double d = r - rb3[u][c3];
                            d2 = fma(d, d, d2);
                            v2 = fma(v, v, v2);
                        }
                        a.vv_v[c3][slot[u]] = v;
                    }
                    if (a.vv_mode == 2 && a.vv_check_skin && d2 > a.vv_half_skin2) atomicCAS(a.err, 0, 3);
                    dmax = max(dmax, __float_as_uint(__double2float_ru(d2)));
                    smax = max(smax, __float_as_uint(__double2float_ru(v2 * a.vv_dt * a.vv_dt)));
                }
            }
            if (a.vv_mode == 2 && a.vv_maxd2) {
                dmax = __reduce_max_sync(0xffffffffu, dmax);
                if (lane == 0 && dmax > *a.vv_maxd2) atomicMax(a.vv_maxd2, dmax);
            }
            if (a.vv_mode == 2 && a.vv_maxstep) {      // this step's drift: what the two-level list's prune cadence is decided on
                smax = __reduce_max_sync(0xffffffffu, smax);
                if (lane == 0 && smax > *a.vv_maxstep) atomicMax(a.vv_maxstep, smax);
            }
            if (push) {
                // publish: when every brick holding atoms a neighbour needs has been advanced, raise that neighbour's flag
                // (stores to peer memory -> system-scope fence by every writer -> producers' barrier -> counter -> flag)
                const int bzi = vbid / (g.nbx * g.nby);
                const bool in_lo = bzi < a.p2p_lo_layers, in_hi = bzi >= a.p2p_hi_layer0;
                if (in_lo || in_hi) {
                    __threadfence_system();
                    bar_sync(1 + 2 * NBUF, PN);
                    if (tid == 0) {
                        if (in_lo && atomicAdd(&a.p2p_done[0], 1) + 1 == a.p2p_nlo) { __threadfence_system(); st_release_sys(a.flag_lo_peer, a.publish_epoch); }
                        if (in_hi && atomicAdd(&a.p2p_done[1], 1) + 1 == a.p2p_nhi) { __threadfence_system(); st_release_sys(a.flag_hi_peer, a.publish_epoch); }
                    }
                }
            }
        };
        // N3: the forces of a released brick's home atoms leave shared memory (and the accumulators are cleared for the next brick);
        // thread tid handles the same home atoms as in advance_atoms, which reads these forces back from global memory
        auto flush_forces = [&](const BrickBuf &Bf, int vbid, int vnh) {
            const int2 *vrecipe = a.recipe + (size_t)vbid * a.rcap;
            for (int h = tid; h < vnh && (h >> 5) < a.gmax; h += PN) {
                const int st = a.homeidx[((size_t)vbid * a.gmax + (h >> 5)) * 32 + (h & 31)];
                const int slot = vrecipe[st].x;
                a.fx[slot] = Bf.acc[3 * h]; a.fy[slot] = Bf.acc[3 * h + 1]; a.fz[slot] = Bf.acc[3 * h + 2];
                Bf.acc[3 * h] = 0.0; Bf.acc[3 * h + 1] = 0.0; Bf.acc[3 * h + 2] = 0.0;
            }
        };
        // ghosts written by the neighbours (peer memory): a brick whose halo reaches ghost planes waits for the neighbour's flag
        bool seen_lo = !(VV && P2P && a.wait_epoch != 0), seen_hi = seen_lo;
        auto wait_flag = [&](const unsigned long long *f) {
            if (tid == 0) {
                const long long t0 = flp_clock();
                while (ld_acquire_sys(f) < a.wait_epoch) {
                    __nanosleep(64);
                    if (flp_clock() - t0 > (1ll << 32)) { atomicCAS(a.err, 0, 6); break; }      // ~2 s: the neighbour never published
                }
                __threadfence_system();
            }
            bar_sync(1 + 2 * NBUF, PN);
        };
        // ---- TMA variant: row groups of a brick travel global -> raw ring by bulk copies, issued one group ahead ----------
        unsigned gq = 0, gw = 0;                    // row groups issued / consumed so far (the same in every producer thread)
        auto flags_for = [&](int vbid) {            // P2P: ghosts this brick reads must be complete before they are copied
            if (VV && P2P) {
                const int bzi = vbid / (g.nbx * g.nby);
                if (!seen_lo && bzi < a.p2p_lo_layers) { wait_flag(a.flag_from_lo); seen_lo = true; if (TMA) fence_proxy_async(); }
                if (!seen_hi && bzi >= a.p2p_hi_layer0) { wait_flag(a.flag_from_hi); seen_hi = true; if (TMA) fence_proxy_async(); }
            }
        };
        auto seg_load = [&](int sb, int vbid) {
            for (int t = tid; t < a.segcap; t += PN) segs[sb * a.segcap + t] = a.seg[(size_t)vbid * a.segcap + t];
        };
        // every producer warp issues the copies of the segments it will transform itself (segment t belongs to warp t mod NPROD):
        // cp.async.bulk runs on the uniform datapath, one copy at a time per warp, so four warps issue four at a time
        auto group_issue = [&](int sb, int vnrows, int gi) {
            const int slot = gq & 1;
            const int s0 = 2 * gi * a.raw_rows, s1 = min(2 * vnrows, 2 * (gi + 1) * a.raw_rows);
            double *rw = rawring + (size_t)slot * 3 * a.rawlen;
            unsigned bytes = 0;
            for (int t = s0 + warp + FLP_NPROD * lane; t < s1; t += FLP_NPROD * 32) {
                const int4 e = segs[sb * a.segcap + t];
                const int n = e.y & 0xffff;
                if (n) bytes += 24u * (unsigned)(((e.y >> 16) + n + 1) & ~1);
            }
            bytes = __reduce_add_sync(0xffffffffu, bytes);
            if (lane == 0) mbar_expect_tx(&tma_bar[slot], bytes);       // (one arrival per producer warp completes the phase)
            __syncwarp();
            for (int t = s0 + warp + FLP_NPROD * lane; t < s1; t += FLP_NPROD * 32) {
                const int4 e = segs[sb * a.segcap + t];
                const int n = e.y & 0xffff;
                if (!n) continue;
                const unsigned nb8 = 8u * (unsigned)(((e.y >> 16) + n + 1) & ~1), ro = (unsigned)e.z >> 16;
                bulk_g2s(rw + ro, a.sx + e.x, nb8, &tma_bar[slot]);
                bulk_g2s(rw + a.rawlen + ro, a.sy + e.x, nb8, &tma_bar[slot]);
                bulk_g2s(rw + 2 * a.rawlen + ro, a.sz + e.x, nb8, &tma_bar[slot]);
            }
            gq++;
        };
        if (TMA && brick < nbricks) {
            const int bid0 = FC_BRICK_OF(a, brick);
            seg_load(0, bid0);
            bar_sync(1 + 2 * NBUF, PN);
            flags_for(bid0);
            group_issue(0, brick_geom(g, bid0).nrows, 0);
        }
        for (int k = 0;; k++) {
            const int b = k % NBUF;
            const BrickBuf B = brick_buf(smem_raw + b * bufsz, a.cap, a.ncs_max, MULTI);
            if (brick >= nbricks) {
                FLP_T(tw0);
                if (k >= NBUF && !empty_wait(b, k / NBUF - 1)) return;
                FLP_TACC(0, flp_clock() - tw0);
                if (N3 && held_bid[b] >= 0) flush_forces(B, held_bid[b], held_nh[b]);
                if (tid == 0) B.scal[4] = -1;
                __threadfence_block();
                full_arrive(b);
                if (VV) {
                    // the atoms of the brick just released, then (after its release) those of the brick in the other buffer
                    FLP_T(ta0);
                    if (held_bid[b] >= 0) advance_atoms(held_bid[b], held_nh[b]);
                    FLP_TACC(2, flp_clock() - ta0);
#pragma unroll
                    for (int q = 1; q < NBUF; q++) {
                        const int ob = (k + q) % NBUF;
                        if (held_bid[ob] >= 0) {
                            FLP_T(tw1);
                            if (!empty_wait(ob, (k - NBUF + q) / NBUF)) return;      // (its last fill was step k - NBUF + q)
                            FLP_T(ta1);
                            if (N3) flush_forces(brick_buf(smem_raw + ob * bufsz, a.cap, a.ncs_max, MULTI), held_bid[ob], held_nh[ob]);
                            advance_atoms(held_bid[ob], held_nh[ob]);
                            FLP_TACC(0, ta1 - tw1);
                            FLP_TACC(2, flp_clock() - ta1);
                        }
                    }
                }
#if FLP_TIMING
                if (warp == 0 && lane == 0) {
                    tacc[3] = flp_clock() - t_begin;         // producers' total
                    for (int q = 0; q < 4; q++) atomicAdd(a.timing + q, (unsigned long long)tacc[q]);
                    atomicAdd(a.timing + 6, (unsigned long long)k);
                }
#endif
                break;
            }
            const int bid = FC_BRICK_OF(a, brick);
            const BrickGeom bg = brick_geom(g, bid);
            // the recipe written by k_list_build at the last re-binning: slot and staged-cell coordinates of every staged
            // atom, so staging is one coalesced load, three gathers and ~30 instructions per atom, with no table or scan
            const int2 *recipe = a.recipe + (size_t)bid * a.rcap;
            const int n1 = min(hdr_n1, cap1);                         // staged atoms + 1
            const int nh = hdr_nh;
            const double invM = 1.0 / g.M;
            const int ux0 = bg.hx0 - R, uy0 = bg.hy0 - R, uz0 = (g.zwrap ? bg.hz0 : g.zglob0 + bg.hz0) - R;
            const double bcx = ((double)bg.hx0 + 0.5 * bg.nhx) * invM, bcy = ((double)bg.hy0 + 0.5 * bg.nhy) * invM,
                         bcz = ((double)(uz0 + R) + 0.5 * bg.nhz) * invM;
            constexpr int U = FLP_STAGE_U;
            int2 rc[U];
            double sx[U], sy[U], sz[U];
            auto load_batch = [&](int i0) {
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int i = i0 + u * PN;
                    rc[u] = i < n1 ? recipe[i] : make_int2(0, 0);
                }
#pragma unroll
                // L2-only loads: ghost slots are written by the neighbouring GPUs (peer memory), L1 may hold a stale line
                for (int u = 0; u < U; u++) {
                    if (P2P) { sx[u] = __ldcg(a.sx + rc[u].x); sy[u] = __ldcg(a.sy + rc[u].x); sz[u] = __ldcg(a.sz + rc[u].x); }
                    else { sx[u] = a.sx[rc[u].x]; sy[u] = a.sy[rc[u].x]; sz[u] = a.sz[rc[u].x]; }
                }
            };
            auto store_batch = [&](int i0) {
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int i = i0 + u * PN;
                    if (i >= n1) continue;
                    const int cc = rc[u].y;
                    const double cx = ((double)(ux0 + (cc & 255)) + 0.5) * invM, cy = ((double)(uy0 + ((cc >> 8) & 255)) + 0.5) * invM,
                                 cz = ((double)(uz0 + (N3 ? (cc >> 16) & 31 : cc >> 16)) + 0.5) * invM;
                    if (N3) B.phome[i] = (uint16_t)((unsigned)cc >> 21);
                    double dx = sx[u] - cx, dy = sy[u] - cy, dz = sz[u] - cz;
                    dx -= rint_magic(dx); dy -= rint_magic(dy); dz -= rint_magic(dz);      // image nearest to the staged cell
                    const double px = a.L * (dx + (cx - bcx)), py = a.L * (dy + (cy - bcy)), pzv = a.L * (dz + (cz - bcz));
                    B.pxy[i] = make_double2(px, py);
                    B.pz[i] = pzv;
                    const __half2 hxy = __floats2half2_rn((float)px, (float)py), hz0h = __floats2half2_rn((float)pzv, 0.0f);
                    B.ph[i] = make_uint2(*reinterpret_cast<const unsigned *>(&hxy), *reinterpret_cast<const unsigned *>(&hz0h));
                    if (MULTI) B.ptyp[i] = (uint8_t)a.type[rc[u].x];
                }
            };
            if (!TMA) {
                flags_for(bid);
                // the first batch is requested BEFORE the buffer is free: the producers wait for the consumers ~40 % of the
                // time, and the staging latency that follows the hand-over is what the consumers then wait for
                load_batch(1 + tid);
            }
            if (tid == 0) claimed[k & 1] = gridDim.x + atomicAdd(a.brick_counter, 1);
            bar_sync(1 + 2 * NBUF, PN);        // producers only; ids 1..NBUF are full[], NBUF+1..2*NBUF empty[]
            const int nb = claimed[k & 1];
            {   // one brick ahead: header into registers, recipe into L2 (it streams from HBM; the coordinates it points
                // at were written by k_vv just before this kernel and are L2 hits)
                if (nb < nbricks) {
                    const int nbid = FC_BRICK_OF(a, nb);
                    hdr_n1 = a.brickhdr[2 * nbid];
                    hdr_nh = a.brickhdr[2 * nbid + 1];
                    if (TMA) seg_load((k + 1) & 1, nbid);
                    else {
                        const unsigned char *nr = reinterpret_cast<const unsigned char *>(a.recipe + (size_t)nbid * a.rcap);
                        for (int t = tid * 128; t < a.rcap * 8; t += PN * 128) prefetch_l2(nr + t);
                    }
                }
            }
            {   // the consumers claim this brick's tasks dynamically: bring every group's entry counts and first chunks into L2
                const int ng = min(((nh << SPL) + 31) >> 5, a.gmax);
                for (int t = tid; t < ng * 8; t += PN) {      // 2 chunks x 512 B = 8 lines of 128 B per group
                    const size_t gs = (size_t)bid * a.gmax + (t >> 3);
                    const int part = t & 7;
                    if (LM == 2) {
                        if (part == 0 && (t >> 3) % 32 == 0) prefetch_l2(a.inner_n + gs);
                        prefetch_l2(reinterpret_cast<const unsigned char *>(a.inner8 + gs * a.lcap8 * 32) + part * 128);
                        continue;
                    }
                    if (part == 0) prefetch_l2(a.list_n + gs * 32);
                    prefetch_l2(reinterpret_cast<const unsigned char *>(a.list8 + gs * a.lcap8 * 32) + part * 128);
                }
            }
            FLP_T(tw2);
            if (k >= NBUF && !empty_wait(b, k / NBUF - 1)) return;    // empty[b]: the consumers are done with this buffer
            FLP_TB(ts0, B.scal + 3);
            FLP_TACC(0, ts0 - tw2);
            if (N3 && held_bid[b] >= 0) flush_forces(B, held_bid[b], held_nh[b]);
            if (tid == 0) {
                if (N3) B.phome[0] = 0;
                B.pxy[0] = make_double2(1e30, 1e30);
                B.pz[0] = 1e30;
                const __half2 far = __floats2half2_rn(60000.0f, 60000.0f), farz = __floats2half2_rn(60000.0f, 0.0f);
                B.ph[0] = make_uint2(*reinterpret_cast<const unsigned *>(&far), *reinterpret_cast<const unsigned *>(&farz));
                if (MULTI) B.ptyp[0] = 0;
                B.scal[0] = nh;
                B.scal[1] = n1;
                B.scal[3] = 0;              // task cursor
                B.scal[4] = brick;
            }
            if (TMA) {
                bar_sync(1 + 2 * NBUF, PN);         // the next brick's segment table is complete
                const int ngr = (bg.nrows + a.raw_rows - 1) / a.raw_rows;
                const int sb = k & 1;
                for (int gi = 0; gi < ngr; gi++) {
                    // one group ahead: the next group of this brick, or the first one of the next brick (its buffer need not be free)
                    if (gi + 1 < ngr) group_issue(sb, bg.nrows, gi + 1);
                    else if (nb < nbricks) {
                        const int nbid = FC_BRICK_OF(a, nb);
                        flags_for(nbid);
                        group_issue((k + 1) & 1, brick_geom(g, nbid).nrows, 0);
                    }
                    const int slot = gw & 1;
                    mbar_wait(&tma_bar[slot], (gw >> 1) & 1, a.err);
                    gw++;
                    const double *rw = rawring + (size_t)slot * 3 * a.rawlen;
                    const int s0 = 2 * gi * a.raw_rows, s1 = min(2 * bg.nrows, 2 * (gi + 1) * a.raw_rows);
                    for (int t = s0 + warp; t < s1; t += FLP_NPROD) {
                        const int4 e = segs[sb * a.segcap + t];
                        const int n = e.y & 0xffff;
                        if (!n) continue;
                        const int mis = e.y >> 16, st0 = e.z & 0xffff, ro = (int)((unsigned)e.z >> 16) + mis;
                        // image nearest to the segment's centre in x (a segment spans less than half the box: checked by the
                        // host) and to the row's cell in y and z
                        const double cxc = ((double)(ux0 + (e.w & 255)) + 0.5 * ((e.w >> 8) & 255)) * invM,
                                     cyc = ((double)(uy0 + ((e.w >> 16) & 255)) + 0.5) * invM, czc = ((double)(uz0 + (int)((unsigned)e.w >> 24)) + 0.5) * invM;
                        const double ox = cxc - bcx, oy = cyc - bcy, oz = czc - bcz;
                        constexpr int TU = 4;                  // atoms per lane in flight (the chain load -> image -> store is pure latency)
                        const int nlim = min(n, n1 - st0);
                        for (int r0 = lane; r0 < nlim; r0 += 32 * TU) {
                            double vx[TU], vy[TU], vz[TU];
#pragma unroll
                            for (int u = 0; u < TU; u++) {
                                const int q = ro + min(r0 + 32 * u, nlim - 1);
                                vx[u] = rw[q]; vy[u] = rw[a.rawlen + q]; vz[u] = rw[2 * a.rawlen + q];
                            }
#pragma unroll
                            for (int u = 0; u < TU; u++) {
                                const int r = r0 + 32 * u;
                                if (r >= nlim) break;
                                const int i = st0 + r;
                                double dx = vx[u] - cxc, dy = vy[u] - cyc, dz = vz[u] - czc;
                                dx -= rint_magic(dx); dy -= rint_magic(dy); dz -= rint_magic(dz);
                                const double px = a.L * (dx + ox), py = a.L * (dy + oy), pzv = a.L * (dz + oz);
                                B.pxy[i] = make_double2(px, py);
                                B.pz[i] = pzv;
                                const __half2 hxy = __floats2half2_rn((float)px, (float)py), hz0h = __floats2half2_rn((float)pzv, 0.0f);
                                B.ph[i] = make_uint2(*reinterpret_cast<const unsigned *>(&hxy), *reinterpret_cast<const unsigned *>(&hz0h));
                                if (MULTI) B.ptyp[i] = (uint8_t)a.type[e.x + mis + r];
                            }
                        }
                    }
                    bar_sync(1 + 2 * NBUF, PN);     // this ring slot may be overwritten by the group after next
                }
            } else {
                store_batch(1 + tid);
                for (int i0 = 1 + tid + U * PN; i0 < n1; i0 += U * PN) {
                    load_batch(i0);
                    store_batch(i0);
                }
            }
            __threadfence_block();
            full_arrive(b);                                           // full[b]
            FLP_T(ta2);
            FLP_TACC(1, ta2 - ts0);
            if (VV) {
                if (held_bid[b] >= 0) advance_atoms(held_bid[b], held_nh[b]);     // released before this staging started
                held_bid[b] = bid; held_nh[b] = nh;
            }
            FLP_TACC(2, flp_clock() - ta2);
            brick = nb;
        }
        return;
    }

    // ================================ consumers ================================
    const int ctid = tid - FLP_NPROD * 32;
    const __half thr = __float2half_ru(LM == 1 ? a.rp2h : a.rc2h);      // (a prune step keeps what lies inside rc + skin2)
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    unsigned long long npair = 0;
    double sig2_0 = 0, tt_0 = 0;
    if (!MULTI) { const double2 pr = ljt[0]; sig2_0 = pr.x; tt_0 = pr.y; }
    int pre_n = 0;
    uint4 pre0 = zero4, pre1 = zero4;
    // entry count and first two chunks of (brick, group) into registers, the following chunks into L2
    auto request = [&](int bid_, int grp_) {
        const size_t gs = (size_t)bid_ * a.gmax + grp_;
        if (LM == 2) {          // the inner list: chunks per task (the same for every lane), first two chunks
            pre_n = a.inner_n[gs];
            pre0 = a.inner8[gs * a.lcap8 * 32 + lane];
            pre1 = a.inner8[(gs * a.lcap8 + 1) * 32 + lane];
#pragma unroll
            for (int k = 2; k < 2 + FL_AHEAD; k++) prefetch_l2(a.inner8 + (gs * a.lcap8 + k) * 32 + lane);
            return;
        }
        pre_n = a.list_n[gs * 32 + lane];
        pre0 = a.list8[gs * a.lcap8 * 32 + lane];
        pre1 = a.list8[(gs * a.lcap8 + 1) * 32 + lane];
#pragma unroll
        for (int k = 2; k < 2 + FL_AHEAD; k++) prefetch_l2(a.list8 + (gs * a.lcap8 + k) * 32 + lane);
    };

    int carry = -1;                   // FLP_CARRY: the task already claimed in the brick this warp enters next
    for (int k = 0;; k++) {
        const int b = k % NBUF;
        const BrickBuf B = brick_buf(smem_raw + b * bufsz, a.cap, a.ncs_max, MULTI);
        FLP_T(tc0);
        if (!full_wait(b, k / NBUF)) break;                           // full[b]
        const int brick = B.scal[4];
#if FLP_TIMING
        FLP_TACC(4, flp_clock_after(B.scal + 4) - tc0);
#endif
        if (brick < 0) break;
        const int bid = FC_BRICK_OF(a, brick);
        const double2 *pxy = B.pxy;
        const double *pz = B.pz;
        const uint2 *ph = B.ph;
        const uint8_t *ptyp = B.ptyp;
        const uint16_t *phome = B.phome;
        double *acc = B.acc;
        const int2 *recipe = a.recipe + (size_t)bid * a.rcap;
        const int nh = B.scal[0] << SPL;      // (virtual) home atoms: with split lists every home atom is two lanes
        const int ngroups = (nh + 31) >> 5;
        if (ngroups > a.gmax) atomicCAS(a.err, 0, 5);
        const int ntask = min(ngroups, a.gmax);
        // tasks are claimed from the buffer's cursor; the first claim of a brick finds the list head in L2 (the producers
        // prefetched it), later ones are requested one task ahead
        int grp = carry;
        if (!(FLP_CARRY && FLP_MBAR) || carry < 0) {
            if (lane == 0) grp = atomicAdd(&B.scal[3], 1);
            grp = __shfl_sync(0xffffffffu, grp, 0);
            if (grp < ntask) request(bid, grp);
        }
        carry = -1;

        while (grp < ntask) {
            const int h = (grp << 5) + lane;
            const bool active = h < nh;
            const int hh = active ? h : (grp << 5);
            const int me = a.homeidx[((size_t)bid * a.gmax + grp) * 32 + (hh & 31)];
            EMDEE_CHECK(me >= 1 && me < B.scal[1] && pre_n <= a.lcap8 * 8 && grp < a.gmax, a.err);
            const int slot_i = recipe[me].x;
            const double2 q0 = pxy[me];
            const double pix = q0.x, piy = q0.y, piz = pz[me];
            const uint2 hme = ph[me];
            const __half2 ixy = *reinterpret_cast<const __half2 *>(&hme.x), izw = *reinterpret_cast<const __half2 *>(&hme.y);
            const double2 *ljrow = ljt;
            if (MULTI) ljrow = ljt + (int)ptyp[me] * a.ntypes;
            double fx = 0, fy = 0, fz = 0, e = 0, w = 0;

            const size_t gs_task = (size_t)bid * a.gmax + grp;
            // LM 2: pre_n counts the CHUNKS of the task's inner list (one count per task); else the entries of this lane's list
            int nent = LM == 2 ? 0 : pre_n;
            const int nch = LM == 2 ? pre_n : (nent + 7) >> 3;
            const int nchmax = LM == 2 ? nch : __reduce_max_sync(0xffffffffu, nch);
            const uint4 *lp = a.list8 + gs_task * a.lcap8 * 32 + lane;
            uint4 e0 = 0 < nch ? pre0 : zero4;
            uint4 e1 = 1 < nch ? pre1 : zero4;
            int cnt = 0;
            uint16_t *qp = queue + ctid;
            unsigned tmin = 0xffffffffu;
            unsigned long long np = 0;

            // the arithmetic of one pair, given the partner's staged index and coordinates
            auto pair_math = [&](int j, double2 j0, double jz) {
                const double vx = pix - j0.x, vy = piy - j0.y, vz = piz - jz;
                const double r2 = fma(vz, vz, fma(vy, vy, vx * vx));
                const int t = __double2hiint(r2) - (a.rc2hi - 1);     // pair_in_range: t < 0 inside, t <= 2 borderline
                tmin = min(tmin, (unsigned)t);
                double sig2 = sig2_0, tt = tt_0;
                if (MULTI) { const double2 pr = ljrow[ptyp[j]]; sig2 = pr.x; tt = pr.y; }
                double Eg = 0, Wg = 0;
                double qf = lj_pair_q<EW>(r2, sig2, tt, a.fast, false, 0.0, Eg, Wg);
                qf = t < 0 ? qf : 0.0;
                fx = fma(qf, vx, fx); fy = fma(qf, vy, fy); fz = fma(qf, vz, fz);
                if (EW) { e += t < 0 ? Eg : 0.0; w += t < 0 ? Wg : 0.0; }
                if (COUNT) np += t < 0 ? 1 : 0;
                if (N3) {      // the partner is a home atom of this brick: it does not list me, its share is added here
                    const unsigned hj = phome[j];
                    if (hj != 0 && t < 0) {
                        double *aj = acc + 3 * (hj - 1);
                        atomicAdd(aj, -qf * vx); atomicAdd(aj + 1, -qf * vy); atomicAdd(aj + 2, -qf * vz);
                    }
                }
            };
            auto pair_eval = [&](int idx) {
                const int j = queue[max(idx, -1) * FLP_QS + ctid];
                pair_math(j, pxy[j], pz[j]);
                return j;
            };
            // LM 1: the log of drained entries, four per call (the same number of calls in every lane); two calls make a chunk
            uint4 lg = zero4;
            bool lhalf = false;
            int lchunk = 0;
            uint4 *ip = LM == 1 ? a.inner8 + gs_task * a.lcap8 * 32 + lane : nullptr;
            auto log4 = [&](int j0, int j1, int j2, int j3) {
                const unsigned w0 = (unsigned)j0 | ((unsigned)j1 << 16), w1 = (unsigned)j2 | ((unsigned)j3 << 16);
                if (!lhalf) { lg.x = w0; lg.y = w1; }
                else {
                    lg.z = w0; lg.w = w1;
                    if (lchunk < a.lcap8) ip[(size_t)lchunk * 32] = lg;
                    else atomicCAS(a.err, 0, 5);
                    lchunk++;
                }
                lhalf = !lhalf;
            };
            auto drain = [&](int depth) {
                for (int kk = 0; kk < depth; kk += ILP) {
                    int jj[ILP];
#pragma unroll
                    for (int u = 1; u <= ILP; u++) jj[u - 1] = pair_eval(cnt - u - kk);
                    if (LM == 1) log4(jj[0], jj[1 % ILP], jj[2 % ILP], jj[3 % ILP]);
                }
                cnt = max(cnt - depth, 0);
                qp = queue + cnt * FLP_QS + ctid;
            };
            auto test = [&](unsigned j, uint2 hj) {
                const __half2 dxy = __hsub2(*reinterpret_cast<const __half2 *>(&hj.x), ixy);
                const __half2 dzw = __hsub2(*reinterpret_cast<const __half2 *>(&hj.y), izw);
                const __half2 s = __hfma2(dzw, dzw, __hmul2(dxy, dxy));
                if (__hle(__hadd(__low2half(s), __high2half(s)), thr)) {
                    EMDEE_CHECK(cnt < QCAP, a.err);
                    *qp = (uint16_t)j; qp += FLP_QS; cnt++;
                }
            };

            int ngrp = 0;
            if (LM == 2) {
                // replay of the inner list: every row is a pair to evaluate (or the dummy atom); two groups of four per chunk
                const uint4 *lp2 = a.inner8 + gs_task * a.lcap8 * 32 + lane;
                const int claim_at = max(nch - 3, 0);       // this warp's next task is claimed (and its head requested) near the end
                ngrp = ntask;
                for (int c = 0; c < nch; c++) {
                    const uint4 e2 = c + 2 < nch ? lp2[(size_t)(c + 2) * 32] : zero4;
                    if (c + 2 + FL_AHEAD < nch) prefetch_l2(lp2 + (size_t)(c + 2 + FL_AHEAD) * 32);
                    if (c == claim_at) {
                        if (lane == 0) ngrp = atomicAdd(&B.scal[3], 1);
                        ngrp = __shfl_sync(0xffffffffu, ngrp, 0);
                        if (ngrp < ntask) request(bid, ngrp);
                    }
                    const unsigned q[8] = {e0.x & 0xffffu, e0.x >> 16, e0.y & 0xffffu, e0.y >> 16, e0.z & 0xffffu, e0.z >> 16, e0.w & 0xffffu, e0.w >> 16};
                    EMDEE_CHECK((int)max(max(max(q[0], q[1]), max(q[2], q[3])), max(max(q[4], q[5]), max(q[6], q[7]))) < B.scal[1], a.err);
#pragma unroll
                    for (int hq = 0; hq < 8; hq += 4) {
                        double2 jxy[4];
                        double jz[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) { jxy[u] = pxy[q[hq + u]]; jz[u] = pz[q[hq + u]]; }
#pragma unroll
                        for (int u = 0; u < 4; u++) pair_math((int)q[hq + u], jxy[u], jz[u]);
                    }
                    e0 = e1; e1 = e2;
                }
                if (nch == 0) {      // (a task without rows still has to claim)
                    if (lane == 0) ngrp = atomicAdd(&B.scal[3], 1);
                    ngrp = __shfl_sync(0xffffffffu, ngrp, 0);
                    if (ngrp < ntask) request(bid, ngrp);
                }
            }
            for (int c = 0; LM != 2 && c < nchmax; c++) {
                const uint4 e2 = c + 2 < nch ? lp[(size_t)(c + 2) * 32] : zero4;
                if (c + 2 + FL_AHEAD < nch) prefetch_l2(lp + (size_t)(c + 2 + FL_AHEAD) * 32);
                const unsigned j0 = e0.x & 0xffffu, j1 = e0.x >> 16, j2 = e0.y & 0xffffu, j3 = e0.y >> 16;
                const unsigned j4 = e0.z & 0xffffu, j5 = e0.z >> 16, j6 = e0.w & 0xffffu, j7 = e0.w >> 16;
                EMDEE_CHECK((int)max(max(max(j0, j1), max(j2, j3)), max(max(j4, j5), max(j6, j7))) < B.scal[1], a.err);
                const uint2 h0 = ph[j0], h1 = ph[j1], h2 = ph[j2], h3 = ph[j3], h4 = ph[j4], h5 = ph[j5], h6 = ph[j6], h7 = ph[j7];
                if (FUSE && c > 0) {
                    // pop ILP entries (an empty stack yields the dummy atom) and fetch their coordinates before anything is
                    // pushed; the tests/pushes of this chunk and the FP64 arithmetic of the popped pairs then share one block
                    int pj[ILP];
                    double2 pxyj[ILP];
                    double pzj[ILP];
#pragma unroll
                    for (int u = 0; u < ILP; u++) pj[u] = queue[max(cnt - 1 - u, -1) * FLP_QS + ctid];
#pragma unroll
                    for (int u = 0; u < ILP; u++) { pxyj[u] = pxy[pj[u]]; pzj[u] = pz[pj[u]]; }
                    cnt = max(cnt - ILP, 0);
                    qp = queue + cnt * FLP_QS + ctid;
                    test(j0, h0); test(j1, h1); test(j2, h2); test(j3, h3);
                    test(j4, h4); test(j5, h5); test(j6, h6); test(j7, h7);
#pragma unroll
                    for (int u = 0; u < ILP; u++) pair_math(pj[u], pxyj[u], pzj[u]);
                    if (LM == 1) log4(pj[0], pj[1 % ILP], pj[2 % ILP], pj[3 % ILP]);
                } else {
                    test(j0, h0); test(j1, h1); test(j2, h2); test(j3, h3);
                    test(j4, h4); test(j5, h5); test(j6, h6); test(j7, h7);
                }
                e0 = e1; e1 = e2;
                const int over = __reduce_max_sync(0xffffffffu, cnt) - (QCAP - 8);
                if (over > 0) drain((max(over, FL_MINPOP) + ILP - 1) & ~(ILP - 1));
            }
            // claim this warp's next task of the brick and request its list head; the final drain hides the round trip.  (Claimed
            // at the START of a task, the first warps to reach a brick took two tasks each and left none for the others -- with the
            // mbarrier hand-over the warps arrive one by one -- so half the warps worked on each buffer and nothing was staged ahead.)
            if (LM != 2) {
                if (lane == 0) ngrp = atomicAdd(&B.scal[3], 1);
                ngrp = __shfl_sync(0xffffffffu, ngrp, 0);
            }
            if (LM == 2) {}
            else if (ngrp < ntask) request(bid, ngrp);
            else if (FLP_CARRY && FLP_MBAR && mbar_test(&hand_bar[(k + 1) % NBUF], ((k + 1) / NBUF) & 1)) {
                const BrickBuf Bn = brick_buf(smem_raw + ((k + 1) % NBUF) * bufsz, a.cap, a.ncs_max, MULTI);
                const int nbrick = Bn.scal[4];
                if (nbrick >= 0) {
                    if (lane == 0) carry = atomicAdd(&Bn.scal[3], 1);
                    carry = __shfl_sync(0xffffffffu, carry, 0);
                    if (carry < min((Bn.scal[0] + 31) >> 5, a.gmax)) request(FC_BRICK_OF(a, nbrick), carry);
                }
            }
            if (LM != 2) drain((__reduce_max_sync(0xffffffffu, cnt) + ILP - 1) & ~(ILP - 1));
            if (LM == 1) {      // the last chunk of the log, padded with dummies; rows of the task's inner list
                if (lhalf) {
                    lg.z = 0u; lg.w = 0u;
                    if (lchunk < a.lcap8) ip[(size_t)lchunk * 32] = lg;
                    else atomicCAS(a.err, 0, 5);
                    lchunk++;
                }
                if (lane == 0) a.inner_n[gs_task] = min(lchunk, a.lcap8);
            }

            if (N3 && tmin <= 2u) {
                // pairs within 3e-6 of rc2 were left out by the hot loop (both sides): add the ones the oracle's decision keeps
                const int nslots = ((nent + 7) >> 3) << 3;
                const uint16_t *ent = reinterpret_cast<const uint16_t *>(lp);
                for (int kk = 0; kk < nslots; kk++) {
                    const int j = ent[((kk >> 3) << 8) + (kk & 7)];
                    if (j == 0) continue;
                    const double2 j0 = pxy[j];
                    const double vx = pix - j0.x, vy = piy - j0.y, vz = piz - pz[j];
                    const double r2 = fma(vz, vz, fma(vy, vy, vx * vx));
                    if (pair_in_range(r2, a.rc2hi) != 0) continue;
                    double xval = 0.0;
                    if (!exact_in_range<P2P>(a.sx, a.sy, a.sz, slot_i, recipe[j].x, a.L, a.model, &xval)) continue;
                    double2 pr = ljrow[0];
                    if (MULTI) pr = ljrow[ptyp[j]];
                    double Eg = 0, Wg = 0;
                    const double qf = lj_pair_q<false>(r2, pr.x, pr.y, a.fast, true, xval, Eg, Wg);
                    fx = fma(qf, vx, fx); fy = fma(qf, vy, fy); fz = fma(qf, vz, fz);
                    const unsigned hj = phome[j];
                    if (hj != 0) {
                        double *aj = acc + 3 * (hj - 1);
                        atomicAdd(aj, -qf * vx); atomicAdd(aj + 1, -qf * vy); atomicAdd(aj + 2, -qf * vz);
                    }
                }
            }
            if (!N3 && tmin <= 2u) {     // this lane met a pair within 3e-6 of rc2: redo its list with the oracle's decision
                if (LM == 2) nent = a.list_n[gs_task * 32 + lane];      // (the list of the last re-binning: a superset of the inner list)
                LaneRedo rd;
                rd.pxy = pxy; rd.pz = pz; rd.ptyp = ptyp; rd.ljt = ljt; rd.cs = nullptr; rd.gbase = nullptr; rd.recipe = recipe;
                rd.ncs = 0; rd.ntypes = a.ntypes; rd.me = me; rd.slot_i = slot_i; rd.nent = nent;
                rd.entries = reinterpret_cast<const uint16_t *>(lp);
                rd.sx = a.sx; rd.sy = a.sy; rd.sz = a.sz; rd.L = a.L; rd.model = a.model; rd.fast = a.fast; rd.rc2hi = a.rc2hi;
                double f3[5];
                careful_lane<MULTI, EW, P2P>(rd, f3, &np);
                fx = f3[0]; fy = f3[1]; fz = f3[2];
                if (EW) { e = f3[3]; w = f3[4]; }
            }
            if (COUNT) npair += np;
            bool writer = active;
            if (SPL) {      // the two halves of an atom's list were evaluated by neighbouring lanes: the even one stores the sum
                fx += __shfl_xor_sync(0xffffffffu, fx, 1); fy += __shfl_xor_sync(0xffffffffu, fy, 1); fz += __shfl_xor_sync(0xffffffffu, fz, 1);
                if (EW) { e += __shfl_xor_sync(0xffffffffu, e, 1); w += __shfl_xor_sync(0xffffffffu, w, 1); }
                writer = active && !(lane & 1);
            }
            if (writer) {
                if (N3) {      // own sums join the reactions other lanes have added; the producers write the totals out
                    double *ai = acc + 3 * h;
                    atomicAdd(ai, fx); atomicAdd(ai + 1, fy); atomicAdd(ai + 2, fz);
                } else if (store_f) { a.fx[slot_i] = fx; a.fy[slot_i] = fy; a.fz[slot_i] = fz; }
                if (EW) { a.en[slot_i] = 0.5 * e; a.vir[slot_i] = 0.5 * w; }
            }
            grp = ngrp;
        }
        if (VV || N3) __threadfence_block();                          // the producers read this brick's forces after empty[b]
        empty_arrive(b);                                              // empty[b]
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) npair += __shfl_xor_sync(0xffffffffu, npair, o);
        if (lane == 0 && npair) atomicAdd(a.digest, npair);
    }
#if FLP_TIMING
    if (lane == 0) {
        atomicAdd(a.timing + 4, (unsigned long long)tacc[4]);
        atomicAdd(a.timing + 5, (unsigned long long)(flp_clock() - t_begin));      // consumers' total, summed over warps
    }
#endif
}
