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
// This kernel serves single-point evaluations (forces / energies / virials, audits).  The velocity-Verlet loop
// uses the pair list instead: k_list_build (list_build.cuh) on a re-binning step, k_force_list (force_list.cuh)
// on every step; both stage a brick in the same order as this kernel.
// Geometry in the drain: every staged atom carries FP64 coordinates in the brick's frame (periodic image
// resolved per staged cell), so a separation is three subtractions; the cutoff decision is made on that r2
// with integer compares and handed to the oracle's exact rounding sequence only inside a 3e-6 band around
// rc2 (pair_in_range, lj_pair.cuh) -- the pair set stays bit-exact.
#pragma once
#include "lj_pair.cuh"

#define FC_QCAP 64          // per-lane queue entries (uint16 staged indices)
#define FC_QCHUNK 16        // candidates scanned between queue-capacity checks
#define FC_MINPOP 8         // a partial drain pops at least this many entries per lane
#define FC_MAX_HOMEROWS 64  // by*bz of the largest supported brick
#define FC_MAX_TYPES 16     // LJ parameter classes held as a pair table in shared memory

#define FC_MAX_NCS_SMALL 512   // staged cells of the bricks k_brick_keep_max handles (static tables)
#define FC_DIMTAB 96         // 3 x 32 doubles: a staged dimension has at most 32 cells

#define FC_MAX_BLOCK 384     // launch bound: 12 warps, up to 170 registers per thread

struct CellArgs {
    GridDesc g;
    const int32_t *cell_start;   // local cells + 1
    const double *sx, *sy, *sz, *hs, *ts;
    const int32_t *id, *type, *xbase;
    const uint64_t *xmask;
    const double2 *ljtab;             // ntypes^2 entries {(half_sigma_a + half_sigma_b)^2, twice_sqrt_eps_a * twice_sqrt_eps_b}
    int ntypes;                       // 0: more than FC_MAX_TYPES classes, per-atom parameters are gathered from global
    double *fx, *fy, *fz, *en, *vir;
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
    int block_first;                  // first brick of this launch (launches may cover a z-layer range) ...
    int block_split, block_first2;    // ... or two ranges: launch index i >= block_split maps to block_first2 + (i - block_split)
    int block_split2, block_first3;   // ... or three: i >= block_split + block_split2 maps to block_first3 + (i - block_split - block_split2)
    // pair list: home atom h of a brick (flattened over its home rows) belongs to group h/32, lane h%32;
    // chunk c of that atom is the uint4 list8[((brick*gmax + h/32)*lcap8 + c)*32 + h%32] = 8 x (staged index + 1)
    uint4 *list8;
    uint16_t *list_n;                 // entries stored per home atom: list_n[(brick*gmax + h/32)*32 + h%32]
    int gmax, lcap8;                  // groups per brick, chunk capacity per atom
    float rl2f;                       // FP32 threshold of the list: (rc + skin)^2 * (1 + margin)
    int rc2hi;                        // high word of rc2 (pair_in_range)
    LJFast fast;
    float rc2h;                       // k_force_list: FP16 pre-cull threshold (conservative, set by the host)
    float rl2h;                       // k_list_build: FP16 threshold of the list, (rc + skin)^2 + rounding bound
    // two-level list of k_force_list_p (LM): the inner list of a warp task -- the rows a prune step drained, in the layout of list8
    // (chunk c of lane l: inner8[((brick*gmax + group)*lcap8 + c)*32 + l]) -- its chunks per task, the prune step's FP16 threshold
    uint4 *inner8;
    int *inner_n;
    float rp2h;
    int qcap;                         // per-lane stack entries of k_force_list_p
    // compacted staging (dense cells): k_list_build keeps only the staged atoms within rc + skin of the brick's home box
    // (about 76 % of the 27 cells around a one-cell brick) and numbers them in staging order; the recipe it writes lists those
    // atoms only, so k_force_list_p stages the compacted brick without knowing about it
    int compact;
    int split;                        // 1: two lanes of the stepping kernel per home atom (few home atoms per brick: dense cells)
    double keep2;                     // (rc + skin)^2 with a margin for rounding
    unsigned *vv_maxstep;             // fused integrator: max over atoms of |r(n+1) - r(n)|^2 (float bits; the host resets it at a prune step)
    // staging recipe, written by k_list_build and valid until the next re-binning (the persistent kernel stages from it
    // instead of rebuilding the cell table on every step):
    int2 *recipe;                     // recipe[brick*rcap + i], i = staged index + 1: {global slot, cx | cy<<8 | cz<<16}
    uint16_t *homeidx;                // homeidx[(brick*gmax + h/32)*32 + h%32] = staged index + 1 of home atom h
    int *brickhdr;                    // brickhdr[2*brick] = staged atoms + 1, [2*brick+1] = home atoms
    int rcap;
    // velocity-Verlet fused into the stepping kernel's epilogue (VV variant of k_force_list_p): the atom's own thread
    // completes step n (second half-kick) and starts step n+1 (first half-kick, drift) as soon as its force is known;
    // the new scaled positions go to a second buffer because other bricks still stage the old ones
    int vv_mode;                      // 1: second half-kick only (last step), 2: + first half-kick and drift of the next step
    double vv_dt, vv_half_skin2;
    double *vv_v[3], *vv_r[3], *vv_snew[3];
    const double *vv_rb[3], *vv_mass;
    unsigned *vv_maxd2;               // adaptive re-binning (may be null)
    int vv_check_skin;
    int *brick_counter;               // persistent kernel: bricks beyond the first of each block are claimed here (zeroed per launch)
    unsigned long long *timing;       // -DFLP_TIMING=1 builds: cycle counters of the persistent kernel's roles (else unused)
    // Slab decomposition with peer-mapped halos (k_force_list_p, VV variant): the integrator writes the new scaled position of
    // every atom of my boundary planes straight into the neighbour's ghost slots over NVLink (peer memory), then raises a flag
    // there once all bricks of that side are done; bricks whose halo reaches ghost planes wait for the neighbour's flag.
    int p2p;                          // 0: off (single GPU, or halo by ncclSend/ncclRecv)
    int p2p_lo_layers, p2p_hi_layer0; // brick z-layers [0, lo_layers) read lower ghosts; layers >= hi_layer0 read upper ghosts
    int p2p_nlo, p2p_nhi;             // bricks holding atoms the lower / upper neighbour needs (they publish when all are advanced)
    int lo_send_a, lo_send_n, hi_send_a, hi_send_n;   // my slot ranges the neighbours hold as ghosts
    double *peer_lo[3], *peer_hi[3];  // neighbours' NEXT-step scaled-position arrays (peer-mapped)
    const long long *peer_info;       // [0] lower neighbour's first upper-ghost slot (my lower planes land there); upper ghosts start at 0
    unsigned long long *flag_lo_peer, *flag_hi_peer;      // flags I raise: in the lower neighbour's "from upper" word, the upper neighbour's "from lower" word
    const unsigned long long *flag_from_lo, *flag_from_hi; // flags I wait on (local memory, written by the neighbours)
    unsigned long long wait_epoch;    // ghosts of this launch are complete when the flags reach this value (0: already complete)
    unsigned long long publish_epoch; // value I raise after my boundary atoms are advanced (0: nothing is published)
    int *p2p_done;                    // [0] lower-side bricks advanced, [1] upper-side (zeroed per launch)
    // TMA staging of the persistent kernel: segment table written by k_list_build -- per brick, two entries per staged (y, z) row
    // (the part before and behind the periodic seam): {first slot rounded down to even, atoms | misalignment << 16,
    // staged index + 1 of the first atom | offset in the raw group << 16, first staged cell x | cells << 8 | cy << 16 | cz << 24}
    int4 *seg;
    int segcap;                       // entries per brick
    int raw_rows, rawlen;             // rows per raw group, doubles per coordinate array of a group
};

// Brick handled by launch index i (a launch covers one or two contiguous ranges of bricks).
#define FC_BRICK_OF(a, i)                                                                                      \
    ((i) < (a).block_split ? (a).block_first + (i)                                                             \
                           : ((i) - (a).block_split < (a).block_split2 ? (a).block_first2 + ((i) - (a).block_split) \
                                                                       : (a).block_first3 + ((i) - (a).block_split - (a).block_split2)))

// Brick geometry shared by the kernels that stage a brick.
struct BrickGeom {
    int hx0, hy0, hz0;      // first home cell (x, y global; z local)
    int nhx, nhy, nhz;      // home cells per dimension
    int sxn, syn, szn;      // staged cells per dimension (home + 2R)
    int nrows, ncs;         // staged rows (y,z) and staged cells
};
__device__ __forceinline__ BrickGeom brick_geom(const GridDesc &g, int bid)
{
    BrickGeom b;
    int t = bid;
    const int bxi = t % g.nbx; t /= g.nbx;
    const int byi = t % g.nby;
    const int bzi = t / g.nby;
    b.hx0 = bxi * g.bx; b.hy0 = byi * g.by; b.hz0 = g.zhome0 + bzi * g.bz;
    b.nhx = min(g.bx, g.M - b.hx0); b.nhy = min(g.by, g.M - b.hy0); b.nhz = min(g.bz, g.zhome0 + g.nzhome - b.hz0);
    b.sxn = b.nhx + 2 * g.R; b.syn = b.nhy + 2 * g.R; b.szn = b.nhz + 2 * g.R;
    b.nrows = b.syn * b.szn; b.ncs = b.nrows * b.sxn;
    return b;
}

// Global slot of staged atom j (binary search of the staged-cell table; used by the rare exact-cull path).
__device__ __noinline__ int staged_slot(int j, const int *cs, const int *gbase, int ncs)
{
    int lo = 0, hi = ncs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (cs[mid] <= j) lo = mid; else hi = mid - 1;
    }
    return gbase[lo] + (j - cs[lo]);
}

// The oracle's cutoff decision for one pair (dist2() of the oracle on the scaled coordinates), taken only
// for pairs whose local-frame r2 is within 3*2^-20 of rc2.  Also returns the oracle's own clamped x
// (src/lennard_jones.jl:36-37: x < 0 -> 0, x > 1 -> 0, x == 1 -> 0.5) so that the switching function is
// evaluated on the same branch.
// CG: L2-only loads.  In a slab decomposition with peer-mapped halos the ghost positions are written by the neighbouring
// GPUs while the kernel runs, and L1 may hold a stale copy of a line that straddles owned and ghost slots.  (A separate
// instantiation: the plain-load version must stay exactly as it is -- its loads are part of what the compiler schedules
// around in the callers' hot loops.)
__device__ __forceinline__ double ld_cg_f64(const double *p)      // L2-only load the compiler may schedule like a plain one
{
    double v;
    asm("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
template <bool CG = false>
__device__ __noinline__ bool exact_in_range(const double *sx, const double *sy, const double *sz, int slot_i, int slot_j,
                                            double L, const LJModel m, double *xval)
{
    double vx, vy, vz;
    double r2;
    if (CG)
        r2 = min_image_r2(ld_cg_f64(sx + slot_i), ld_cg_f64(sy + slot_i), ld_cg_f64(sz + slot_i), ld_cg_f64(sx + slot_j), ld_cg_f64(sy + slot_j),
                          ld_cg_f64(sz + slot_j), L, vx, vy, vz);
    else
        r2 = min_image_r2(sx[slot_i], sy[slot_i], sz[slot_i], sx[slot_j], sy[slot_j], sz[slot_j], L, vx, vy, vz);
    double x = __dmul_rn(__dsub_rn(r2, m.rs2), m.id2);
    x = (x < 0.0 || x > 1.0) ? 0.0 : (x == 1.0 ? 0.5 : x);
    *xval = x;
    return r2 <= m.rc2;
}

__host__ __device__ inline size_t fc_smem_bytes(int cap, int ncs_max, int block, bool typed)
{
    size_t b = (size_t)cap * (2 * sizeof(double2) + sizeof(float4));
    if (!typed) b += (size_t)(cap + (cap & 1)) * sizeof(int);   // slot of every staged atom (per-atom parameter gathers)
    b += (size_t)FC_MAX_TYPES * FC_MAX_TYPES * sizeof(double2);
    b += FC_DIMTAB * sizeof(double);                 // scaled cell centres of the three staged dimensions
    b += (size_t)(ncs_max + 1) * sizeof(int) * 3;    // cs[], gbase[], ccoord[]
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

// ---- staging shared by k_force_cells and k_force_list ---------------------------------------------------
// Phase A: count and first global slot of every staged cell, its coordinates inside the staged box, and the
// unwrapped scaled centres of the staged cells along each dimension.  The caller turns cs[] into an exclusive
// prefix afterwards.
// tid / BLOCK: index and size of the thread group that stages (the whole block, or the producer warps of the
// persistent kernel).
__device__ __forceinline__ void stage_cell_table(const CellArgs &a, const BrickGeom &bg, int *cs, int *gbase, int *ccoord, double *ctab,
                                                 int tid, int BLOCK)
{
    const GridDesc &g = a.g;
    const int R = g.R, M = g.M;
    for (int t = tid; t < bg.ncs; t += BLOCK) {
        const int cx = t % bg.sxn, row = t / bg.sxn, cy = row % bg.syn, cz = row / bg.syn;
        const int gx = wrap_mod(bg.hx0 - R + cx, M), gy = wrap_mod(bg.hy0 - R + cy, M);
        int lz = bg.hz0 - R + cz;
        if (g.zwrap) lz = wrap_mod(lz, M);
        const int lc = gx + M * (gy + M * lz);
        const int s0 = a.cell_start[lc];
        gbase[t] = s0;
        cs[t] = a.cell_start[lc + 1] - s0;
        ccoord[t] = cx | (cy << 8) | (cz << 16);
    }
    const int uz0 = (g.zwrap ? bg.hz0 : g.zglob0 + bg.hz0) - R;
    for (int t = tid; t < FC_DIMTAB; t += BLOCK) {
        const int d = t >> 5, k = t & 31;
        const int u = (d == 0 ? bg.hx0 - R : d == 1 ? bg.hy0 - R : uz0) + k;
        ctab[t] = ((double)u + 0.5) / M;
    }
}

// Phase B: every staged atom [first, last) in the brick's frame.  For staged index idx the cell is found by a
// fixed-length binary search of the prefix table (three atoms per thread in flight, so their global loads
// overlap); p = L * ((s - c) - rint(s - c) + (c - b)) with c the unwrapped scaled centre of the staged cell
// (the periodic image nearest to that cell) and b the brick centre.  store(idx, slot, cell code, px, py, pz).
template <int U = 3, class Store>
__device__ __forceinline__ void stage_atoms(const CellArgs &a, const BrickGeom &bg, const int *cs, const int *gbase, const int *ccoord,
                                            const double *ctab, int first, int last, int tid, int BLOCK, Store store)
{
    const GridDesc &g = a.g;
    const int M = g.M;
    const int uz0 = (g.zwrap ? bg.hz0 : g.zglob0 + bg.hz0);
    const double bcx = ((double)bg.hx0 + 0.5 * bg.nhx) / M, bcy = ((double)bg.hy0 + 0.5 * bg.nhy) / M, bcz = ((double)uz0 + 0.5 * bg.nhz) / M;
    int nsteps = 0;
    while ((1 << nsteps) < bg.ncs) nsteps++;
    for (int i0 = first + tid; i0 < last; i0 += U * BLOCK) {
        int idx[U], lo[U], slot[U];
        bool ok[U];
        double sx[U], sy[U], sz[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            idx[u] = i0 + u * BLOCK;
            ok[u] = idx[u] < last;
            lo[u] = 0;
        }
        for (int st = nsteps - 1; st >= 0; st--) {       // last t with cs[t] <= idx (empty cells repeat a value)
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int mid = lo[u] + (1 << st);
                if (mid < bg.ncs && cs[mid] <= idx[u]) lo[u] = mid;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            slot[u] = gbase[lo[u]] + (idx[u] - cs[lo[u]]);
            sx[u] = sy[u] = sz[u] = 0.0;
            if (ok[u]) { sx[u] = a.sx[slot[u]]; sy[u] = a.sy[slot[u]]; sz[u] = a.sz[slot[u]]; }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!ok[u]) continue;
            const int cc = ccoord[lo[u]];
            const double cx = ctab[cc & 255], cy = ctab[32 + ((cc >> 8) & 255)], cz = ctab[64 + (cc >> 16)];
            double dx = sx[u] - cx, dy = sy[u] - cy, dz = sz[u] - cz;
            dx -= rint_magic(dx); dy -= rint_magic(dy); dz -= rint_magic(dz);      // image nearest to the staged cell
            store(idx[u], slot[u], cc, a.L * (dx + (cx - bcx)), a.L * (dy + (cy - bcy)), a.L * (dz + (cz - bcz)));
        }
    }
}

template <bool F, bool EW, bool EXCL, bool AUDIT, bool TYPED>
__global__ void __launch_bounds__(FC_MAX_BLOCK, 1) k_force_cells(CellArgs a)
{
    const int BLOCK = blockDim.x;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GridDesc &g = a.g;
    const int cap = a.cap;
    double2 *pxy = reinterpret_cast<double2 *>(smem_raw);      // {x, y} in the brick's frame
    double2 *pzm = pxy + cap;                                  // {z, (type << 32) | id}
    float4 *prel = reinterpret_cast<float4 *>(pzm + cap);      // FP32 copy {x, y, z, |p|^2} for the scan
    constexpr bool typed = TYPED;    // LJ classes through the shared-memory pair table; else per-atom gathers
    int *sslot = reinterpret_cast<int *>(prel + cap);
    double2 *ljt = reinterpret_cast<double2 *>(sslot + (typed ? 0 : cap + (cap & 1)));
    double *ctab = reinterpret_cast<double *>(ljt + FC_MAX_TYPES * FC_MAX_TYPES);
    int *cs = reinterpret_cast<int *>(ctab + FC_DIMTAB);
    int *gbase = cs + (a.ncs_max + 1);
    int *ccoord = gbase + (a.ncs_max + 1);
    int *hstart = ccoord + (a.ncs_max + 1);
    int *tstart = hstart + (FC_MAX_HOMEROWS + 1);
    int *scal = tstart + (FC_MAX_HOMEROWS + 1);
    uint16_t *queue = reinterpret_cast<uint16_t *>(
        smem_raw + ((reinterpret_cast<unsigned char *>(scal + 8) - smem_raw + 15) & ~(size_t)15));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = g.R;

    // ---- brick geometry ---------------------------------------------------------------------
    const int bid = FC_BRICK_OF(a, (int)blockIdx.x);
    const BrickGeom bg = brick_geom(g, bid);
    const int nhx = bg.nhx, nhy = bg.nhy, nhz = bg.nhz;
    const int sxn = bg.sxn, syn = bg.syn, ncs = bg.ncs;

    // ---- phase A: staged-cell table (count and first global slot of every staged cell) --------
    stage_cell_table(a, bg, cs, gbase, ccoord, ctab, tid, BLOCK);
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

    // ---- phase B: stage every atom of the brick and its halo in the brick's frame ---------------
    stage_atoms(a, bg, cs, gbase, ccoord, ctab, 0, nstaged, tid, BLOCK, [&](int idx, int slot, int, double px, double py, double pz) {
        if (!typed) sslot[idx] = slot;
        pxy[idx] = make_double2(px, py);
        pzm[idx] = make_double2(pz, __hiloint2double(typed ? a.type[slot] : 0, a.id[slot]));
        float4 p;
        p.x = (float)px; p.y = (float)py; p.z = (float)pz;
        p.w = fmaf(p.z, p.z, fmaf(p.y, p.y, p.x * p.x));
        prel[idx] = p;
    });
    __syncthreads();

    // ---- phase C: warp tasks -----------------------------------------------------------------
    const double L = a.L;
    const float scan2f = a.rc2f;                          // what the window scan accepts
    const int nwin = 2 * R + 1;
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
        const double pix = q0.x, piy = q0.y, piz = q1.x;
        const int32_t idi = __double2loint(q1.y), typi = __double2hiint(q1.y);
        double hsi = 0, tsi = 0;
        if (!typed) { hsi = a.hs[slot_i]; tsi = a.ts[slot_i]; }
        const double2 *ljrow = ljt + typi * a.ntypes;
        int32_t xb = 0; uint64_t xm = 0;
        if (EXCL) { xb = a.xbase[slot_i]; xm = a.xmask[slot_i]; }
        double fx = 0, fy = 0, fz = 0, e = 0, w = 0;
        int cnt = 0;                            // entries on this lane's stack: queue[k*BLOCK + tid], k < cnt
        uint16_t *qp = queue + tid;             // next free entry

        // one stack entry, branch-free except for the rare exact-cull path: invalid or culled entries contribute nothing
        auto pair_eval = [&](int idx, bool valid) {
            const int j = valid ? (int)queue[idx * BLOCK + tid] : me;
            const double2 j0 = pxy[j], j1 = pzm[j];
            const double vx = pix - j0.x, vy = piy - j0.y, vz = piz - j1.x;
            const double r2 = fma(vz, vz, fma(vy, vy, vx * vx));
            const int where = pair_in_range(r2, a.rc2hi);
            bool in = where < 0, xover = false;
            double xval = 0.0;
            if (where == 0 && valid && j != me) {          // within 3e-6 of rc2: the oracle's rounding sequence decides
                const int slot_j = typed ? staged_slot(j, cs, gbase, ncs) : sslot[j];
                in = exact_in_range(a.sx, a.sy, a.sz, slot_i, slot_j, L, a.model, &xval);
                xover = true;
            }
            bool ok = valid && (j != me) && in;
            const int32_t idj = __double2loint(j1.y);
            if (EXCL) ok = ok && !pair_excluded(xb, xm, idj);
            double sig2, tt;
            if (!typed) {
                const int slot_j = sslot[j];
                const double sig = hsi + a.hs[slot_j];
                sig2 = sig * sig;
                tt = tsi * a.ts[slot_j];
            } else {
                const double2 pr = ljrow[__double2hiint(j1.y)];   // one class: every lane reads entry 0 (broadcast)
                sig2 = pr.x; tt = pr.y;
            }
            double Eg = 0, Wg = 0;
            const double qf = lj_pair_q<EW>(r2, sig2, tt, a.fast, xover, xval, Eg, Wg);
            if (ok) {
                if (F) { fx = fma(qf, vx, fx); fy = fma(qf, vy, fy); fz = fma(qf, vz, fz); }
                if (EW) { e += Eg; w += Wg; }
            }
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
                pair_eval(cnt - 1 - k, k < cnt);
                pair_eval(cnt - 2 - k, k + 1 < cnt && k + 1 < depth);
                pair_eval(cnt - 3 - k, k + 2 < cnt && k + 2 < depth);
                pair_eval(cnt - 4 - k, k + 3 < cnt && k + 3 < depth);
            }
            cnt = max(cnt - depth, 0);
            qp = queue + cnt * BLOCK + tid;
        };
        // a candidate that passed the FP32 test is stacked for evaluation
        auto accept = [&](int p, float) { *qp = (uint16_t)p; qp += BLOCK; cnt++; };

        // -------- scan: (2R+1)^2 rows, one shared window of cells [cxa-R, cxb+R] per row --------
        for (int rw = 0; rw < nwin * nwin; rw++) {
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
                    if (r0 <= scan2f) accept(p, r0);
                    if (r1 <= scan2f) accept(p + 1, r1);
                    if (r2 <= scan2f) accept(p + 2, r2);
                    if (r3 <= scan2f) accept(p + 3, r3);
                }
                for (; p < pe; p++) {
                    const float4 c = prel[p];
                    const float r = fmaf(c.x, m2x, fmaf(c.y, m2y, fmaf(c.z, m2z, c.w + pp)));
                    if (r <= scan2f) accept(p, r);
                }
                const int over = __reduce_max_sync(0xffffffffu, cnt) - (FC_QCAP - FC_QCHUNK);
                if (over > 0) drain(max(over, FC_MINPOP));
            }
        }
        drain(__reduce_max_sync(0xffffffffu, cnt));
        if (active) {
            if (F) { a.fx[slot_i] = fx; a.fy[slot_i] = fy; a.fz[slot_i] = fz; }
            if (EW) {
                e *= 0.5; w *= 0.5;                         // src/nonbonded.jl:93-94
                a.en[slot_i] = e; a.vir[slot_i] = w;
            }
        }
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
