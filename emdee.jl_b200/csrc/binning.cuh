// binning.cuh -- cell assignment and the (cell, id) counting sort.
//
// Replaces the reference's linked-list construction (distribute!, clean_cells!, collect_baskets!,
// renew_cells!, src/cells.jl:46-174: O(N*cells), racy) with a one-pass radix on the cell key:
//   cell index (src/cells.jl:79-85)  ->  histogram  ->  exclusive scan (warp-shuffle scan)
//   ->  scatter  ->  rank inside each cell by atom id  ->  gather of the per-atom arrays.
// The result is the unique stable order by (cell, id), independent of atomics' arrival order, so the
// integer outputs (cell index, population, permutation) are bit-exact against the oracle.
#pragma once
#include "common.cuh"

// v = min(floor(Int32, M*(s - floor(s))), M-1) for one scaled coordinate (src/cells.jl:79-84 + Q7).
__device__ __forceinline__ int cell_coord(double s, int M)
{
    const double fr = __dsub_rn(s, floor(s));
    const int v = (int)floor(__dmul_rn((double)M, fr));
    return v < M - 1 ? v : M - 1;
}

// s = r/L (src/nonbonded.jl:60-61, src/cells.jl:79-81): IEEE division, one per coordinate per atom.
__global__ void k_scale_positions(int64_t n, const double *__restrict__ rx, const double *__restrict__ ry,
                                  const double *__restrict__ rz, double L, double *__restrict__ sx,
                                  double *__restrict__ sy, double *__restrict__ sz)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sx[i] = __ddiv_rn(rx[i], L);
    sy[i] = __ddiv_rn(ry[i], L);
    sz[i] = __ddiv_rn(rz[i], L);
}

// Global cell coordinates of slots [first, first+n) and the histogram of LOCAL cell indices.
//   gcell[i] = x + M*(y + M*z)   (0-based global index; the reference's index is this + 1)
//   lcell    = x + M*(y + M*(z - zglob0))  with z taken periodically into the local plane range
__global__ void k_cell_index(int64_t first, int64_t n, const double *__restrict__ sx, const double *__restrict__ sy,
                             const double *__restrict__ sz, int M, int zglob0, int nzt, int zwrap,
                             int32_t *__restrict__ gcell, int32_t *__restrict__ lcell, int32_t *__restrict__ count,
                             int *__restrict__ err)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = first + k;
    const int x = cell_coord(sx[i], M), y = cell_coord(sy[i], M), z = cell_coord(sz[i], M);
    gcell[i] = x + M * (y + M * z);
    int zl = z - zglob0;
    if (!zwrap) {            // slab: bring z into [zglob0, zglob0+M) then it must fall in the local planes
        if (zl < 0) zl += M;
        if (zl >= M) zl -= M;
        if (zl >= nzt) { atomicCAS(err, 0, 1); zl = nzt - 1; }
    }
    const int lc = x + M * (y + M * zl);
    lcell[i] = lc;
    atomicAdd(count + lc, 1);
}

// update_cells! (src/cells.jl:196-222) starts by recomputing every atom's cell (clean_cells!, :79-85) and unlinking the atoms
// whose cell changed: count those "movers" against the cell each atom was sorted into at the last binning, and track the
// largest displacement since then (decides whether a pair list built with a skin is still valid).
__global__ void k_count_movers(int64_t first, int64_t n, const double *__restrict__ sx, const double *__restrict__ sy,
                               const double *__restrict__ sz, int M, const int32_t *__restrict__ gcell,
                               const double *__restrict__ rx, const double *__restrict__ ry, const double *__restrict__ rz,
                               const double *__restrict__ bx, const double *__restrict__ by, const double *__restrict__ bz,
                               unsigned long long *__restrict__ movers, unsigned *__restrict__ maxd2)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int moved = 0;
    unsigned d2b = 0;
    if (k < n) {
        const int64_t i = first + k;
        const int x = cell_coord(sx[i], M), y = cell_coord(sy[i], M), z = cell_coord(sz[i], M);
        moved = (x + M * (y + M * z)) != gcell[i];
        const double dx = rx[i] - bx[i], dy = ry[i] - by[i], dz = rz[i] - bz[i];
        d2b = __float_as_uint(__double2float_ru(fma(dz, dz, fma(dy, dy, dx * dx))));
    }
    const unsigned any = __ballot_sync(0xffffffffu, moved);
    d2b = __reduce_max_sync(0xffffffffu, d2b);
    if ((threadIdx.x & 31) == 0) {
        if (any) atomicAdd(movers, (unsigned long long)__popc(any));
        if (d2b > *maxd2) atomicMax(maxd2, d2b);
    }
}

// ---- exclusive scan of int32 counts, three phases, warp-shuffle scans inside each block ----------
#define SCAN_BLOCK 512
#define SCAN_ITEMS 8
__device__ __forceinline__ int warp_inclusive_scan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
// Each block scans SCAN_BLOCK*SCAN_ITEMS entries; writes the exclusive scan in place and its total.
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_block(const int32_t *in, int32_t *out,
                                                           int64_t n, int32_t *__restrict__ block_sum,
                                                           int32_t *__restrict__ maxval)
{
    __shared__ int wsum[SCAN_BLOCK / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t base = ((int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x) * SCAN_ITEMS;
    int v[SCAN_ITEMS], t = 0, mx = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        mx = max(mx, v[k]);
        t += v[k];
    }
    const int inc = warp_inclusive_scan(t, lane);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int s = (lane < SCAN_BLOCK / 32) ? wsum[lane] : 0;
        const int si = warp_inclusive_scan(s, lane);
        if (lane < SCAN_BLOCK / 32) wsum[lane] = si - s;
        if (lane == SCAN_BLOCK / 32 - 1 && block_sum) block_sum[blockIdx.x] = si;
    }
    __syncthreads();
    int run = wsum[w] + inc - t;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (maxval) {
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0 && mx > 0) atomicMax(maxval, mx);
    }
}
__global__ void k_scan_add(int32_t *__restrict__ out, int64_t n, const int32_t *__restrict__ block_off)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] += block_off[i / (SCAN_BLOCK * SCAN_ITEMS)];
}

// Scatter: provisional place of every slot inside its cell (arrival order, fixed by k_rank_in_cell).
__global__ void k_scatter(int64_t first, int64_t n, const int32_t *__restrict__ lcell,
                          const int32_t *__restrict__ cell_start, int32_t *__restrict__ fill,
                          int32_t *__restrict__ order)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = first + k;
    const int c = lcell[i];
    order[cell_start[c] + atomicAdd(fill + c, 1)] = (int32_t)i;
}

// Deterministic order inside each cell: ascending global id.  One thread per sorted position p;
// rank = number of cell members with a smaller id.  Cells hold ~2-100 atoms, so this is O(n) reads
// from L1 per atom.  dest[p] = old slot that belongs at new slot p.
__global__ void k_rank_in_cell(int64_t pfirst, int64_t n, const int32_t *__restrict__ order,
                               const int32_t *__restrict__ lcell, const int32_t *__restrict__ cell_start,
                               const int32_t *__restrict__ id, int32_t *__restrict__ src_of_new)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t p = pfirst + k;
    const int old = order[p];
    const int c = lcell[old];
    const int myid = id[old];
    const int b = cell_start[c], e = cell_start[c + 1];
    int rank = 0;
    for (int q = b; q < e; q++) rank += (id[order[q]] < myid);
    src_of_new[b + rank] = old;
}

// Gather the per-atom arrays into the new order.  One thread per new slot.
struct GatherArgs {
    int64_t pfirst, n;
    const int32_t *src_of_new;
    const int32_t *gcell_old, *lcell_old;
    AtomArrays in, out;
    int32_t *gcell_new, *lcell_new;
    int32_t *slot_of_id;
    int has_vel, has_excl;
};
__global__ void k_gather(GatherArgs a)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const int64_t p = a.pfirst + k;
    const int o = a.src_of_new[p];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double r = a.in.r[c][o];
        a.out.r[c][p] = r;
        a.out.rb[c][p] = r;
        a.out.s[c][p] = a.in.s[c][o];
        if (a.has_vel) a.out.v[c][p] = a.in.v[c][o];
    }
    a.out.hs[p] = a.in.hs[o];
    a.out.ts[p] = a.in.ts[o];
    a.out.mass[p] = a.in.mass[o];
    const int id = a.in.id[o];
    a.out.id[p] = id;
    a.out.type[p] = a.in.type[o];
    if (a.has_excl) { a.out.xbase[p] = a.in.xbase[o]; a.out.xmask[p] = a.in.xmask[o]; }
    a.gcell_new[p] = a.gcell_old[o];
    a.lcell_new[p] = a.lcell_old[o];
    if (a.slot_of_id) a.slot_of_id[id] = (int32_t)p;
}
