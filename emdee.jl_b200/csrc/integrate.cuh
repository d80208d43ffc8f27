// integrate.cuh -- velocity-Verlet update and host<->slot-order transfer kernels.
//
// The reference has no integrator; the update below is the oracle's definition
// (oracle_vv_steps):  v += (dt/2m) f ; r += dt v ; f = F(r) ; v += (dt/2m) f,
// each line one fma so that CPU and GPU round identically per step.
#pragma once
#include "common.cuh"

struct VVArgs {
    int64_t first, n;          // owned slots
    double *r[3], *s[3], *v[3];
    const double *rb[3];
    const double *f[3];
    const double *mass;
    double dt, L, half_skin2;  // (skin/2)^2
    int pending_kick;          // complete the previous step's second half-kick first
    int drift;                 // 1: kick + drift (+ rescale s), 0: kick only
    int check_skin;
    int *err;
    unsigned *maxd2;           // adaptive re-binning: max over atoms of |r - r_bin|^2 as float bits (atomicMax); may be null
};

// One kernel per step: [second half-kick of step n-1] + first half-kick + drift of step n.
// Reads r, v, f (+ r_bin for the skin check), writes r, v, s: 120 B/atom-step + 48 B for s and r_bin.
__global__ void k_vv(VVArgs a)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const int64_t i = a.first + k;
    const double h = __ddiv_rn(0.5 * a.dt, a.mass[i]);
    double d2 = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double f = a.f[c][i];
        double v = a.v[c][i];
        if (a.pending_kick) v = __fma_rn(h, f, v);
        if (a.drift) {
            v = __fma_rn(h, f, v);
            const double r = __fma_rn(a.dt, v, a.r[c][i]);
            a.r[c][i] = r;
            a.s[c][i] = __ddiv_rn(r, a.L);     // scaled position for the next force evaluation
            const double d = r - a.rb[c][i];
            d2 = fma(d, d, d2);
        }
        a.v[c][i] = v;
    }
    if (a.drift && a.check_skin && d2 > a.half_skin2) atomicCAS(a.err, 0, 3);
    if (a.maxd2) {      // positive floats order like their bit patterns (rounded up: the decision must be conservative)
        const unsigned m = __reduce_max_sync(__activemask(), __float_as_uint(__double2float_ru(d2)));
        if ((threadIdx.x & 31) == 0 && m > *a.maxd2) atomicMax(a.maxd2, m);
    }
}

// K = sum 1/2 m v^2 over owned slots; per-block partials, summed on the host in block order.
__global__ void k_kinetic(int64_t first, int64_t n, const double *__restrict__ vx, const double *__restrict__ vy,
                          const double *__restrict__ vz, const double *__restrict__ mass, double *__restrict__ partial)
{
    __shared__ double sh[256];
    double acc = 0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = first + k;
        acc += 0.5 * mass[i] * (vx[i] * vx[i] + vy[i] * vy[i] + vz[i] * vz[i]);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// v *= factor over owned slots (velocity-rescaling thermostats, driven from the host between emdee_vv_step calls)
__global__ void k_scale3(int64_t first, int64_t n, double factor, double *__restrict__ vx, double *__restrict__ vy, double *__restrict__ vz)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = first + k;
    vx[i] *= factor;
    vy[i] *= factor;
    vz[i] *= factor;
}

// host layout (3xN column-major, id order) <-> slot order
__global__ void k_set3(int64_t first, int64_t n, const int32_t *__restrict__ id, const double *__restrict__ in,
                       double *__restrict__ d0, double *__restrict__ d1, double *__restrict__ d2)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = first + k;
    const int64_t g = id[i];
    d0[i] = in[3 * g]; d1[i] = in[3 * g + 1]; d2[i] = in[3 * g + 2];
}
__global__ void k_get3(int64_t first, int64_t n, const int32_t *__restrict__ id, const double *__restrict__ d0,
                       const double *__restrict__ d1, const double *__restrict__ d2, double *__restrict__ out)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = first + k;
    const int64_t g = id[i];
    out[3 * g] = d0[i]; out[3 * g + 1] = d1[i]; out[3 * g + 2] = d2[i];
}
// The same for a CYCLIC window of the id-ordered host array: rows id0, id0+1, ... (mod N), n of them, held contiguously in
// window order on the device side of the transfer.  Slab ranks move only the rows of the atoms they own; the window is cyclic
// because the periodic box makes the first and the last slab own a few atoms from the other end of the id order.
// Slots [own0, own1) are atoms the rank owns (the window must cover them); ghosts outside the window keep their values (the
// next re-binning refreshes every ghost from its owner).
__device__ __forceinline__ int64_t window_row(int32_t id, int64_t id0, int64_t N)
{
    int64_t g = (int64_t)id - id0;
    return g < 0 ? g + N : g;
}
__global__ void k_set3_range(int64_t first, int64_t cnt, int64_t own0, int64_t own1, const int32_t *__restrict__ id, int64_t id0, int64_t n,
                             int64_t N, const double *__restrict__ in, double *__restrict__ d0, double *__restrict__ d1,
                             double *__restrict__ d2, int *__restrict__ err)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    const int64_t i = first + k;
    const int64_t g = window_row(id[i], id0, N);
    if (g >= n) {
        if (i >= own0 && i < own1) atomicCAS(err, 0, 7);
        return;
    }
    d0[i] = in[3 * g]; d1[i] = in[3 * g + 1]; d2[i] = in[3 * g + 2];
}
__global__ void k_get3_range(int64_t first, int64_t cnt, const int32_t *__restrict__ id, int64_t id0, int64_t n, int64_t N,
                             const double *__restrict__ d0, const double *__restrict__ d1, const double *__restrict__ d2,
                             double *__restrict__ out, int *__restrict__ err)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    const int64_t i = first + k;
    const int64_t g = window_row(id[i], id0, N);
    if (g >= n) {      // slab rank (err != null): an owned atom outside the window means a stale window, its row would be silently missing
        if (err) atomicCAS(err, 0, 7);
        return;
    }
    out[3 * g] = d0[i]; out[3 * g + 1] = d1[i]; out[3 * g + 2] = d2[i];
}
__global__ void k_get1_range(int64_t first, int64_t cnt, const int32_t *__restrict__ id, int64_t id0, int64_t n, int64_t N,
                             const double *__restrict__ d, double *__restrict__ out, int *__restrict__ err)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    const int64_t g = window_row(id[first + k], id0, N);
    if (g < n) out[g] = d[first + k];
    else if (err) atomicCAS(err, 0, 7);
}
// which of `nbuckets` equal id ranges hold an atom of slots [first, first + cnt)
__global__ void k_id_buckets(int64_t first, int64_t cnt, const int32_t *__restrict__ id, int64_t N, int nbuckets, unsigned char *__restrict__ occupied)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    occupied[(int64_t)id[first + k] * nbuckets / N] = 1;
}

// ---- results of a single-point evaluation leaving in chunks of z planes (emdee_compute_nonbonded_into) ---------------------------
// A chunk is the slot range of the cells [cell_bound[k], cell_bound[k + 1]) (z is the slowest digit of the cell index, so a range of
// planes is one contiguous slot range); its atoms' rows of the id-ordered arrays are marked in `nbuckets` equal id ranges.
#define CHUNKS_MAX 32
struct ChunkPlan {
    int nchunks;
    int cell_bound[CHUNKS_MAX + 1];
};
__global__ void k_chunk_id_buckets(int64_t n, ChunkPlan plan, const int32_t *__restrict__ cell_start, const int32_t *__restrict__ id,
                                   int64_t N, int nbuckets, unsigned char *__restrict__ occupied)
{
    __shared__ int bound[CHUNKS_MAX + 1];
    if (threadIdx.x <= plan.nchunks) bound[threadIdx.x] = cell_start[plan.cell_bound[threadIdx.x]];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int k = 0;
    while (k + 1 < plan.nchunks && i >= bound[k + 1]) k++;
    occupied[(int64_t)k * nbuckets + (int64_t)id[i] * nbuckets / N] = 1;
}
// slot order -> id order for the atoms of cells [c0, c1): forces as N x 3 rows, energies, virials (null: not wanted)
__global__ void k_get_chunk(const int32_t *__restrict__ cell_start, int c0, int c1, const int32_t *__restrict__ id,
                            const double *__restrict__ f0, const double *__restrict__ f1, const double *__restrict__ f2,
                            const double *__restrict__ en, const double *__restrict__ vir, double *__restrict__ outF,
                            double *__restrict__ outE, double *__restrict__ outW)
{
    const int64_t first = cell_start[c0], n = cell_start[c1] - first;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = first + k, g = id[i];
        if (outF) { outF[3 * g] = f0[i]; outF[3 * g + 1] = f1[i]; outF[3 * g + 2] = f2[i]; }
        if (outE) outE[g] = en[i];
        if (outW) outW[g] = vir[i];
    }
}

template <typename T>
__global__ void k_set1(int64_t first, int64_t n, const int32_t *__restrict__ id, const T *__restrict__ in,
                       int stride, int off, T *__restrict__ d)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = first + k;
    d[i] = in[(int64_t)stride * id[i] + off];
}
template <typename T>
__global__ void k_get1(int64_t first, int64_t n, const int32_t *__restrict__ id, const T *__restrict__ d,
                       T *__restrict__ out, T add)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = first + k;
    out[id[i]] = d[i] + add;
}
__global__ void k_iota(int64_t n, int32_t *__restrict__ d)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = (int32_t)i;
}
template <typename T>
__global__ void k_fill(int64_t n, T *__restrict__ d, T v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = v;
}

// FP64 FMA throughput probe: 8 independent dependent-chains per thread, 2 flops per DFMA.
__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double seed, double *__restrict__ out)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;   // never true; keeps the chains alive
}
