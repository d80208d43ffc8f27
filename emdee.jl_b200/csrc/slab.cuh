// slab.cuh -- kernels of the 1-D slab decomposition (one process per GPU, z planes of the cell grid).
//
// The reference is single-GPU; this is the B200-native scaling mechanism the north star
// prescribes: the slowest digit of the reference's cell index (z, src/cells.jl:85) is split into
// contiguous plane ranges, one per rank.  Because atoms are sorted by (cell, id) with z slowest, the
// R boundary planes a neighbour needs as ghosts are ONE contiguous slot range per side, and the
// ghosts a rank receives land in one contiguous range at the head (from below) or tail (from
// above) of its arrays: the per-step halo exchange is plain ncclSend/ncclRecv of array slices with
// no pack/unpack kernels.  Only migration at re-binning gathers scattered leavers into a buffer.
#pragma once
#include "binning.cuh"

#define MIG_FIELDS 12   // doubles per migrating atom: r(3) v(3) hs ts mass | (id,type) | xbase | xmask

// Classify owned atoms at a slab re-bin.  Planes [z0, z0+nz) are mine; an atom that left goes to the
// lower / upper neighbour if it is within R planes of my range (it cannot have moved further between
// re-bins), else err.  first_time: every rank still holds all atoms and simply drops the foreign ones.
// Leavers and foreigners get lcell = ncell (never scattered).
__global__ void k_cell_index_slab(int64_t first, int64_t n, const double *__restrict__ sx, const double *__restrict__ sy,
                                  const double *__restrict__ sz, int M, int z0, int nz, int R, int first_time,
                                  int ncell, int32_t *__restrict__ gcell, int32_t *__restrict__ lcell,
                                  int32_t *__restrict__ count, int32_t *__restrict__ sendcount,
                                  int32_t *__restrict__ list_lo, int32_t *__restrict__ list_hi, int list_cap,
                                  int *__restrict__ err)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = first + k;
    const int x = cell_coord(sx[i], M), y = cell_coord(sy[i], M), z = cell_coord(sz[i], M);
    gcell[i] = x + M * (y + M * z);
    int dz = z - z0;
    if (dz < 0) dz += M;
    if (dz < nz) {
        const int lc = x + M * (y + M * (dz + R));
        lcell[i] = lc;
        atomicAdd(count + lc, 1);
        return;
    }
    lcell[i] = ncell;
    if (first_time) return;
    if (dz - nz < M - dz) {          // nearer to my top: moved up
        const int p = atomicAdd(sendcount + 1, 1);
        if (p < list_cap) list_hi[p] = (int32_t)i; else atomicCAS(err, 0, 4);
    } else {
        const int p = atomicAdd(sendcount + 0, 1);
        if (p < list_cap) list_lo[p] = (int32_t)i; else atomicCAS(err, 0, 4);
    }
}

// Gather the leavers of one direction into a field-major buffer (field f of atom a at buf[f*n + a]).
__global__ void k_pack_migrants(int n, const int32_t *__restrict__ list, AtomArrays A, double *__restrict__ buf)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int i = list[a];
#pragma unroll
    for (int c = 0; c < 3; c++) { buf[(size_t)c * n + a] = A.r[c][i]; buf[(size_t)(3 + c) * n + a] = A.v[c][i]; }
    buf[(size_t)6 * n + a] = A.hs[i];
    buf[(size_t)7 * n + a] = A.ts[i];
    buf[(size_t)8 * n + a] = A.mass[i];
    buf[(size_t)9 * n + a] = __hiloint2double(A.type[i], A.id[i]);
    buf[(size_t)10 * n + a] = __longlong_as_double((long long)A.xbase[i]);
    buf[(size_t)11 * n + a] = __longlong_as_double((long long)A.xmask[i]);
}

// Append arrivals behind the owned atoms and bin them (they must fall into my planes).
__global__ void k_unpack_migrants(int n, const double *__restrict__ buf, int64_t first_slot, AtomArrays A, double L,
                                  int M, int z0, int nz, int R, int32_t *__restrict__ gcell, int32_t *__restrict__ lcell,
                                  int32_t *__restrict__ count, int *__restrict__ err)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int64_t i = first_slot + a;
    double s[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double r = buf[(size_t)c * n + a];
        A.r[c][i] = r;
        s[c] = __ddiv_rn(r, L);
        A.s[c][i] = s[c];
        A.v[c][i] = buf[(size_t)(3 + c) * n + a];
    }
    A.hs[i] = buf[(size_t)6 * n + a];
    A.ts[i] = buf[(size_t)7 * n + a];
    A.mass[i] = buf[(size_t)8 * n + a];
    A.id[i] = __double2loint(buf[(size_t)9 * n + a]);
    A.type[i] = __double2hiint(buf[(size_t)9 * n + a]);
    A.xbase[i] = (int32_t)__double_as_longlong(buf[(size_t)10 * n + a]);
    A.xmask[i] = (uint64_t)__double_as_longlong(buf[(size_t)11 * n + a]);
    const int x = cell_coord(s[0], M), y = cell_coord(s[1], M), z = cell_coord(s[2], M);
    gcell[i] = x + M * (y + M * z);
    int dz = z - z0;
    if (dz < 0) dz += M;
    if (dz >= nz) { atomicCAS(err, 0, 1); dz = nz - 1; }
    const int lc = x + M * (y + M * (dz + R));
    lcell[i] = lc;
    atomicAdd(count + lc, 1);
}

// Ghost atoms at a re-binning: everything the force kernels need of a boundary atom, packed field-major into ONE message per
// neighbour (field f of atom a at buf[f*n + a]) -- nine separate array slices per side cost 36 NCCL operations per re-binning.
#define GHOST_FIELDS 8   // doubles per ghost atom: s(3) hs ts | (id,type) | xbase | xmask
__global__ void k_pack_ghosts(int64_t first, int n, AtomArrays A, double *__restrict__ buf)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int64_t i = first + a;
#pragma unroll
    for (int c = 0; c < 3; c++) buf[(size_t)c * n + a] = A.s[c][i];
    buf[(size_t)3 * n + a] = A.hs[i];
    buf[(size_t)4 * n + a] = A.ts[i];
    buf[(size_t)5 * n + a] = __hiloint2double(A.type[i], A.id[i]);
    buf[(size_t)6 * n + a] = __longlong_as_double((long long)A.xbase[i]);
    buf[(size_t)7 * n + a] = __longlong_as_double((long long)A.xmask[i]);
}
__global__ void k_unpack_ghosts(int64_t first, int n, const double *__restrict__ buf, AtomArrays A)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const int64_t i = first + a;
#pragma unroll
    for (int c = 0; c < 3; c++) A.s[c][i] = buf[(size_t)c * n + a];
    A.hs[i] = buf[(size_t)3 * n + a];
    A.ts[i] = buf[(size_t)4 * n + a];
    A.id[i] = __double2loint(buf[(size_t)5 * n + a]);
    A.type[i] = __double2hiint(buf[(size_t)5 * n + a]);
    A.xbase[i] = (int32_t)__double_as_longlong(buf[(size_t)6 * n + a]);
    A.xmask[i] = (uint64_t)__double_as_longlong(buf[(size_t)7 * n + a]);
}

// Scatter that skips the trash cell (leavers / foreign atoms).
__global__ void k_scatter_slab(int64_t first, int64_t n, const int32_t *__restrict__ lcell, int ncell,
                               const int32_t *__restrict__ cell_start, int32_t *__restrict__ fill,
                               int32_t *__restrict__ order)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = first + k;
    const int c = lcell[i];
    if (c >= ncell) return;
    order[cell_start[c] + atomicAdd(fill + c, 1)] = (int32_t)i;
}
