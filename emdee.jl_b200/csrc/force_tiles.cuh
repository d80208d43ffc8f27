// force_tiles.cuh -- ALLPAIRS_REFERENCE mode: the reference's own 32x32 shuffle-tile algorithm
// (compute_tile!, src/nonbonded.jl:44-107) restated in FP64 with tail masking.
//
// One warp per tile (I-block, J-block); lane l owns row atom I=32(bI-1)+l and column atom
// J=32(bJ-1)+l.  In iteration m lane l meets column atom of lane (l+m)%32 (:68), evaluates the pair
// and hands the reaction back to that lane's J accumulators through lane (l+32-m)%32 (:69,76,80,84).
// Diagonal tiles run 31 iterations and drop the J side (:67,96).  Epilogue: FP64 atomics of f,
// E/2, W/2 per atom (:88-104).  With CULL the same kernel serves EMDEE_CUTOFF for boxes too small
// for a cell grid (pair kept iff r2 <= rc2 and not excluded).
#pragma once
#include "lj_pair.cuh"

struct TileArgs {
    const int32_t *tiles;   // ntiles x (I,J), 1-based block ids (src/nonbonded.jl:18-26)
    int64_t ntiles;
    int64_t N;              // number of atoms (ids 0..N-1)
    const int32_t *slot_of_id;  // nullptr: slot == id
    const double *sx, *sy, *sz; // scaled positions by slot
    const double *hs, *ts;
    const int32_t *id, *xbase;
    const uint64_t *xmask;
    double *fx, *fy, *fz, *en, *vir;   // outputs by slot (zeroed by the launcher, :112-114)
    unsigned long long *digest;        // CULL only: {accepted pairs, sum hash, xor hash, pair-list cursor}
    int32_t *pairs;                    // CULL only: optional pair list (2 x pair_cap)
    long long pair_cap;
    double L;
    LJModel model;
};

template <bool F, bool E, bool W, bool CULL, bool EXCL>
__global__ void __launch_bounds__(128) k_force_tiles(TileArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= a.ntiles) return;
    const int bI = a.tiles[2 * tile], bJ = a.tiles[2 * tile + 1];
    const bool diag = (bI == bJ);
    const int64_t I = (int64_t)(bI - 1) * 32 + lane, J = (int64_t)(bJ - 1) * 32 + lane;
    const bool vI = I < a.N, vJ = J < a.N;
    const int si = vI ? (a.slot_of_id ? a.slot_of_id[I] : (int)I) : 0;
    const int sj = vJ ? (a.slot_of_id ? a.slot_of_id[J] : (int)J) : 0;

    const double xi = a.sx[si], yi = a.sy[si], zi = a.sz[si], hsi = a.hs[si], tsi = a.ts[si];
    const double xj = a.sx[sj], yj = a.sy[sj], zj = a.sz[sj], hsj = a.hs[sj], tsj = a.ts[sj];
    int32_t xb = 0; uint64_t xm = 0;
    if (EXCL) { xb = a.xbase[si]; xm = a.xmask[si]; }
    const int32_t idj = (int32_t)J;
    const double c60id2 = 60.0 * a.model.id2;

    double fix = 0, fiy = 0, fiz = 0, ei = 0, wi = 0;
    double fjx = 0, fjy = 0, fjz = 0, ej = 0, wj = 0;
    unsigned long long npairs = 0, hsum = 0, hxor = 0;

    const int niter = 32 - (diag ? 1 : 0);
    for (int m = 1; m <= niter; m++) {
        const int pj = (lane + m) & 31;        // :68 partner column lane
        const int pk = (lane + 32 - m) & 31;   // :69 lane whose partner is this lane
        const double pxj = shfl_f64(xj, pj), pyj = shfl_f64(yj, pj), pzj = shfl_f64(zj, pj);
        const double phs = shfl_f64(hsj, pj), pts = shfl_f64(tsj, pj);
        const bool pv = __shfl_sync(0xffffffffu, (int)vJ, pj) != 0;
        const int32_t pid = __shfl_sync(0xffffffffu, idj, pj);
        double vx, vy, vz, Eg = 0, Wg = 0, qx = 0, qy = 0, qz = 0;
        bool ok = vI && pv;
        if (ok) {
            const double r2 = min_image_r2(xi, yi, zi, pxj, pyj, pzj, a.L, vx, vy, vz);   // :70-71
            if (CULL && !(r2 <= a.model.rc2)) ok = false;
            if (EXCL && ok && pair_excluded(xb, xm, pid)) ok = false;
            if (ok) {
                const double inv = rcp_fast(r2);
                lj_interaction(r2, inv, hsi + phs, tsi * pts, a.model, c60id2, Eg, Wg);   // :72
                const double q = Wg * inv;                                                // :74
                qx = q * vx; qy = q * vy; qz = q * vz;
                if (CULL && (!diag || (int32_t)I < pid)) {
                    const int32_t lo = min((int32_t)I, pid), hi = max((int32_t)I, pid);
                    const uint64_t hh = pair_hash(lo, hi);
                    npairs++; hsum += hh; hxor ^= hh;
                    if (a.pairs) {
                        const unsigned long long k = atomicAdd(a.digest + 3, 1ull);
                        if ((long long)k < a.pair_cap) { a.pairs[2 * k] = lo; a.pairs[2 * k + 1] = hi; }
                    }
                }
            }
        }
        if (F) {
            fix += qx; fiy += qy; fiz += qz;                                              // :75
            fjx -= shfl_f64(qx, pk); fjy -= shfl_f64(qy, pk); fjz -= shfl_f64(qz, pk);    // :76
        }
        if (E) { ei += Eg; ej += shfl_f64(Eg, pk); }                                      // :79-80
        if (W) { wi += Wg; wj += shfl_f64(Wg, pk); }                                      // :83-84
    }
    if (vI) {                                                                             // :88-94
        if (F) { atomicAdd(a.fx + si, fix); atomicAdd(a.fy + si, fiy); atomicAdd(a.fz + si, fiz); }
        if (E) atomicAdd(a.en + si, 0.5 * ei);
        if (W) atomicAdd(a.vir + si, 0.5 * wi);
    }
    if (!diag && vJ) {                                                                    // :96-104
        if (F) { atomicAdd(a.fx + sj, fjx); atomicAdd(a.fy + sj, fjy); atomicAdd(a.fz + sj, fjz); }
        if (E) atomicAdd(a.en + sj, 0.5 * ej);
        if (W) atomicAdd(a.vir + sj, 0.5 * wj);
    }
    if (CULL) {
        for (int o = 16; o > 0; o >>= 1) {
            npairs += __shfl_xor_sync(0xffffffffu, npairs, o);
            hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
            hxor ^= __shfl_xor_sync(0xffffffffu, hxor, o);
        }
        if (lane == 0 && npairs) { atomicAdd(a.digest, npairs); atomicAdd(a.digest + 1, hsum); atomicXor(a.digest + 2, hxor); }
    }
}
