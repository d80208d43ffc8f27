// pairs14.cuh -- 1-4 scaling of the Lennard-Jones interaction (lj14scale).
//
// The reference parses lj14scale from the force-field file (src/modelling.jl:199; test/data/dibenzo-p-dioxin-in-water.xml:84
// has 0.5) and never applies it; the oracle pins the meaning (oracle_pairs14_correction): a pair of atoms three bonds
// apart interacts with lj14scale times the energy / virial / force of an ordinary pair.  The force kernels evaluate those pairs
// at full strength like any other (they are a handful per molecule, known from the topology, and tagging them in the pair
// list would cost every pair of every system an instruction); this kernel then adds (scale - 1) x the interaction of every
// listed pair that is inside the cutoff: one thread per pair, the oracle's exact rounding sequence for the cutoff decision
// (min_image_r2) and interaction() itself (lj_interaction), FP64 atomics for the two atoms' sums.
#pragma once
#include "lj_pair.cuh"

struct Pairs14Args {
    int64_t n;
    const int32_t *ij;            // (n, 2) global atom ids, i < j
    const int32_t *slot_of_id;
    const double *sx, *sy, *sz, *hs, *ts;
    double L, cm1;                // box edge, scale - 1
    LJModel model;
    int bitmask;
    double *fx, *fy, *fz, *en, *vir;
};

__global__ void k_pairs14(Pairs14Args a)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n) return;
    const int si = a.slot_of_id[a.ij[2 * k]], sj = a.slot_of_id[a.ij[2 * k + 1]];
    double vx, vy, vz;
    const double r2 = min_image_r2(a.sx[si], a.sy[si], a.sz[si], a.sx[sj], a.sy[sj], a.sz[sj], a.L, vx, vy, vz);
    if (!(r2 <= a.model.rc2)) return;
    const double inv = __ddiv_rn(1.0, r2);
    double Eg, Wg;
    lj_interaction(r2, inv, a.hs[si] + a.hs[sj], a.ts[si] * a.ts[sj], a.model, 60.0 * a.model.id2, Eg, Wg);
    if (a.bitmask & EMDEE_FORCES) {
        const double q = a.cm1 * (Wg * inv);
        atomicAdd(a.fx + si, q * vx); atomicAdd(a.fy + si, q * vy); atomicAdd(a.fz + si, q * vz);
        atomicAdd(a.fx + sj, -q * vx); atomicAdd(a.fy + sj, -q * vy); atomicAdd(a.fz + sj, -q * vz);
    }
    if (a.bitmask & EMDEE_ENERGIES) { atomicAdd(a.en + si, 0.5 * a.cm1 * Eg); atomicAdd(a.en + sj, 0.5 * a.cm1 * Eg); }
    if (a.bitmask & EMDEE_VIRIALS) { atomicAdd(a.vir + si, 0.5 * a.cm1 * Wg); atomicAdd(a.vir + sj, 0.5 * a.cm1 * Wg); }
}
