// lj_pair.cuh -- FP64 pair geometry and the Lennard-Jones pair function for sm_100a.
//
// Restates src/nonbonded.jl:40,60-61,70-71,74 (minimum image, r2, force from virial) and
// src/lennard_jones.jl:25-42 (interaction).  Integer-valued decisions (which pairs are inside the
// cutoff, which branch of the clamp on line 37 applies) use exactly the rounding sequence of the
// oracle; the remaining arithmetic is free to use fused multiply-adds and one shared reciprocal
// (tolerance 1e-10 on E/W, 1e-9 F_rms on forces -- BASELINE.json north_star).
#pragma once
#include "common.cuh"

// rint() (ties-to-even) for |d| < 2^51 as two FP64-pipe additions instead of a conversion-pipe
// FRND: adding and subtracting 1.5*2^52 rounds to an integer in the current (nearest-even) mode.
// __dadd_rn is never re-associated or contracted by the compiler.
__device__ __forceinline__ double rint_magic(double d)
{
    const double K = 6755399441055744.0;  // 1.5 * 2^52
    return __dadd_rn(__dadd_rn(d, K), -K);
}

// Minimum-image separation and its square in the pinned order (oracle dist2()):
//   d = s_i - s_j ; d -= rint(d) ; v = L*d ; r2 = fma(vz,vz, fma(vy,vy, vx*vx))
// Bit-exact with the oracle, so `r2 <= rc2` selects the same pair set (src/cells.jl:241,246,260).
__device__ __forceinline__ double min_image_r2(double six, double siy, double siz, double sjx, double sjy,
                                               double sjz, double L, double &vx, double &vy, double &vz)
{
    double dx = __dsub_rn(six, sjx), dy = __dsub_rn(siy, sjy), dz = __dsub_rn(siz, sjz);
    dx = __dsub_rn(dx, rint_magic(dx));
    dy = __dsub_rn(dy, rint_magic(dy));
    dz = __dsub_rn(dz, rint_magic(dz));
    vx = __dmul_rn(L, dx);
    vy = __dmul_rn(L, dy);
    vz = __dmul_rn(L, dz);
    return __fma_rn(vz, vz, __fma_rn(vy, vy, __dmul_rn(vx, vx)));
}

// 1/a for a normal, positive a: MUFU.RCP64H seed (about 20 bits) + two Newton steps on the FP64
// pipe; error about 1 ulp.  r2 of a real pair is far from the denormal/overflow ranges the
// compiler's generic division path guards against.
__device__ __forceinline__ double rcp_fast(double a)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    e = fma(-a, y, 1.0);
    return fma(y, e, y);
}

// interaction(r2, model, atom_i, atom_j) -- src/lennard_jones.jl:25-42.
//   sig = half_sigma_i + half_sigma_j (:29), tt = twice_sqrt_eps_i * twice_sqrt_eps_j (:33),
//   inv_r2 = 1/r2 (the reference divides twice: :31 and src/nonbonded.jl:74).
// Returns E*g and -r d(E*g)/dr.  The clamp on :37, x *= 0.5(sign(x) - sign(x-1)), is reproduced
// branch by branch on the same x = (r2 - rs2)*id2: x<0 -> 0, x>1 -> 0 (full LJ beyond the cutoff,
// SURVEY F4), x==1 -> 0.5.
__device__ __forceinline__ void lj_interaction(double r2, double inv_r2, double sig, double tt, const LJModel &m,
                                               double c60id2, double &Eg, double &Wg)
{
    double s2 = sig * sig * inv_r2;                 // :31
    double s6 = s2 * s2 * s2;                       // :32
    double e4s6 = tt * s6;                          // :33
    double E = fma(e4s6, s6, -e4s6);                // :34  e4s6*(s6-1)
    double W0 = e4s6 * fma(12.0, s6, -6.0);         // :35  6 e4s6 (2 s6 - 1)
    double x = __dmul_rn(__dsub_rn(r2, m.rs2), m.id2);   // :36
    x = (x < 0.0 || x > 1.0) ? 0.0 : (x == 1.0 ? 0.5 : x);   // :37
    double x2 = x * x;                              // :38
    double p = fma(-6.0, x2, fma(15.0, x, -10.0));
    double g = fma(x * x2, p, 1.0);                 // :39  1 + x^3 (15x - 6x^2 - 10)
    double t = x * (1.0 - x);                       //      x^2 (1 - 2x + x^2) = (x(1-x))^2
    double G = (t * t) * c60id2 * r2;               // :40  60 x^2 (1-x)^2 id2 r2
    Eg = E * g;                                     // :41
    Wg = fma(W0, g, E * G);
}

// ---- the stepping path's pair function -------------------------------------------------------------
// Same quantities as lj_interaction(), re-associated for the FP64 pipe (the tolerance is 1e-10 on E/W and
// 1e-9 F_rms on forces, not bit equality; the pair SET stays bit-exact, see pair_in_range below):
//   * one reciprocal: MUFU.RCP64H seed (input mantissa truncated to 20 bits, so e <= 2^-20) and ONE cubic
//     correction y(1 + e + e^2), error e^3 ~ 1e-18 -- three DFMAs instead of four;
//   * sigma^2 and the epsilon product come from the class table;
//   * x = fma(r2, id2, -rs2*id2); x < 0 is clamped with an integer test of the sign bit (ALU pipe, not a
//     DSETP); the x >= 1 branches (x == 1 gives 0.5, x > 1 gives 0) can only be reached within rounding of
//     rc2, where the exact slow path evaluates the oracle's own clamped x and passes it in (xover, xval);
//   * q = W/r2 = W0*g/r2 + E*G/r2 with G/r2 = 60 id2 (x(1-x))^2, so r2 cancels.
// 22 FP64-pipe instructions for q (+ 6 for the geometry, + 3 to accumulate the force).
struct LJFast {
    double id2, nrs2id2, c60id2;   // 1/(rc2-rs2), -rs2*id2, 60*id2
};
template <bool EW>
__device__ __forceinline__ double lj_pair_q(double r2, double sig2, double tt, const LJFast &m, bool xover, double xval,
                                            double &Eg, double &Wg)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
    const double e = fma(-r2, y, 1.0);
    const double inv = fma(y, fma(e, e, e), y);
    const double s2 = sig2 * inv;                    // :31
    const double s6 = s2 * s2 * s2;                  // :32
    const double e4s6 = tt * s6;                     // :33
    const double E = fma(e4s6, s6, -e4s6);           // :34
    const double W0 = e4s6 * fma(12.0, s6, -6.0);    // :35
    double x = fma(r2, m.id2, m.nrs2id2);            // :36
    x = __double2hiint(x) < 0 ? 0.0 : x;             // :37, x < 0 branch
    x = xover ? xval : x;                            // :37 as the oracle evaluated it (borderline pairs only)
    const double x2 = x * x;                         // :38
    const double g = fma(x * x2, fma(-6.0, x2, fma(15.0, x, -10.0)), 1.0);   // :39
    const double t = x - x2;                         // x(1-x)
    const double q = fma(E * m.c60id2, t * t, (W0 * g) * inv);               // (:35*g + E*:40)/r2
    if (EW) { Eg = E * g; Wg = q * r2; }             // :41
    return q;
}

// Cutoff decision from a local-frame r2 (rounding differs from the oracle's by ~1e-15 relative):
//   hi word of r2 well below / above that of rc2  -> decided here with two integer compares;
//   within 3 * 2^-20 of rc2 (a few hundred pairs per million-atom step) -> the caller evaluates the oracle's
//   exact rounding sequence (min_image_r2 on the scaled coordinates) and decides on that.
// Returns -1 inside, +1 outside, 0 borderline.
__device__ __forceinline__ int pair_in_range(double r2, int rc2hi)
{
    const int t = __double2hiint(r2) - (rc2hi - 1);
    return t < 0 ? -1 : ((unsigned)t <= 2u ? 0 : 1);
}

// Exclusion test: bit (j - base_i) of mask_i over the window [base_i, base_i+64).
__device__ __forceinline__ bool pair_excluded(int32_t base_i, uint64_t mask_i, int32_t id_j)
{
    uint32_t o = (uint32_t)(id_j - base_i);
    return o < 64u && ((mask_i >> o) & 1ull);
}

// splitmix64 finaliser: pair hash of the audit digest.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t pair_hash(int32_t i, int32_t j)  // i<j, 0-based global ids
{
    return mix64(((uint64_t)(uint32_t)i << 32) | (uint64_t)(uint32_t)j);
}

__device__ __forceinline__ double shfl_f64(double v, int src)
{
    return __shfl_sync(0xffffffffu, v, src);
}
__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
