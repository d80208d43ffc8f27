// common.cuh -- shared declarations of libemdee_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/emdee_b200.h"

#define EMDEE_WARP 32

// -DEMDEE_CHECKS=1: explicit bounds checks on the hand-rolled shared-memory structures of the list kernels (staged indices,
// per-lane stacks and rows, task indices); a violation raises device flag 8, which the next getter / emdee_synchronize reports.
// compute-sanitizer is not available on the GPU pool this library is developed on; the GPU tests are run against a library
// built this way instead (tools/gpu_checked.sh).  Off by default: the checks cost instructions in the hot loops.
#ifndef EMDEE_CHECKS
#define EMDEE_CHECKS 0
#endif
#if EMDEE_CHECKS
#define EMDEE_CHECK(cond, errptr) do { if (!(cond)) atomicCAS((errptr), 0, 8); } while (0)
#else
#define EMDEE_CHECK(cond, errptr) do { } while (0)
#endif

// ---- error handling: thread-local message + status code -------------------------------------
void emdee_set_error(const char *fmt, ...);
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            emdee_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return EMDEE_ERR_CUDA;                                                                  \
        }                                                                                           \
    } while (0)
#define EMDEE_TRY(expr)                   \
    do {                                  \
        int _s = (expr);                  \
        if (_s != EMDEE_OK) return _s;    \
    } while (0)
#define EMDEE_FAIL(code, ...)             \
    do {                                  \
        emdee_set_error(__VA_ARGS__);     \
        return (code);                    \
    } while (0)

// ---- model constants handed to kernels by value (LennardJonesModel, src/lennard_jones.jl:6-11) --
struct LJModel {
    double rc2, rs2, id2;  // cutoff^2, switch^2, 1/(rc2-rs2)
};

// Per-atom state in the CURRENT slot order (id order before the first binning, (cell,id) order after).
// Structure of arrays; every array has `cap` entries.  Slots [0,nlo) are lower ghosts, [nlo,nlo+nown)
// owned atoms, [nlo+nown, ntot) upper ghosts (ghosts exist only in a slab decomposition).
struct AtomArrays {
    double *r[3] = {nullptr, nullptr, nullptr};   // positions
    double *s[3] = {nullptr, nullptr, nullptr};   // scaled positions s = r/L (src/nonbonded.jl:60-61)
    double *v[3] = {nullptr, nullptr, nullptr};   // velocities
    double *rb[3] = {nullptr, nullptr, nullptr};  // positions at the last binning (skin check)
    double *hs = nullptr, *ts = nullptr;          // LJAtom.half_sigma, LJAtom.twice_sqrt_eps
    double *mass = nullptr;
    int32_t *id = nullptr;                        // global atom id of the slot
    int32_t *type = nullptr;                      // LJ parameter class (index into the pair table)
    int32_t *xbase = nullptr;                     // exclusion window base (global id)
    uint64_t *xmask = nullptr;                    // exclusion bits
};

// Geometry of the local cell grid handed to the force kernel.
struct GridDesc {
    int M;        // cells per dimension of the global grid (src/cells.jl:36)
    int R;        // neighbour reach in cells
    int nzt;      // local z planes including ghost planes
    int zhome0;   // first home plane (local index)
    int nzhome;   // number of home planes
    int zwrap;    // 1: local grid is the whole periodic grid (single GPU), 0: explicit ghost planes
    int zglob0;   // global z of local plane 0 (may be negative)
    int bx, by, bz;     // home brick shape in cells
    int nbx, nby, nbz;  // bricks per dimension
};

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
