/* abi_example.c -- the C ABI of include/emdee_b200.h used from plain C (what any FFI binds: Julia ccall, ctypes, cgo).
 *
 *   gcc -std=c11 -Iinclude examples/abi_example.c -o abi_example -Lemdee.jl_b200/csrc -lemdee_b200 -Wl,-rpath,$PWD/emdee.jl_b200/csrc
 *
 * Mirrors `compute_nonbonded!(forces, energies, virials, positions, L, tiles, model, atoms, Val(FORCES|ENERGIES|VIRIALS))`
 * (src/nonbonded.jl:109-120) on a small cubic lattice with the one-shot host-array entry point, then the same system
 * through a device-resident handle in CUTOFF mode.  Without a B200 every computing call fails with EMDEE_ERR_CUDA and the
 * program reports that (exit code 3) -- there is no CPU fallback. */
#include <stdio.h>
#include <stdlib.h>

#include "emdee_b200.h"

#define CHECK(call)                                                                          \
    do {                                                                                     \
        int st_ = (call);                                                                    \
        if (st_ != EMDEE_OK) {                                                               \
            fprintf(stderr, "%s -> status %d: %s\n", #call, st_, emdee_last_error());        \
            return st_ == EMDEE_ERR_CUDA ? 3 : 1;                                            \
        }                                                                                    \
    } while (0)

int main(void)
{
    enum { n = 6, N = n * n * n };
    const double L = 1.2 * n, cutoff = 2.5, sw = 2.0;
    double *pos = malloc(sizeof(double) * 3 * N), *atoms = malloc(sizeof(double) * 2 * N);
    double *f = malloc(sizeof(double) * 3 * N), *e = malloc(sizeof(double) * N), *w = malloc(sizeof(double) * N);
    if (!pos || !atoms || !f || !e || !w) return 2;
    for (int i = 0; i < N; i++) {
        pos[3 * i + 0] = 1.2 * (i % n) + 0.01 * (i % 7);
        pos[3 * i + 1] = 1.2 * ((i / n) % n) + 0.01 * (i % 5);
        pos[3 * i + 2] = 1.2 * (i / (n * n)) + 0.01 * (i % 3);
        atoms[2 * i + 0] = 0.5;   /* LennardJonesAtom(1, 1) = LJAtom(sigma/2, 2 sqrt(eps)), src/lennard_jones.jl:13 */
        atoms[2 * i + 1] = 2.0;
    }
    printf("libemdee_b200 version %d\n", emdee_version());

    /* one-shot form: the reference's call with host arrays (default tile list, the reference's all-pairs semantics) */
    CHECK(emdee_compute_nonbonded_host(N, pos, L, cutoff, sw, atoms, NULL, 0, EMDEE_ALLPAIRS_REFERENCE, 2,
                                       EMDEE_FORCES | EMDEE_ENERGIES | EMDEE_VIRIALS, f, e, w));
    double E = 0;
    for (int i = 0; i < N; i++) E += e[i];
    printf("all pairs: sum of per-atom energies %.12g\n", E);

    /* handle form: data stays on the GPU; cell list + cutoff pair set */
    emdee_ctx *ctx = NULL;
    emdee_system *sys = NULL;
    CHECK(emdee_create(&ctx, 0));
    CHECK(emdee_system_create(ctx, N, L, &sys));
    CHECK(emdee_set_model(sys, cutoff, sw));
    CHECK(emdee_set_lj_atoms(sys, atoms));
    CHECK(emdee_set_positions(sys, pos));
    CHECK(emdee_bin(sys, 1));
    CHECK(emdee_compute_nonbonded(sys, EMDEE_CUTOFF, EMDEE_FORCES | EMDEE_ENERGIES | EMDEE_VIRIALS));
    double W = 0;
    int64_t npairs = 0;
    CHECK(emdee_get_totals(sys, &E, &W, &npairs));
    printf("cutoff: E %.12g  W %.12g  pairs %lld\n", E, W, (long long)npairs);
    /* the reference's call shape on the resident system: the three output arrays are host memory, filled in one call */
    CHECK(emdee_compute_nonbonded_into(sys, EMDEE_CUTOFF, EMDEE_FORCES | EMDEE_ENERGIES | EMDEE_VIRIALS, f, e, w));
    E = 0;
    for (int64_t i = 0; i < N; i++) E += e[i];
    printf("cutoff, host arrays: sum of per-atom energies %.12g\n", E);
    CHECK(emdee_system_destroy(sys));
    CHECK(emdee_destroy(ctx));
    free(pos); free(atoms); free(f); free(e); free(w);
    return 0;
}
