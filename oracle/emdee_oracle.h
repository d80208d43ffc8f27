/* emdee_oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE ONLY, see emdee_oracle.c). */
#ifndef EMDEE_ORACLE_H
#define EMDEE_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int oracle_num_threads(void);
void oracle_set_num_threads(int n);
uint64_t oracle_mix64(uint64_t z);

/* src/lennard_jones.jl:6-18,25-42 */
void oracle_lj_model_f64(double cutoff, double sw, double out[3]);
void oracle_lj_model_f32(double cutoff, double sw, float out[3]);
void oracle_lj_atom_f64(double eps, double sigma, double out[2]);
void oracle_lj_atom_f32(double eps, double sigma, float out[2]);
void oracle_interaction_f64(double r2, const double model[3], const double ai[2], const double aj[2], double out[2]);
void oracle_interaction_f32(float r2, const float model[3], const float ai[2], const float aj[2], float out[2]);

/* src/nonbonded.jl:18-26 */
int64_t oracle_tiles(int64_t N, int32_t *out);

/* src/nonbonded.jl:122-155 (ALLPAIRS_REFERENCE, naive order) */
void oracle_naive_allpairs_f64(int64_t N, const double *pos, double L, const double model[3], const double *atoms,
                               double *forces, double *energies, double *virials);
void oracle_naive_allpairs_f32(int64_t N, const float *pos, float L, const float model[3], const float *atoms,
                               float *forces, float *energies, float *virials);

/* src/nonbonded.jl:44-120 (ALLPAIRS_REFERENCE, tile order) */
void oracle_tiles_allpairs_f64(int64_t N, const double *pos, double L, const int32_t *tiles, int64_t ntiles,
                               const double model[3], const double *atoms, int bitmask,
                               double *forces, double *energies, double *virials);
void oracle_tiles_allpairs_f32(int64_t N, const float *pos, float L, const int32_t *tiles, int64_t ntiles,
                               const float model[3], const float *atoms, int bitmask,
                               float *forces, float *energies, float *virials);

/* src/cells.jl:36,79-85 */
int32_t oracle_cells_per_dimension(double L, double cutoff, int ndiv);
void oracle_cell_index_f64(int64_t N, const double *pos, double L, int32_t M, int32_t *index);
void oracle_cell_index_f32(int64_t N, const float *pos, float L, int32_t M, int32_t *index);

/* CUTOFF mode: src/cells.jl:240-270 pair set, src/lennard_jones.jl pair math */
int64_t oracle_pair_set_brute(int64_t N, const double *pos, double L, double rc2,
                              const int32_t *excl_base, const uint64_t *excl_mask,
                              int32_t *ij, int64_t cap, uint64_t digest[3]);
int64_t oracle_pair_set_cells(int64_t N, const double *pos, double L, double cutoff, int ndiv,
                              const int32_t *excl_base, const uint64_t *excl_mask, int32_t *ij, int64_t cap);
int oracle_cutoff_cells(int64_t N, const double *pos, double L, double cutoff, double sw, const double *atoms,
                        int ndiv, const int32_t *excl_base, const uint64_t *excl_mask, int bitmask,
                        double *forces, double *energies, double *virials,
                        double totals[2], int64_t *npairs, uint64_t digest[3]);

/* 1-4 scaling as a correction (lj14scale, src/modelling.jl:199; never applied by the reference) */
int64_t oracle_pairs14_correction(int64_t N, const double *pos, double L, double cutoff, double sw, const double *atoms,
                                  const int32_t *ij, int64_t n14, double scale, int bitmask,
                                  double *forces, double *energies, double *virials, double dtot[2]);

/* velocity-Verlet (defined by the oracle) */
int oracle_vv_steps(int64_t N, double *pos, double *vel, double *forces, const double *mass, double L,
                    double cutoff, double sw, const double *atoms, int ndiv,
                    const int32_t *excl_base, const uint64_t *excl_mask, double dt, int64_t nsteps);

#ifdef __cplusplus
}
#endif
#endif
