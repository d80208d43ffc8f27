/* emdee_b200.h -- C ABI of libemdee_b200.so, the B200 (sm_100a) implementation of EmDee.jl's
 * nonbonded hot path: cell binning -> cutoff pair loop -> Lennard-Jones energy/force/virial
 * (-> velocity-Verlet).
 *
 * The reference has no FFI seam on this path: it is exported Julia functions operating on CuArrays
 * (SURVEY section 8b).  Each entry point below is what a Julia shim keeping the reference's names
 * would bind with `ccall` (see INTEGRATION.md and julia/EmDee.jl); the reference symbol replaced is
 * cited as file:line relative to the reference tree.  The author's own ccall style (library
 * constant, Ptr{T} arrays, Cint sizes) is visible at src/molecular_graphs.jl:4,73-80.
 *
 * Conventions
 *   - plain C types only; every function returns an int status (EMDEE_OK == 0);
 *     emdee_last_error() returns the thread's last message.
 *   - all floating point is double (the reference is Float32; FP64 is the north-star precision).
 *   - host arrays are caller-owned and caller-sized; positions/velocities/forces are 3xN
 *     column-major (x,y,z of atom 0, then atom 1, ...) exactly like the reference's 3xN matrices
 *     (src/nonbonded.jl:52-61); LJ atoms are 2xN {half_sigma, twice_sqrt_eps} like `LJAtom`
 *     (src/lennard_jones.jl:15-18).  Atom ids at the boundary are 0-based positions in those arrays.
 *   - device memory is library-owned behind the opaque handles.  There is NO CPU fallback: every
 *     entry point that computes fails with EMDEE_ERR_CUDA when no sm_100 device is usable.
 *   - one context per process and GPU; calls on one context are not re-entrant.
 */
#ifndef EMDEE_B200_H
#define EMDEE_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct emdee_ctx emdee_ctx;
typedef struct emdee_system emdee_system;

enum {
    EMDEE_OK = 0,
    EMDEE_ERR_INVALID = 1,   /* bad argument (N<=0, L<=0, not 0<rs<rc, rc>L/2, null pointer, ...) */
    EMDEE_ERR_CUDA = 2,      /* CUDA runtime error or no usable device */
    EMDEE_ERR_NCCL = 3,      /* NCCL error */
    EMDEE_ERR_STATE = 4,     /* call order (e.g. compute before positions/model are set) */
    EMDEE_ERR_CAPACITY = 5,  /* caller buffer or internal capacity too small */
    EMDEE_ERR_SKIN = 6       /* an atom moved more than skin/2 since the last binning */
};

/* Output selection, src/nonbonded.jl:12-14 */
enum { EMDEE_FORCES = 1, EMDEE_ENERGIES = 2, EMDEE_VIRIALS = 4 };

/* Pair-set semantics.
 * EMDEE_CUTOFF: pairs i<j with r2 <= rc2 (predicate of src/cells.jl:241,246,260), cell-list kernels.
 * EMDEE_ALLPAIRS_REFERENCE: every minimum-image pair, bug-compatible with the shipped tile kernel
 *   (src/nonbonded.jl:44-107: no cull, full LJ beyond the cutoff). */
enum { EMDEE_CUTOFF = 0, EMDEE_ALLPAIRS_REFERENCE = 1 };

int emdee_version(void);
const char *emdee_last_error(void);

/* ---- context: one GPU, its streams, optionally one rank of a slab decomposition ---- */
int emdee_create(emdee_ctx **ctx, int device);
int emdee_destroy(emdee_ctx *ctx);
/* Slab decomposition over `nranks` processes (one per GPU): rank 0 creates a 128-byte id with
 * emdee_comm_unique_id, the host runtime broadcasts it, every rank calls emdee_comm_init. */
int emdee_comm_unique_id(char id[128]);
int emdee_comm_init(emdee_ctx *ctx, int rank, int nranks, const char id[128]);
int emdee_device_info(emdee_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, int64_t *mem_bytes);
/* Measured FP64 FMA throughput of this GPU (FLOP/s, 2 per DFMA) -- the roofline denominator. */
int emdee_measure_fp64_peak(emdee_ctx *ctx, double *flops_per_s, double *ms);

/* ---- system: N atoms in a cubic periodic box of edge L (scalar L: src/nonbonded.jl:60,70) ---- */
int emdee_system_create(emdee_ctx *ctx, int64_t N, double L, emdee_system **sys);
int emdee_system_destroy(emdee_system *sys);
/* LennardJonesModel(cutoff, switch): src/lennard_jones.jl:6-11 */
int emdee_set_model(emdee_system *sys, double cutoff, double switch_dist);
/* Vector{LJAtom} from LennardJonesAtom(eps, sigma): src/lennard_jones.jl:13-18 */
int emdee_set_lj_atoms(emdee_system *sys, const double *half_sigma_twice_sqrt_eps_2xN);
/* positions argument of compute_nonbonded! / Cells / update_cells!: src/nonbonded.jl:109, src/cells.jl:176,196 */
int emdee_set_positions(emdee_system *sys, const double *pos_3xN);
int emdee_set_velocities(emdee_system *sys, const double *vel_3xN);
int emdee_set_masses(emdee_system *sys, const double *mass_N);
/* Intramolecular exclusions as a bitmask (adjacency source: src/modelling.jl:19-23,297-304):
 * pair (i,j) is excluded iff 0 <= j-base[i] < 64 and bit (j-base[i]) of mask[i] is set. NULL clears. */
int emdee_set_exclusions(emdee_system *sys, const int32_t *base_N, const uint64_t *mask_N);
/* 1-4 scaling: lj14scale of the force field (src/modelling.jl:199, test/data/dibenzo-p-dioxin-in-water.xml:84; parsed and
 * never applied by the reference).  Pairs (i<j, global ids) three bonds apart interact with `scale` times an ordinary pair's
 * energy / virial / force in EMDEE_CUTOFF evaluations and in emdee_vv_step.  NULL or n = 0 clears.  Single GPU. */
int emdee_set_pairs14(emdee_system *sys, const int32_t *ij_2xn, int64_t n, double scale);
/* Extra cell-edge margin so that binning stays valid while atoms move < skin/2 (0 = reference cells). */
int emdee_set_skin(emdee_system *sys, double skin);

/* Cells(r, L, cutoff; ndiv) and update_cells!(cells, r, L): src/cells.jl:176-194,196-222.
 * M = floor(ndiv*L/(cutoff+skin)) (src/cells.jl:36), cell index of src/cells.jl:79-85, atoms sorted by
 * (cell, id).  Replaces distribute!/clean_cells!/collect_baskets!/renew_cells! (src/cells.jl:46-174). */
int emdee_bin(emdee_system *sys, int ndiv);
/* update_cells!(cells, r, L), src/cells.jl:196-222, incremental: after emdee_set_positions, recompute every atom's cell; if no
 * atom changed cell, the sorted order and cell table are kept (and the pair list too while no atom has moved more than skin/2
 * since the binning); otherwise the movers are re-linked by a full (cell, id) re-sort.  *movers = atoms whose cell changed
 * (-1 when the full path was taken without counting: first use, new cutoff or skin, slab decomposition). */
int emdee_update_cells(emdee_system *sys, int64_t *movers);
int emdee_get_cells_per_dimension(emdee_system *sys, int32_t *M);
int emdee_get_cell_index(emdee_system *sys, int32_t *index_N);           /* 1-based, id order (Cells.index) */
int emdee_get_cell_population(emdee_system *sys, int32_t *pop_M3);       /* Cells.population */
int emdee_get_cell_order(emdee_system *sys, int32_t *perm_N, int32_t *cell_start_M3p1); /* ids sorted by (cell,id) */

/* compute_nonbonded!(forces, energies, virials, positions, L, tiles, model, atoms, Val(bitmask)):
 * src/nonbonded.jl:109-120 (launcher) and :44-107 (compute_tile!).  Results stay on the device;
 * fetch them with emdee_get_*.  Asynchronous like the reference launch; getters synchronise.
 * mode EMDEE_CUTOFF needs emdee_bin first; EMDEE_ALLPAIRS_REFERENCE uses the tile list set by
 * emdee_set_tiles (default: nonbonded_computation_tiles(N), src/nonbonded.jl:18-26). */
int emdee_compute_nonbonded(emdee_system *sys, int mode, int bitmask);
/* The same evaluation with the reference's output arguments: compute_nonbonded!(forces, energies, virials, ...) fills three
 * HOST arrays in id order (src/nonbonded.jl:122-155; forces N rows of 3, row-major).  Equivalent to emdee_compute_nonbonded
 * followed by emdee_get_forces / _energies / _virials for the outputs `bitmask` selects (the other pointers may be NULL), but on
 * one GPU the box is evaluated in chunks of z planes and the rows of a finished chunk travel to the host while the next chunk
 * computes (falls back to the plain sequence when atom ids do not run with z, or EMDEE_PIPE=0).  Synchronous. */
int emdee_compute_nonbonded_into(emdee_system *sys, int mode, int bitmask, double *forces_Nx3, double *energies_N, double *virials_N);
int emdee_set_tiles(emdee_system *sys, const int32_t *tiles_2xT, int64_t ntiles);
int emdee_get_positions(emdee_system *sys, double *pos_3xN);
int emdee_get_velocities(emdee_system *sys, double *vel_3xN);
int emdee_get_forces(emdee_system *sys, double *forces_3xN);
int emdee_get_energies(emdee_system *sys, double *energies_N);
int emdee_get_virials(emdee_system *sys, double *virials_N);
/* sum of per-atom energies / virials and the number of pairs in the set (this rank's atoms). */
int emdee_get_totals(emdee_system *sys, double *E, double *W, int64_t *npairs);

/* Cyclic windows of the id-ordered host arrays (additive; a slab rank then moves only the rows of the atoms it owns over PCIe
 * instead of all N): rows *id_first, *id_first + 1, ... (mod N), *count of them, are the smallest such window covering every
 * atom this rank owns, to 1/4096 of N (all N before the first emdee_bin of a decomposed system; cyclic because the periodic box
 * makes the first and last slab own a few atoms from the other end of the id order).  The *_range calls take the FULL 3xN / N
 * host array and touch only the window's rows: emdee_set_positions_range uploads them and fails if the window misses an atom
 * the rank owns (ghosts outside it are refreshed from their owners by the next emdee_bin, which the call makes mandatory
 * anyway); the getters write the rows of the atoms this rank owns (zeros for other rows of the window), leave the rest of
 * the array alone, and on a slab rank fail likewise if an owned atom lies outside the window -- ownership changes at every
 * emdee_bin, so ask for the window again after it (one GPU owns every atom: any window of rows may be read there). */
int emdee_get_local_id_range(emdee_system *sys, int64_t *id_first, int64_t *count);
int emdee_set_positions_range(emdee_system *sys, int64_t id_first, int64_t count, const double *pos_3xN);
int emdee_get_forces_range(emdee_system *sys, int64_t id_first, int64_t count, double *forces_3xN);
int emdee_get_energies_range(emdee_system *sys, int64_t id_first, int64_t count, double *energies_N);
int emdee_get_virials_range(emdee_system *sys, int64_t id_first, int64_t count, double *virials_N);

/* Pair-set audit (the pair enumeration find_action_partners1! was heading to, src/cells.jl:224-297):
 * sorted (i<j) pairs for small N, or (count, sum hash, xor hash) with hash = splitmix64((i<<32)|j). */
int emdee_pair_set(emdee_system *sys, int32_t *ij_2xcap, int64_t cap, int64_t *n);
int emdee_pair_set_digest(emdee_system *sys, uint64_t out[3]);
/* Audit of the stepping path: number of pairs (i<j) the pair-list kernel evaluated inside the cutoff at the
 * current positions; equals emdee_pair_set_digest's count on the same positions (the list pre-culls are
 * conservative, the final decision is the oracle's).  Needs a valid list, i.e. a preceding emdee_vv_step.
 * Returns -1 in *npairs when the system steps without a list (no skin, EMDEE_LIST=0, > 16 LJ classes). */
int emdee_list_pair_count(emdee_system *sys, int64_t *npairs);
/* Host-only: the conservative FP16 threshold the list kernels pre-cull with, for coordinates within +-half_extent[k]
 * of a brick centre: every pair with r <= rcut satisfies r2_fp16 <= *threshold (tests/test_host_logic.py checks it). */
int emdee_fp16_threshold(const double half_extent[3], double rcut, float *threshold);

/* Velocity-Verlet (absent from the reference): nsteps of
 * v += dt/2m f ; r += dt v ; f = F(r) ; v += dt/2m f, re-binning every `rebin_every` steps
 * (0: never; < 0 with a skin: adaptively, on the first step on which an atom has moved more than skin/2 since the
 * last binning -- one 4-byte read-back per step, and a max over ranks in a slab decomposition).  Needs model, atoms, masses, velocities, one emdee_bin and one emdee_compute_nonbonded. */
int emdee_vv_step(emdee_system *sys, double dt, int64_t nsteps, int rebin_every);
/* Host-only counters of the fused stepping loop since the system was created: re-binnings, and stepping launches that walked
 * the pair list in full, pruned it into the inner list (two-level list: the entries inside rc + skin2; opt-in with EMDEE_SKIN2=<skin2>, e.g. 0.12;
 * off by default), or replayed the inner list.  The inner list is used by one-GPU runs with adaptive re-binning
 * (rebin_every < 0): the per-step read-back tells the host when an atom may have moved skin2 / 2 since the last prune step. */
int emdee_get_step_counters(emdee_system *sys, int64_t out[4]);
int emdee_kinetic_energy(emdee_system *sys, double *K);
/* v *= factor for every atom this rank owns: the device half of a velocity-rescaling thermostat (the reference has
 * none); the host computes the factor from emdee_kinetic_energy between emdee_vv_step calls. */
int emdee_scale_velocities(emdee_system *sys, double factor);
int emdee_synchronize(emdee_system *sys);
/* number of kernels this library has launched on the context since creation */
int emdee_launch_count(emdee_ctx *ctx, int64_t *n);
/* elapsed GPU time of the last compute/vv call's dominant kernel is measured by the caller with
 * these markers, recorded on the library's compute stream */
int emdee_timer_start(emdee_ctx *ctx);
int emdee_timer_stop(emdee_ctx *ctx, double *ms);

/* Per-kernel timing of the dominant kernel (the cell-list force kernel) with CUDA events recorded on the
 * library's stream around every launch between begin and end; end synchronises and returns the summed
 * duration and the number of launches (roofline.achieved in bench.py). */
int emdee_profile_begin(emdee_system *sys);
int emdee_profile_end(emdee_system *sys, double *force_kernel_ms, int64_t *force_kernel_launches);
/* The same split by kernel (valid after emdee_profile_end): kind 0 = k_force_cells (window scan, single-point
 * evaluations), 1 = k_list_build (pair-list build on a re-binning step), 2 = k_force_list(_p) (the stepping kernel). */
int emdee_profile_kind(emdee_system *sys, int kind, double *ms, int64_t *launches);

/* Host-only: how the stepping path is configured after the last emdee_bin (what bench.py names as the dominant
 * kernel): out = {brick cells x, y, z, staged-atom capacity of a brick, list-capable (1: k_list_build + list walk,
 * 0: window scan on every step), persistent (bit 0: k_force_list_p instead of k_force_list, one block per brick; bit 1: bulk-copy
 * staging; bit 2: compacted staging -- only the atoms within rc + skin of the home box; bit 3: shallow per-lane stacks; bit 4: split lists, two lanes per home atom),
 * velocity-Verlet fused into the stepping kernel (1/0), list chunks of 8 entries per atom}. */
int emdee_get_step_config(emdee_system *sys, int32_t out[8]);

/* Slab decomposition info (valid after emdee_bin): atoms owned by this rank and their ids */
int emdee_get_local_count(emdee_system *sys, int64_t *nlocal, int64_t *nghost);
int emdee_get_local_ids(emdee_system *sys, int32_t *ids_nlocal);

/* One-shot mirror of compute_nonbonded! (src/nonbonded.jl:109-120) on host arrays: upload, compute,
 * download.  tiles may be NULL (default list).  Unselected outputs are not written. */
int emdee_compute_nonbonded_host(int64_t N, const double *pos_3xN, double L, double cutoff, double switch_dist,
                                 const double *atoms_2xN, const int32_t *tiles_2xT, int64_t ntiles, int mode,
                                 int ndiv, int bitmask, double *forces_3xN, double *energies_N, double *virials_N);

#ifdef __cplusplus
}
#endif
#endif
