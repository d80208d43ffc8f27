// emdee_b200.cu -- host side of libemdee_b200.so: contexts, systems, launch logic and the C ABI
// declared in include/emdee_b200.h.  sm_100a only; there is no CPU path in this library.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>

#include "common.cuh"
#include "binning.cuh"
#include "force_cells.cuh"
#include "force_list.cuh"
#include "force_list_p.cuh"
#include "list_build.cuh"
#include "force_tiles.cuh"
#include "integrate.cuh"
#include "slab.cuh"
#include "pairs14.cuh"

#define FC_REGS_ESTIMATE 128   // registers per thread of k_force_cells assumed by the residency model
#define FL_REGS_ESTIMATE 128   // same for k_force_list (__launch_bounds__(256, 2))

// NCCL is bound at run time (dlopen) instead of link time: a host process that also imports PyTorch must end
// up with ONE libnccl.so.2 (torch bundles 2.28, the system has 2.27); dlopen by soname returns whichever copy
// is already loaded, and single-GPU users need no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load()
{
    if (g_nccl.handle) return EMDEE_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) EMDEE_FAIL(EMDEE_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
#define NCCL_SYM(field, name)                                                                     \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                                   \
    if (!g_nccl.field) EMDEE_FAIL(EMDEE_ERR_NCCL, "libnccl.so.2 lacks %s", name);
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId") NCCL_SYM(CommInitRank, "ncclCommInitRank") NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(Send, "ncclSend") NCCL_SYM(Recv, "ncclRecv") NCCL_SYM(GroupStart, "ncclGroupStart") NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NCCL_SYM(GetErrorString, "ncclGetErrorString") NCCL_SYM(AllReduce, "ncclAllReduce")
#undef NCCL_SYM
    g_nccl.handle = h;
    return EMDEE_OK;
}
#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclCommDestroy g_nccl.CommDestroy
#define ncclSend g_nccl.Send
#define ncclRecv g_nccl.Recv
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
#define ncclGetErrorString g_nccl.GetErrorString
#define NCCL_TRY(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess) {                                                                    \
            emdee_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, ncclGetErrorString(_r)); \
            return EMDEE_ERR_NCCL;                                                                  \
        }                                                                                           \
    } while (0)

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void emdee_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char *emdee_last_error(void) { return g_err; }
extern "C" int emdee_version(void) { return 100; }

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct emdee_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0, mem_bytes = 0;
    int64_t launches = 0;
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_compute = nullptr, ev_comm = nullptr;
    // emdee_compute_nonbonded_into: results leave for the host on their own stream, chunk by chunk, behind the compute stream
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> chunk_events;
};

struct emdee_system {
    emdee_ctx *ctx = nullptr;
    int64_t N = 0;          // global atom count
    double L = 0;
    // parameters
    bool has_model = false, has_atoms = false, has_pos = false, has_vel = false, has_mass = false, has_excl = false;
    double cutoff = 0, sw = 0, skin = 0;
    LJModel model{};
    // per-atom state, ping-pong for the re-ordering gather
    int64_t cap = 0;
    AtomArrays A[2];
    int cur = 0;
    double *f[3] = {nullptr, nullptr, nullptr}, *en = nullptr, *vir = nullptr;
    int32_t *gcell[2] = {nullptr, nullptr}, *lcell[2] = {nullptr, nullptr};
    int32_t *slot_of_id = nullptr;
    int64_t nlo = 0, nown = 0, nhi = 0;   // lower ghosts, owned, upper ghosts
    // cell grid
    bool binned = false, grid_ok = false;
    bool order_valid = false;                 // slot order and gcell[] are those of a binning with the current cutoff, skin and ndiv
    int ndiv = 0;
    GridDesc g{};
    int64_t ncell = 0, ncell_cap = 0;
    int32_t *count = nullptr, *cell_start = nullptr, *fill = nullptr, *order = nullptr, *src_of_new = nullptr;
    int32_t *block_sum = nullptr, *maxpop = nullptr;
    int64_t steps_since_bin = 0;
    int *brick_counter = nullptr;             // device: brick cursor of the persistent kernel
    // velocity-Verlet fused into the stepping kernel (producer warps advance the atoms of released bricks; EMDEE_FUSE_VV=0
    // switches back to k_vv): second buffer of scaled positions, per-launch mode
    bool fuse_vv = true;
    double *s_alt[3] = {nullptr, nullptr, nullptr};
    int vv_mode = 0, vv_check_skin = 0;
    bool vv_track = false;
    double vv_dt = 0;
    unsigned *maxd2 = nullptr;                // device: max |r - r_bin|^2 since the last binning (float bits), adaptive re-binning
    // slab decomposition (nranks > 1)
    // Scaled positions of a slab system live in ONE allocation (nine arrays: A[0].s, A[1].s, s_alt, behind 64 flag words), mapped
    // into both neighbours with one CUDA IPC handle: the fused integrator writes my boundary atoms' new positions straight into
    // the neighbours' ghost slots over NVLink and raises a flag there (force_list_p.cuh), so a step needs no NCCL call.
    double *spool = nullptr;
    size_t spool_stride = 0;                          // doubles per array
    void *peer_pool[2] = {nullptr, nullptr};          // lower / upper neighbour's pool (the same mapping when both are one peer)
    bool peer_tried = false, peer_ok = false;
    long long *peerinfo = nullptr;                    // device: [0..3] mine, [4..7] from the lower neighbour, [8..11] from the upper one
    unsigned long long epoch = 0;                     // id of the last fused slab launch (flags carry it)
    int ghost_state = 0;                              // ghost positions: 0 complete, 1 stale (exchange them), 2 being written by the neighbours' launch `epoch`
    bool p2p_launch = false;                          // the launch being issued uses peer-mapped halos
    unsigned long long p2p_wait = 0, p2p_publish = 0;
    int hi_layer0 = 0;                                // first brick z-layer whose halo reaches the upper ghost planes (unclamped)
    // values a slab re-binning reads back together with its slot marks (one host synchronisation instead of three)
    bool pre_valid = false;
    int pre_maxpop = 0, pre_cap = 0;
    // re-binning cadence of a slab run with rebin_every < 0: chosen at every re-binning from the largest displacement of the
    // interval that just ended (max over ranks, read back with the marks), so that no step needs a read-back
    int slab_interval = 4;
    double last_dmax_ratio = 0;                       // max displacement of the last interval / (skin/2)
    bool decomposed = false;
    int z0 = 0, nz = 0;                               // my global planes [z0, z0+nz)
    int64_t lo_send_a = 0, lo_send_n = 0, hi_send_a = 0, hi_send_n = 0;   // slot ranges my neighbours need as ghosts
    int brick_lo_end = 0, brick_hi_begin = 0;         // brick z layers [0,lo_end) and [hi_begin,nbz) read ghost planes
    int32_t *sendcount = nullptr, *recvcount = nullptr, *list_lo = nullptr, *list_hi = nullptr;
    double *migbuf[4] = {nullptr, nullptr, nullptr, nullptr};   // send lo, send hi, recv lo, recv hi
    int64_t migcap = 0;
    double *ghostbuf[4] = {nullptr, nullptr, nullptr, nullptr}; // packed ghost atoms: send lo, send hi, recv lo, recv hi
    int64_t ghostcap = 0;
    // force kernel configuration
    int fc_cap = 0, fc_ncs = 0, fc_block = 256, fc_nblocks = 0;
    bool fc_typed = false;
    int fc_shape[3] = {0, 0, 0};
    double fc_per_cell = 0;                   // atoms per cell and reach when the shape was chosen
    int fc_R = 0;
    int fl_block = 192;                       // block size of k_force_list
    bool fl_ilp8 = true;
    bool fl_persistent = false, want_persistent = true;   // k_force_list_p when two staging buffers fit in shared memory
    // staged-atom capacity, stack depth and staging mode of the list kernels: with compaction (dense cells: the 27 cells around a
    // one-cell brick do not fit twice) k_list_build keeps only the atoms within rc + skin of the home box, see CellArgs::compact
    int fl_cap = 0, fl_qcap = FL_QCAP;
    bool fl_compact = false;
    int want_compact = 1;      // EMDEE_COMPACT: 0 never, 1 when nothing else fits (dense cells), 2 whenever the persistent kernel runs on one GPU
    // two lanes of the persistent kernel per home atom (split lists): bricks with at most half as many warp tasks as consumer warps
    bool fl_split = false, want_split = true;
    bool fl_fuse = true;                                  // walk and drain share a basic block
    int reserve_sms = 0, nccl_sms = 4;                    // SMs left free for NCCL during the interior launch of a slab step (EMDEE_NCCL_SMS).
                                                          // With bricks claimed dynamically every SM is busy until the launch ends, so NCCL's kernel would
                                                          // only start then; measured at 2 GPUs: halo wait 85 -> 3 us, interior +14 us, step 1.136 -> 1.083 ms
    size_t fl_smem = 0;
    int fc_gmax = 0;                          // 32-atom groups per brick (pair-list addressing)
    uint4 *list8 = nullptr;                   // pair list: chunks of 8 x uint16 (staged index + 1) per home atom
    uint16_t *list_n = nullptr;               // entries per home atom
    int2 *recipe = nullptr;                   // staging recipe of every brick (k_list_build -> k_force_list_p)
    uint16_t *homeidx = nullptr;
    int *brickhdr = nullptr;
    int64_t recipe_cap = 0, hdr_cap = 0;
    int64_t list_slots = 0;                   // allocated groups
    int lcap8 = 24;                           // chunks per atom (192 entries; grown from the density, or EMDEE_LIST_CHUNKS)
    bool lcap8_forced = false;
    bool list_valid = false, use_list = true;
    // Newton's third law inside the brick (EMDEE_N3=1, fused velocity-Verlet steps only): the list then holds every pair of
    // two home atoms once; the other list kernels need the full list, so the flavour of the valid list is tracked
    bool want_n3 = false, list_n3 = false, build_n3 = false;
    // TMA staging of the persistent kernel (opt-in, EMDEE_TMA=1; the default is the recipe gathers, which measure faster:
    // profiles/README.md): per-brick segment table written by k_list_build, raw ring sized from the longest staged row
    bool want_tma = false, fl_tma = false;
    // two-level list of the fused stepping kernel (k_force_list_p<..., LM>): inner list logged by a prune step, replayed after it
    uint4 *inner8 = nullptr;
    int *inner_n = nullptr;
    int64_t inner_slots = 0;
    int inner_lcap8 = 0;
    double skin2 = 0.0;                       // inner skin (EMDEE_SKIN2, e.g. 0.12; 0: off -- the default: measured, it does not pay at the
                                              // bench's state point: a replay launch takes 1.10 ms, a prune launch 1.39 ms, a plain walk 1.19 ms)
    int lm = 0;                               // mode of the next stepping launch: 0 plain walk, 1 prune, 2 replay
    uint64_t list_gen = 0, inner_gen = 0;     // the inner list is a subset of the list of generation inner_gen
    int steps_since_prune = 0;
    bool inner_at_current = false;            // the last stepping launch pruned or replayed at the positions the system still has
    int64_t counters[4] = {0, 0, 0, 0};       // re-binnings, stepping launches: plain, prune, replay (emdee_get_step_counters)
    int4 *seg = nullptr;
    int64_t seg_cap_total = 0;
    int segcap = 0, raw_rows = 0, rawlen = 0, fc_rowmax = 0, pre_rowmax = 0;
    size_t fc_smem_budget = 0;
    size_t fc_smem = 0;
    double2 *ljtab = nullptr;                 // pair table of the LJ parameter classes
    int ntypes = 0;                           // 0: too many classes, kernels gather per-atom parameters
    unsigned long long *digest = nullptr;     // device {count, sum, xor} + pair counter at [3]
    int *err = nullptr;                       // device error flag
    int *brick_max = nullptr;
    // 1-4 pairs (lj14scale): corrected after every CUTOFF force evaluation (pairs14.cuh)
    int32_t *pairs14 = nullptr;
    int64_t n14 = 0;
    double scale14 = 1.0;
    // tiles for ALLPAIRS_REFERENCE
    int32_t *tiles = nullptr;
    int64_t ntiles = 0;
    bool tiles_default = true;
    // results
    int last_mode = -1, last_bitmask = 0;
    bool forces_valid = false, kick_pending = false;
    // per-launch events of the force kernel (emdee_profile_begin/end)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    std::vector<int> prof_mode;               // launch mode of every event pair (0 scan, 1 list build, 2 list walk)
    size_t prof_used = 0;
    std::vector<cudaEvent_t> prof_mid;        // slab runs: end of the interior launch, start of the boundary launch
    std::vector<size_t> prof_mid_of;          // index of the event pair each mid pair belongs to
    double prof_ms[4] = {0, 0, 0, 0};         // per-kind totals of the last profile (emdee_profile_kind)
    int64_t prof_n[4] = {0, 0, 0, 0};
    // outputs of the audit passes (pair digest / pair set): the audit re-evaluates the pairs, its forces, energies and virials go
    // here so that the results of the last compute (and the bits of the next half-kick) are left alone
    double *aud[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int64_t aud_cap = 0;
    // scratch for host transfers
    double *tmp = nullptr;
    size_t tmp_bytes = 0;
    // run_cells on a range of brick layers (emdee_compute_nonbonded_into evaluates the box in chunks of z layers); 0 layers: all
    int range_first = 0, range_count = 0;
};

#define LAUNCH_1D(ctx, kernel, n, ...)                                                        \
    do {                                                                                      \
        if ((n) > 0) {                                                                        \
            kernel<<<(unsigned)ceil_div64((n), 256), 256, 0, (ctx)->stream>>>(__VA_ARGS__);   \
            (ctx)->launches++;                                                                \
        }                                                                                     \
    } while (0)

static int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) EMDEE_FAIL(EMDEE_ERR_CUDA, "kernel launch %s failed: %s", what, cudaGetErrorString(e));
    return EMDEE_OK;
}

template <typename T>
static int dev_alloc(T **p, size_t n)
{
    *p = nullptr;
    if (n == 0) n = 1;
    CUDA_TRY(cudaMalloc((void **)p, n * sizeof(T)));
    return EMDEE_OK;
}
template <typename T>
static void dev_free(T *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

static int ensure_audit_scratch(emdee_system *s)
{
    if (s->aud_cap >= s->cap) return EMDEE_OK;
    for (int k = 0; k < 5; k++) {
        if (s->aud[k]) cudaFree(s->aud[k]);
        s->aud[k] = nullptr;
        CUDA_TRY(cudaMalloc((void **)&s->aud[k], sizeof(double) * (size_t)s->cap));
    }
    s->aud_cap = s->cap;
    return EMDEE_OK;
}
static int ensure_tmp(emdee_system *s, size_t bytes)
{
    if (bytes <= s->tmp_bytes) return EMDEE_OK;
    dev_free(s->tmp);
    CUDA_TRY(cudaMalloc((void **)&s->tmp, bytes));
    s->tmp_bytes = bytes;
    return EMDEE_OK;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int emdee_create(emdee_ctx **out, int device)
{
    if (!out) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_create: null output pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        EMDEE_FAIL(EMDEE_ERR_CUDA, "emdee_create: no CUDA device (%s); this library has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_create: device %d out of range [0,%d)", device, ndev);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        EMDEE_FAIL(EMDEE_ERR_CUDA, "emdee_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                   device, prop.major, prop.minor);
    CUDA_TRY(cudaSetDevice(device));
    emdee_ctx *c = new emdee_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    c->mem_bytes = prop.totalGlobalMem;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    *out = c;
    return EMDEE_OK;
}

extern "C" int emdee_destroy(emdee_ctx *c)
{
    if (!c) return EMDEE_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
    if (c->comm) ncclCommDestroy(c->comm);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_compute) cudaEventDestroy(c->ev_compute);
    if (c->ev_comm) cudaEventDestroy(c->ev_comm);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    for (cudaEvent_t e : c->chunk_events) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return EMDEE_OK;
}

extern "C" int emdee_comm_unique_id(char id[128])
{
    if (!id) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_comm_unique_id: null buffer");
    static_assert(NCCL_UNIQUE_ID_BYTES == 128, "ncclUniqueId is expected to be 128 bytes");
    EMDEE_TRY(nccl_load());
    ncclUniqueId u;
    NCCL_TRY(ncclGetUniqueId(&u));
    memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
    return EMDEE_OK;
}
extern "C" int emdee_comm_init(emdee_ctx *c, int rank, int nranks, const char id[128])
{
    if (!c) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_comm_init: null context");
    if (nranks < 1 || rank < 0 || rank >= nranks) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_comm_init: rank %d of %d", rank, nranks);
    if (c->comm) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_comm_init: communicator already initialised");
    if (nranks == 1) return EMDEE_OK;
    if (!id) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_comm_init: null id");
    CUDA_TRY(cudaSetDevice(c->device));
    EMDEE_TRY(nccl_load());
    ncclUniqueId u;
    memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
    NCCL_TRY(ncclCommInitRank(&c->comm, nranks, u, rank));
    {   // the halo exchange must get SMs ahead of the (persistent, SM-filling) force kernel that becomes ready together with it
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_compute, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming));
    c->rank = rank;
    c->nranks = nranks;
    return EMDEE_OK;
}

extern "C" int emdee_device_info(emdee_ctx *c, int *sm_count, int *cc_major, int *cc_minor, int64_t *mem_bytes)
{
    if (!c) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_device_info: null context");
    if (sm_count) *sm_count = c->sm_count;
    if (cc_major) *cc_major = c->cc_major;
    if (cc_minor) *cc_minor = c->cc_minor;
    if (mem_bytes) *mem_bytes = (int64_t)c->mem_bytes;
    return EMDEE_OK;
}

extern "C" int emdee_launch_count(emdee_ctx *c, int64_t *n)
{
    if (!c || !n) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_launch_count: null argument");
    *n = c->launches;
    return EMDEE_OK;
}
extern "C" int emdee_timer_start(emdee_ctx *c)
{
    if (!c) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_timer_start: null context");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    return EMDEE_OK;
}
extern "C" int emdee_timer_stop(emdee_ctx *c, double *ms)
{
    if (!c || !ms) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_timer_stop: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    float t = 0;
    CUDA_TRY(cudaEventElapsedTime(&t, c->ev0, c->ev1));
    *ms = (double)t;
    return EMDEE_OK;
}

extern "C" int emdee_measure_fp64_peak(emdee_ctx *c, double *flops_per_s, double *ms_out)
{
    if (!c || !flops_per_s) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_measure_fp64_peak: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    double *d = nullptr;
    EMDEE_TRY(dev_alloc(&d, 1));
    const int iters = 1 << 15, blocks = c->sm_count * 8, threads = 256;
    double best = 1e30;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
        k_dfma_peak<<<blocks, threads, 0, c->stream>>>(iters, 1.0 + rep, d);
        c->launches++;
        CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
        CUDA_TRY(cudaEventSynchronize(c->ev1));
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, c->ev0, c->ev1));
        if (rep > 0) best = std::min(best, (double)t);
    }
    dev_free(d);
    EMDEE_TRY(check_launch("k_dfma_peak"));
    const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
    *flops_per_s = flops / (best * 1e-3);
    if (ms_out) *ms_out = best;
    return EMDEE_OK;
}

// ------------------------------------------------------------------------------------------------
// system
// ------------------------------------------------------------------------------------------------
static int alloc_atoms(AtomArrays &A, int64_t cap)
{
    for (int c = 0; c < 3; c++) {
        EMDEE_TRY(dev_alloc(&A.r[c], cap));
        EMDEE_TRY(dev_alloc(&A.s[c], cap + 2));       // bulk copies round a row up to an even number of slots
        EMDEE_TRY(dev_alloc(&A.v[c], cap));
        EMDEE_TRY(dev_alloc(&A.rb[c], cap));
    }
    EMDEE_TRY(dev_alloc(&A.hs, cap));
    EMDEE_TRY(dev_alloc(&A.ts, cap));
    EMDEE_TRY(dev_alloc(&A.mass, cap));
    EMDEE_TRY(dev_alloc(&A.id, cap));
    EMDEE_TRY(dev_alloc(&A.type, cap));
    EMDEE_TRY(dev_alloc(&A.xbase, cap));
    EMDEE_TRY(dev_alloc(&A.xmask, cap));
    return EMDEE_OK;
}
static void free_atoms(AtomArrays &A)
{
    for (int c = 0; c < 3; c++) { dev_free(A.r[c]); dev_free(A.s[c]); dev_free(A.v[c]); dev_free(A.rb[c]); }
    dev_free(A.hs); dev_free(A.ts); dev_free(A.mass); dev_free(A.id); dev_free(A.type); dev_free(A.xbase); dev_free(A.xmask);
}

extern "C" int emdee_system_create(emdee_ctx *c, int64_t N, double L, emdee_system **out)
{
    if (!c || !out) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_system_create: null argument");
    *out = nullptr;
    if (N <= 0 || N > 0x7fffffff) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_system_create: N=%lld must be in [1, 2^31)", (long long)N);
    if (!(L > 0) || !std::isfinite(L)) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_system_create: L=%g must be positive and finite", L);
    CUDA_TRY(cudaSetDevice(c->device));
    emdee_system *s = new emdee_system();
    s->ctx = c;
    s->N = N;
    s->L = L;
    if (const char *e = getenv("EMDEE_LIST")) s->use_list = atoi(e) != 0;
    if (const char *e = getenv("EMDEE_LIST_CHUNKS")) { s->lcap8 = std::max(4, atoi(e)); s->lcap8_forced = true; }
    if (const char *e = getenv("EMDEE_ILP8")) s->fl_ilp8 = atoi(e) != 0;
    if (const char *e = getenv("EMDEE_PERSIST")) s->want_persistent = atoi(e) != 0;
    if (const char *e = getenv("EMDEE_FUSE")) s->fl_fuse = atoi(e) != 0;
    if (const char *e = getenv("EMDEE_FUSE_VV")) s->fuse_vv = atoi(e) != 0;
    if (const char *e = getenv("EMDEE_N3")) s->want_n3 = atoi(e) != 0;
    if (const char *e = getenv("EMDEE_TMA")) s->want_tma = atoi(e) != 0;
    if (const char *e = getenv("EMDEE_SKIN2")) s->skin2 = std::max(0.0, atof(e));
    if (const char *e = getenv("EMDEE_COMPACT")) s->want_compact = atoi(e);
    if (const char *e = getenv("EMDEE_SPLIT")) s->want_split = atoi(e) != 0;
    if (const char *e = getenv("EMDEE_NCCL_SMS")) s->nccl_sms = std::max(0, atoi(e));
    s->cap = N + (c->nranks > 1 ? N / 4 + 1024 : 0);   // head-room for ghost copies in a slab decomposition
    s->nown = N;
    int st = EMDEE_OK;
    auto A = [&](int rc) { if (st == EMDEE_OK) st = rc; };
    A(alloc_atoms(s->A[0], s->cap));
    A(alloc_atoms(s->A[1], s->cap));
    for (int k = 0; k < 3; k++) A(dev_alloc(&s->f[k], s->cap));
    A(dev_alloc(&s->en, s->cap));
    A(dev_alloc(&s->vir, s->cap));
    for (int k = 0; k < 2; k++) { A(dev_alloc(&s->gcell[k], s->cap)); A(dev_alloc(&s->lcell[k], s->cap)); }
    A(dev_alloc(&s->slot_of_id, N));
    A(dev_alloc(&s->order, s->cap));
    A(dev_alloc(&s->src_of_new, s->cap));
    A(dev_alloc(&s->maxd2, 4));              // [0] since the binning, [1] its max over ranks, [2] largest step since the last prune step
    A(dev_alloc(&s->brick_counter, 4));      // [0] brick cursor of the persistent kernel, [1], [2] boundary bricks advanced (peer halos)
    A(dev_alloc(&s->digest, 16));        // [0..3] audit digest + pair counter, [8..15] role timers of FLP_TIMING builds
    A(dev_alloc(&s->err, 1));
    A(dev_alloc(&s->maxpop, 1));
    A(dev_alloc(&s->brick_max, 2));         // staged atoms of the fullest brick, atoms of the longest staged row
    if (c->nranks > 1 && st == EMDEE_OK) {
        for (int b = 0; b < 2; b++)
            for (int k = 0; k < 3; k++) dev_free(s->A[b].s[k]);
        s->spool_stride = ((size_t)s->cap + 2 + 31) & ~(size_t)31;
        A(dev_alloc(&s->spool, 64 + 9 * s->spool_stride));
        if (st == EMDEE_OK) {
            for (int k = 0; k < 3; k++) {
                s->A[0].s[k] = s->spool + 64 + (size_t)k * s->spool_stride;
                s->A[1].s[k] = s->spool + 64 + (size_t)(3 + k) * s->spool_stride;
                s->s_alt[k] = s->spool + 64 + (size_t)(6 + k) * s->spool_stride;
            }
            if (cudaMemsetAsync(s->spool, 0, 64 * sizeof(double), c->stream) != cudaSuccess) st = EMDEE_ERR_CUDA;
        }
        A(dev_alloc(&s->peerinfo, 12));
    }
    if (c->nranks > 1) {
        A(dev_alloc(&s->sendcount, 2));
        A(dev_alloc(&s->recvcount, 2));
        A(dev_alloc(&s->list_lo, s->cap));
        A(dev_alloc(&s->list_hi, s->cap));
    }
    if (st != EMDEE_OK) { emdee_system_destroy(s); return st; }
    // slot == id until the first binning; unit masses; zero velocities; no exclusions
    LAUNCH_1D(c, k_iota, N, N, s->A[0].id);
    LAUNCH_1D(c, k_iota, N, N, s->slot_of_id);
    LAUNCH_1D(c, k_fill<double>, N, N, s->A[0].mass, 1.0);
    for (int k = 0; k < 3; k++) CUDA_TRY(cudaMemsetAsync(s->A[0].v[k], 0, sizeof(double) * N, c->stream));
    CUDA_TRY(cudaMemsetAsync(s->A[0].type, 0, sizeof(int32_t) * N, c->stream));
    CUDA_TRY(cudaMemsetAsync(s->A[0].xbase, 0, sizeof(int32_t) * N, c->stream));
    CUDA_TRY(cudaMemsetAsync(s->A[0].xmask, 0, sizeof(uint64_t) * N, c->stream));
    CUDA_TRY(cudaMemsetAsync(s->err, 0, sizeof(int), c->stream));
    CUDA_TRY(cudaMemsetAsync(s->digest, 0, 16 * sizeof(unsigned long long), c->stream));
    EMDEE_TRY(check_launch("system init"));
    *out = s;
    return EMDEE_OK;
}

extern "C" int emdee_system_destroy(emdee_system *s)
{
    if (!s) return EMDEE_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->spool) {        // the scaled positions are slices of the pool
        for (int k = 0; k < 3; k++) { s->A[0].s[k] = nullptr; s->A[1].s[k] = nullptr; s->s_alt[k] = nullptr; }
        if (s->peer_pool[0]) cudaIpcCloseMemHandle(s->peer_pool[0]);
        if (s->peer_pool[1] && s->peer_pool[1] != s->peer_pool[0]) cudaIpcCloseMemHandle(s->peer_pool[1]);
        dev_free(s->spool);
        dev_free(s->peerinfo);
    }
    free_atoms(s->A[0]);
    free_atoms(s->A[1]);
    for (int k = 0; k < 3; k++) dev_free(s->f[k]);
    dev_free(s->en); dev_free(s->vir);
    for (int k = 0; k < 2; k++) { dev_free(s->gcell[k]); dev_free(s->lcell[k]); }
    dev_free(s->slot_of_id); dev_free(s->order); dev_free(s->src_of_new);
    dev_free(s->count); dev_free(s->cell_start); dev_free(s->fill); dev_free(s->block_sum);
    dev_free(s->ljtab); dev_free(s->digest); dev_free(s->maxd2); dev_free(s->brick_counter);
    dev_free(s->inner8); dev_free(s->inner_n);
    for (int k = 0; k < 3; k++) dev_free(s->s_alt[k]);
    dev_free(s->err); dev_free(s->maxpop); dev_free(s->brick_max); dev_free(s->tiles); dev_free(s->tmp);
    for (int k = 0; k < 5; k++) dev_free(s->aud[k]);
    for (cudaEvent_t e : s->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : s->prof_mid) cudaEventDestroy(e);
    dev_free(s->list8); dev_free(s->list_n); dev_free(s->recipe); dev_free(s->homeidx); dev_free(s->brickhdr);
    dev_free(s->sendcount); dev_free(s->recvcount); dev_free(s->list_lo); dev_free(s->list_hi);
    for (int k = 0; k < 4; k++) dev_free(s->migbuf[k]);
    for (int k = 0; k < 4; k++) dev_free(s->ghostbuf[k]);
    dev_free(s->pairs14);
    dev_free(s->seg);
    delete s;
    return EMDEE_OK;
}

#define SYS_ENTER(s, name)                                                            \
    if (!(s)) EMDEE_FAIL(EMDEE_ERR_INVALID, name ": null system");                    \
    emdee_ctx *c = (s)->ctx;                                                          \
    CUDA_TRY(cudaSetDevice(c->device));                                               \
    AtomArrays &A = (s)->A[(s)->cur];                                                 \
    (void)A;                                                                          \
    const int64_t ntot = (s)->nlo + (s)->nown + (s)->nhi;                             \
    (void)ntot;

extern "C" int emdee_set_model(emdee_system *s, double cutoff, double sw)
{
    SYS_ENTER(s, "emdee_set_model");
    if (!(sw > 0) || !(cutoff > sw) || !std::isfinite(cutoff))
        EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_set_model: need 0 < switch (%g) < cutoff (%g)", sw, cutoff);
    if (cutoff > 0.5 * s->L)
        EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_set_model: cutoff %g exceeds L/2 = %g (minimum image invalid)", cutoff, 0.5 * s->L);
    // LennardJonesModel(cutoff, switch) = new(cutoff^2, switch^2, 1/(cutoff^2 - switch^2)), src/lennard_jones.jl:10
    if (cutoff != s->cutoff) {     // the cell grid and the pair list were sized for the old cutoff (+ skin)
        s->binned = false;
        s->list_valid = false;
        s->order_valid = false;
    }
    s->cutoff = cutoff;
    s->sw = sw;
    s->model.rc2 = cutoff * cutoff;
    s->model.rs2 = sw * sw;
    s->model.id2 = 1.0 / (cutoff * cutoff - sw * sw);
    s->has_model = true;
    s->forces_valid = false;
    return EMDEE_OK;
}

extern "C" int emdee_set_skin(emdee_system *s, double skin)
{
    SYS_ENTER(s, "emdee_set_skin");
    if (!(skin >= 0) || !std::isfinite(skin)) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_set_skin: skin=%g must be >= 0", skin);
    s->skin = skin;
    s->binned = false;
    s->order_valid = false;
    return EMDEE_OK;
}

static int upload(emdee_system *s, const void *host, size_t bytes)
{
    EMDEE_TRY(ensure_tmp(s, bytes));
    CUDA_TRY(cudaMemcpyAsync(s->tmp, host, bytes, cudaMemcpyHostToDevice, s->ctx->stream));
    return EMDEE_OK;
}

extern "C" int emdee_set_lj_atoms(emdee_system *s, const double *atoms)
{
    SYS_ENTER(s, "emdee_set_lj_atoms");
    if (!atoms) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_set_lj_atoms: null array");
    EMDEE_TRY(upload(s, atoms, sizeof(double) * 2 * s->N));
    LAUNCH_1D(c, k_set1<double>, ntot, 0, ntot, A.id, s->tmp, 2, 0, A.hs);
    LAUNCH_1D(c, k_set1<double>, ntot, 0, ntot, A.id, s->tmp, 2, 1, A.ts);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    // LJ parameter classes: distinct (half_sigma, twice_sqrt_eps) rows.  Up to FC_MAX_TYPES classes the
    // kernels use a pair table {sigma_ij^2, 4 eps_ij} in shared memory (src/lennard_jones.jl:29,33);
    // beyond that they gather the per-atom parameters.
    std::vector<double> cls;
    std::vector<int32_t> type((size_t)s->N, 0);
    bool many = false;
    for (int64_t i = 0; i < s->N && !many; i++) {
        const double h = atoms[2 * i], t = atoms[2 * i + 1];
        size_t k = 0;
        for (; k < cls.size() / 2; k++)
            if (cls[2 * k] == h && cls[2 * k + 1] == t) break;
        if (k == cls.size() / 2) {
            if (k == FC_MAX_TYPES) { many = true; break; }
            cls.push_back(h); cls.push_back(t);
        }
        type[i] = (int32_t)k;
    }
    s->ntypes = many ? 0 : (int)(cls.size() / 2);
    if (!many) {
        const int T = s->ntypes;
        std::vector<double2> tab((size_t)T * T);
        for (int p = 0; p < T; p++)
            for (int q = 0; q < T; q++) {
                const double sig = cls[2 * p] + cls[2 * q];     // :29, squared here so that the kernels skip one multiply
                tab[(size_t)p * T + q] = make_double2(sig * sig, cls[2 * p + 1] * cls[2 * q + 1]);
            }
        dev_free(s->ljtab);
        EMDEE_TRY(dev_alloc(&s->ljtab, tab.size()));
        CUDA_TRY(cudaMemcpy(s->ljtab, tab.data(), sizeof(double2) * tab.size(), cudaMemcpyHostToDevice));
        EMDEE_TRY(upload(s, type.data(), sizeof(int32_t) * s->N));
        LAUNCH_1D(c, k_set1<int32_t>, ntot, 0, ntot, A.id, (const int32_t *)s->tmp, 1, 0, A.type);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    s->has_atoms = true;
    s->forces_valid = false;
    return check_launch("set_lj_atoms");
}

extern "C" int emdee_set_positions(emdee_system *s, const double *pos)
{
    SYS_ENTER(s, "emdee_set_positions");
    if (!pos) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_set_positions: null array");
    EMDEE_TRY(upload(s, pos, sizeof(double) * 3 * s->N));
    LAUNCH_1D(c, k_set3, ntot, 0, ntot, A.id, s->tmp, A.r[0], A.r[1], A.r[2]);
    LAUNCH_1D(c, k_scale_positions, ntot, ntot, A.r[0], A.r[1], A.r[2], s->L, A.s[0], A.s[1], A.s[2]);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    s->has_pos = true;
    s->binned = false;        // cells must be rebuilt (update_cells!, src/cells.jl:196)
    s->inner_at_current = false;
    s->forces_valid = false;
    s->kick_pending = false;
    return check_launch("set_positions");
}

// ---- cyclic windows of the id-ordered host arrays: a slab rank moves only the rows of the atoms it owns ----------------------
#define ID_BUCKETS 4096
extern "C" int emdee_get_local_id_range(emdee_system *s, int64_t *id_first, int64_t *count)
{
    SYS_ENTER(s, "emdee_get_local_id_range");
    if (!id_first || !count) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_get_local_id_range: null output");
    if (s->nown == s->N) { *id_first = 0; *count = s->N; return EMDEE_OK; }   // not decomposed (yet): every atom is here
    // occupancy of ID_BUCKETS equal id ranges; the window is the complement of the longest cyclic run of empty ones
    EMDEE_TRY(ensure_tmp(s, ID_BUCKETS));
    CUDA_TRY(cudaMemsetAsync(s->tmp, 0, ID_BUCKETS, c->stream));
    LAUNCH_1D(c, k_id_buckets, s->nown, s->nlo, s->nown, A.id, s->N, ID_BUCKETS, reinterpret_cast<unsigned char *>(s->tmp));
    unsigned char occ[ID_BUCKETS];
    CUDA_TRY(cudaMemcpyAsync(occ, s->tmp, ID_BUCKETS, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    EMDEE_TRY(check_launch("k_id_buckets"));
    int best_len = 0, best_start = 0, run = 0;
    for (int k = 0; k < 2 * ID_BUCKETS; k++) {          // twice around: runs that wrap
        if (!occ[k % ID_BUCKETS]) {
            if (++run > best_len && run <= ID_BUCKETS) { best_len = run; best_start = k - run + 1; }
        } else
            run = 0;
    }
    if (best_len >= ID_BUCKETS) { *id_first = 0; *count = 0; return EMDEE_OK; }     // owns nothing
    const int b0 = (best_start + best_len) % ID_BUCKETS, nb = ID_BUCKETS - best_len;   // occupied arc: buckets b0 .. b0 + nb - 1 (cyclic)
    // bucket k holds ids [ceil(k N / B), ceil((k+1) N / B))
    auto lo_of = [&](int64_t k) { return (k * s->N + ID_BUCKETS - 1) / ID_BUCKETS; };
    const int64_t first = lo_of(b0);
    int64_t end = lo_of((int64_t)b0 + nb);             // may exceed N: the window wraps
    if (b0 + nb > ID_BUCKETS) end = s->N + lo_of((int64_t)b0 + nb - ID_BUCKETS);
    *id_first = first;
    *count = std::min<int64_t>(end - first, s->N);
    return EMDEE_OK;
}

// rows [id_first, id_first + count) (mod N) of a full id-ordered host array <-> `count` contiguous rows of the device scratch
static int window_copy(emdee_system *s, void *host_full, int64_t id_first, int64_t count, size_t row_bytes, bool to_device)
{
    emdee_ctx *c = s->ctx;
    char *h = reinterpret_cast<char *>(host_full), *d = reinterpret_cast<char *>(s->tmp);
    const int64_t n1 = std::min(count, s->N - id_first), n2 = count - n1;
    if (to_device) {
        CUDA_TRY(cudaMemcpyAsync(d, h + id_first * row_bytes, n1 * row_bytes, cudaMemcpyHostToDevice, c->stream));
        if (n2 > 0) CUDA_TRY(cudaMemcpyAsync(d + n1 * row_bytes, h, n2 * row_bytes, cudaMemcpyHostToDevice, c->stream));
    } else {
        CUDA_TRY(cudaMemcpyAsync(h + id_first * row_bytes, d, n1 * row_bytes, cudaMemcpyDeviceToHost, c->stream));
        if (n2 > 0) CUDA_TRY(cudaMemcpyAsync(h, d + n1 * row_bytes, n2 * row_bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    return EMDEE_OK;
}
static int window_check(emdee_system *s, const void *p, int64_t id_first, int64_t count, const char *what)
{
    if (!p || id_first < 0 || id_first >= s->N || count < 0 || count > s->N)
        EMDEE_FAIL(EMDEE_ERR_INVALID, "%s: window of %lld rows from row %lld of %lld", what, (long long)count, (long long)id_first, (long long)s->N);
    return EMDEE_OK;
}

// an owned atom outside the caller's window (device flag 7): the window is stale -- ownership changes at every re-binning
static int window_flag(emdee_system *s, const char *what)
{
    emdee_ctx *c = s->ctx;
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, s->err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (flag == 7) {
        CUDA_TRY(cudaMemsetAsync(s->err, 0, sizeof(int), c->stream));
        EMDEE_FAIL(EMDEE_ERR_INVALID, "%s: the window does not cover every atom this rank owns (ask emdee_get_local_id_range again after a re-binning)", what);
    }
    return EMDEE_OK;
}

extern "C" int emdee_set_positions_range(emdee_system *s, int64_t id_first, int64_t count, const double *pos)
{
    SYS_ENTER(s, "emdee_set_positions_range");
    EMDEE_TRY(window_check(s, pos, id_first, count, "emdee_set_positions_range"));
    EMDEE_TRY(ensure_tmp(s, sizeof(double) * 3 * std::max<int64_t>(count, 1)));
    EMDEE_TRY(window_copy(s, const_cast<double *>(pos), id_first, count, 3 * sizeof(double), true));
    LAUNCH_1D(c, k_set3_range, ntot, 0, ntot, s->nlo, s->nlo + s->nown, A.id, id_first, count, s->N, s->tmp, A.r[0], A.r[1], A.r[2], s->err);
    LAUNCH_1D(c, k_scale_positions, ntot, ntot, A.r[0], A.r[1], A.r[2], s->L, A.s[0], A.s[1], A.s[2]);
    EMDEE_TRY(window_flag(s, "emdee_set_positions_range"));
    s->has_pos = true;
    s->binned = false;
    s->inner_at_current = false;
    s->forces_valid = false;
    s->kick_pending = false;
    return check_launch("set_positions_range");
}

static int get3_range(emdee_system *s, double *const src[3], int64_t id_first, int64_t count, double *out, const char *what)
{
    emdee_ctx *c = s->ctx;
    AtomArrays &A = s->A[s->cur];
    EMDEE_TRY(window_check(s, out, id_first, count, what));
    EMDEE_TRY(ensure_tmp(s, sizeof(double) * 3 * std::max<int64_t>(count, 1)));
    CUDA_TRY(cudaMemsetAsync(s->tmp, 0, sizeof(double) * 3 * count, c->stream));
    LAUNCH_1D(c, k_get3_range, s->nown, s->nlo, s->nown, A.id, id_first, count, s->N, src[0], src[1], src[2], s->tmp, c->nranks > 1 ? s->err : nullptr);
    if (c->nranks > 1) EMDEE_TRY(window_flag(s, what));
    EMDEE_TRY(window_copy(s, out, id_first, count, 3 * sizeof(double), false));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return check_launch(what);
}
static int get1_range(emdee_system *s, const double *src, int64_t id_first, int64_t count, double *out, const char *what)
{
    emdee_ctx *c = s->ctx;
    AtomArrays &A = s->A[s->cur];
    EMDEE_TRY(window_check(s, out, id_first, count, what));
    EMDEE_TRY(ensure_tmp(s, sizeof(double) * std::max<int64_t>(count, 1)));
    CUDA_TRY(cudaMemsetAsync(s->tmp, 0, sizeof(double) * count, c->stream));
    LAUNCH_1D(c, k_get1_range, s->nown, s->nlo, s->nown, A.id, id_first, count, s->N, src, s->tmp, c->nranks > 1 ? s->err : nullptr);
    if (c->nranks > 1) EMDEE_TRY(window_flag(s, what));
    EMDEE_TRY(window_copy(s, out, id_first, count, sizeof(double), false));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return check_launch(what);
}

extern "C" int emdee_set_velocities(emdee_system *s, const double *vel)
{
    SYS_ENTER(s, "emdee_set_velocities");
    if (!vel) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_set_velocities: null array");
    EMDEE_TRY(upload(s, vel, sizeof(double) * 3 * s->N));
    LAUNCH_1D(c, k_set3, ntot, 0, ntot, A.id, s->tmp, A.v[0], A.v[1], A.v[2]);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    s->has_vel = true;
    s->kick_pending = false;
    return check_launch("set_velocities");
}

extern "C" int emdee_set_masses(emdee_system *s, const double *mass)
{
    SYS_ENTER(s, "emdee_set_masses");
    if (!mass) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_set_masses: null array");
    for (int64_t i = 0; i < s->N; i++)
        if (!(mass[i] > 0)) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_set_masses: mass[%lld]=%g must be positive", (long long)i, mass[i]);
    EMDEE_TRY(upload(s, mass, sizeof(double) * s->N));
    LAUNCH_1D(c, k_set1<double>, ntot, 0, ntot, A.id, s->tmp, 1, 0, A.mass);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    s->has_mass = true;
    return check_launch("set_masses");
}

extern "C" int emdee_set_exclusions(emdee_system *s, const int32_t *base, const uint64_t *mask)
{
    SYS_ENTER(s, "emdee_set_exclusions");
    s->list_valid = false;         // exclusions are applied when the pair list is built (k_list_build<EXCL>)
    if (!base || !mask) {
        s->has_excl = false;
        s->forces_valid = false;
        return EMDEE_OK;
    }
    EMDEE_TRY(upload(s, base, sizeof(int32_t) * s->N));
    LAUNCH_1D(c, k_set1<int32_t>, ntot, 0, ntot, A.id, (const int32_t *)s->tmp, 1, 0, A.xbase);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    EMDEE_TRY(upload(s, mask, sizeof(uint64_t) * s->N));
    LAUNCH_1D(c, k_set1<uint64_t>, ntot, 0, ntot, A.id, (const uint64_t *)s->tmp, 1, 0, A.xmask);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    s->has_excl = true;
    s->forces_valid = false;
    return check_launch("set_exclusions");
}

extern "C" int emdee_set_pairs14(emdee_system *s, const int32_t *ij, int64_t n, double scale)
{
    SYS_ENTER(s, "emdee_set_pairs14");
    dev_free(s->pairs14);
    s->n14 = 0;
    s->scale14 = 1.0;
    s->forces_valid = false;
    if (!ij || n == 0) return EMDEE_OK;
    if (n < 0 || !std::isfinite(scale)) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_set_pairs14: n=%lld, scale=%g", (long long)n, scale);
    if (c->nranks > 1) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_set_pairs14: 1-4 scaling is not available in a slab decomposition");
    for (int64_t k = 0; k < n; k++)
        if (ij[2 * k] < 0 || ij[2 * k + 1] >= s->N || ij[2 * k] >= ij[2 * k + 1])
            EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_set_pairs14: pair %lld = (%d,%d) must satisfy 0 <= i < j < N", (long long)k, ij[2 * k], ij[2 * k + 1]);
    EMDEE_TRY(dev_alloc(&s->pairs14, (size_t)2 * n));
    CUDA_TRY(cudaMemcpy(s->pairs14, ij, sizeof(int32_t) * 2 * n, cudaMemcpyHostToDevice));
    s->n14 = n;
    s->scale14 = scale;
    return EMDEE_OK;
}

// (scale - 1) x the interaction of every 1-4 pair inside the cutoff, added to the selected outputs of the evaluation just made
static int apply_pairs14(emdee_system *s, int bitmask)
{
    if (s->n14 == 0 || s->scale14 == 1.0) return EMDEE_OK;
    emdee_ctx *c = s->ctx;
    AtomArrays &A = s->A[s->cur];
    Pairs14Args a;
    a.n = s->n14; a.ij = s->pairs14; a.slot_of_id = s->slot_of_id;
    a.sx = A.s[0]; a.sy = A.s[1]; a.sz = A.s[2]; a.hs = A.hs; a.ts = A.ts;
    a.L = s->L; a.cm1 = s->scale14 - 1.0; a.model = s->model; a.bitmask = bitmask;
    a.fx = s->f[0]; a.fy = s->f[1]; a.fz = s->f[2]; a.en = s->en; a.vir = s->vir;
    LAUNCH_1D(c, k_pairs14, a.n, a);
    return check_launch("k_pairs14");
}

extern "C" int emdee_set_tiles(emdee_system *s, const int32_t *tiles, int64_t ntiles)
{
    SYS_ENTER(s, "emdee_set_tiles");
    dev_free(s->tiles);
    s->ntiles = 0;
    s->tiles_default = true;
    if (!tiles) return EMDEE_OK;
    if (ntiles < 0) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_set_tiles: negative tile count");
    const int64_t nb = (s->N + 31) / 32;
    for (int64_t t = 0; t < ntiles; t++)
        if (tiles[2 * t] < 1 || tiles[2 * t] > nb || tiles[2 * t + 1] < 1 || tiles[2 * t + 1] > nb)
            EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_set_tiles: tile %lld = (%d,%d) outside [1,%lld]", (long long)t,
                       tiles[2 * t], tiles[2 * t + 1], (long long)nb);
    EMDEE_TRY(dev_alloc(&s->tiles, (size_t)2 * ntiles));
    CUDA_TRY(cudaMemcpy(s->tiles, tiles, sizeof(int32_t) * 2 * ntiles, cudaMemcpyHostToDevice));
    s->ntiles = ntiles;
    s->tiles_default = false;
    return EMDEE_OK;
}

// nonbonded_computation_tiles(N): n = cld(N,32); for i=0:n-1, j=1:n-i -> (j, j+i), src/nonbonded.jl:18-26
static int default_tiles(emdee_system *s)
{
    const int64_t n = (s->N + 31) / 32, nt = n * (n + 1) / 2;
    if (nt > (int64_t)1 << 28)
        EMDEE_FAIL(EMDEE_ERR_CAPACITY, "ALLPAIRS_REFERENCE: N=%lld needs %lld tiles; the all-pairs mode is for small N", (long long)s->N, (long long)nt);
    std::vector<int32_t> h((size_t)2 * nt);
    int64_t k = 0;
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 1; j <= n - i; j++) { h[2 * k] = (int32_t)j; h[2 * k + 1] = (int32_t)(j + i); k++; }
    dev_free(s->tiles);
    EMDEE_TRY(dev_alloc(&s->tiles, (size_t)2 * nt));
    CUDA_TRY(cudaMemcpy(s->tiles, h.data(), sizeof(int32_t) * 2 * nt, cudaMemcpyHostToDevice));
    s->ntiles = nt;
    return EMDEE_OK;
}

// ------------------------------------------------------------------------------------------------
// binning
// ------------------------------------------------------------------------------------------------
// max over bricks of the number of atoms in the brick plus its halo (shared-memory sizing)
__global__ void k_brick_max(GridDesc g, const int32_t *__restrict__ cell_start, int *__restrict__ out)
{
    const int b0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (b0 >= g.nbx * g.nby * g.nbz) return;
    int b = b0;
    const int bxi = b % g.nbx; b /= g.nbx;
    const int byi = b % g.nby;
    const int bzi = b / g.nby;
    const int M = g.M, R = g.R;
    const int hx0 = bxi * g.bx, hy0 = byi * g.by, hz0 = g.zhome0 + bzi * g.bz;
    const int nhx = min(g.bx, M - hx0), nhy = min(g.by, M - hy0), nhz = min(g.bz, g.zhome0 + g.nzhome - hz0);
    int total = 0, rowmax = 0;
    for (int cz = 0; cz < nhz + 2 * R; cz++)
        for (int cy = 0; cy < nhy + 2 * R; cy++) {
            int lz = hz0 - R + cz;
            if (g.zwrap) lz = wrap_mod(lz, M);
            const int gy = wrap_mod(hy0 - R + cy, M);
            int row = 0;
            for (int cx = 0; cx < nhx + 2 * R; cx++) {
                const int lc = wrap_mod(hx0 - R + cx, M) + M * (gy + M * lz);
                row += cell_start[lc + 1] - cell_start[lc];
            }
            total += row;
            rowmax = max(rowmax, row);
        }
    atomicMax(out, total);
    atomicMax(out + 1, rowmax);        // longest staged (y, z) row: sizes the raw ring of the TMA staging
}

static int exclusive_scan(emdee_system *s, int32_t *data, int64_t n, int32_t *maxval)
{
    emdee_ctx *c = s->ctx;
    const int64_t per = SCAN_BLOCK * SCAN_ITEMS;
    const int64_t nb = ceil_div64(n, per);
    if (nb > per) EMDEE_FAIL(EMDEE_ERR_CAPACITY, "exclusive_scan: %lld entries exceed the two-level scan", (long long)n);
    k_scan_block<<<(unsigned)nb, SCAN_BLOCK, 0, c->stream>>>(data, data, n, s->block_sum, maxval);
    c->launches++;
    if (nb > 1) {
        k_scan_block<<<1, SCAN_BLOCK, 0, c->stream>>>(s->block_sum, s->block_sum, nb, nullptr, nullptr);
        c->launches++;
        LAUNCH_1D(c, k_scan_add, n, data, n, s->block_sum);
    }
    return check_launch("exclusive_scan");
}

// Picks the home-brick shape and the block size.  Shared memory (staged atoms + per-lane stacks) limits
// residency, so the score is the number of warps per SM that actually hold a task:
//   blocks/SM(shared memory, registers) x min(warps per block, warp tasks per brick).
// The staged-atom capacity is the exact maximum over all bricks (k_brick_max), so a launch never
// overflows.  The previous choice is kept across re-binnings while it still fits.
static int brick_capacity(emdee_system *s, int *cap_out, int *rowmax_out)
{
    emdee_ctx *c = s->ctx;
    GridDesc &g = s->g;
    CUDA_TRY(cudaMemsetAsync(s->brick_max, 0, 2 * sizeof(int), c->stream));
    const int nb = g.nbx * g.nby * g.nbz;
    LAUNCH_1D(c, k_brick_max, (int64_t)nb, g, s->cell_start, s->brick_max);
    int mx[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(mx, s->brick_max, sizeof(mx), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *cap_out = std::max(64, (mx[0] + 4) & ~3);
    *rowmax_out = mx[1];
    return check_launch("k_brick_max");
}
// the same for compacted staging: atoms within rc + skin of a brick's home box, largest count over all bricks
static int brick_capacity_compact(emdee_system *s, int *cap_out)
{
    emdee_ctx *c = s->ctx;
    GridDesc &g = s->g;
    AtomArrays &A = s->A[s->cur];
    CellArgs a = {};
    a.g = g;
    a.cell_start = s->cell_start;
    a.sx = A.s[0]; a.sy = A.s[1]; a.sz = A.s[2];
    a.L = s->L;
    a.cell_edge = s->L / g.M;
    const double rl = s->cutoff + s->skin;
    a.keep2 = rl * rl * (1.0 + 1e-6);
    a.block_split = 0x7fffffff; a.block_split2 = 0x7fffffff;
    CUDA_TRY(cudaMemsetAsync(s->brick_max, 0, 2 * sizeof(int), c->stream));
    const int nb = g.nbx * g.nby * g.nbz;
    k_brick_keep_max<<<nb, 256, 0, c->stream>>>(a, s->brick_max);
    c->launches++;
    int mx = 0;
    CUDA_TRY(cudaMemcpyAsync(&mx, s->brick_max, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *cap_out = mx >= (1 << 30) ? 1 << 30 : std::max(64, (mx + 4) & ~3);
    return check_launch("k_brick_keep_max");
}
static void set_brick_shape(emdee_system *s, const int sh[3])
{
    GridDesc &g = s->g;
    const int R = g.R, M = g.M;
    g.bx = std::max(1, std::min(sh[0], M - 2 * R));
    g.by = std::max(1, std::min(sh[1], M - 2 * R));
    g.bz = std::max(1, std::min(sh[2], g.zwrap ? M - 2 * R : g.nzhome));
    g.nbx = (M + g.bx - 1) / g.bx;
    g.nby = (M + g.by - 1) / g.by;
    g.nbz = (g.nzhome + g.bz - 1) / g.bz;
}
static bool list_capable(const emdee_system *s)
{
    // the pair-list kernels read LJ parameters from the class table and index staged atoms with 16 bits
    return s->grid_ok && s->use_list && s->skin > 0 && s->ntypes > 0;
}
// A single-point evaluation can use the list kernels at any skin (the list is built at the positions it is used at);
// energies and virials exist only in the persistent kernel.
static bool single_point_list(const emdee_system *s)
{
    return s->grid_ok && s->use_list && s->ntypes > 0 && s->fl_persistent;
}
static int choose_bricks(emdee_system *s)
{
    emdee_ctx *c = s->ctx;
    GridDesc &g = s->g;
    const int R = g.R;
    const bool typed = s->ntypes > 0;
    const bool listed = s->use_list && s->ntypes > 0;    // the list kernels also serve single-point evaluations (any skin)
    const size_t per_sm = c->smem_optin + 1024;          // usable shared memory per SM (1 KB reserved per block)
    const int64_t ntot = s->nlo + s->nown + s->nhi;
    const double per_cell = (double)ntot / (double)std::max<int64_t>(1, (int64_t)g.M * g.M * g.nzt);
    int maxpop = s->pre_maxpop;
    if (!s->pre_valid) {
        CUDA_TRY(cudaMemcpyAsync(&maxpop, s->maxpop, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    static const int blocks[] = {384, 256, 192, 128, 64};
    static const int lblocks[] = {256, 192, 128, 96, 64};
    // largest k_force_cells block that fits next to `cap` staged atoms (0: none)
    auto cells_block = [&](int cap, int ncs, int forced_block) -> int {
        for (int block : blocks) {
            if (forced_block && block != forced_block) continue;
            if (fc_smem_bytes(cap, ncs, block, typed) <= c->smem_optin) return block;
        }
        return 0;
    };
    // Do two staging buffers of the persistent list kernel fit?  As they are; or, as a last resort (allow_compact: no brick shape
    // fits otherwise -- dense cells), with shallower per-lane stacks, then (one GPU, no TMA / Newton's-third-law variant) with
    // only the atoms within rc + skin of the home box staged.
    struct ListFit { bool ok; int cap, qcap; bool compact; };
    auto list_fit = [&](int cap_full, int ncs, bool allow_compact, ListFit *out) -> int {
        const int nt = std::max(s->ntypes, 1);
        *out = ListFit{false, cap_full, FL_QCAP, false};
        if (!s->want_persistent) return EMDEE_OK;
        for (int q : {FL_QCAP, FLP_QCAP_SMALL})
            if ((q == FL_QCAP || (allow_compact && c->nranks == 1)) && flp_smem_bytes(cap_full, ncs, nt, 2, 0, q) <= c->smem_optin) {
                *out = ListFit{true, cap_full, q, false};
                break;
            }
        if (out->ok && s->want_compact < 2) return EMDEE_OK;
        if (!(allow_compact || s->want_compact >= 2) || !s->want_compact || c->nranks != 1 || s->want_tma || s->want_n3 || ncs > FC_MAX_NCS_SMALL)
            return EMDEE_OK;
        int ccap = 0;
        EMDEE_TRY(brick_capacity_compact(s, &ccap));
        for (int q : {FL_QCAP, FLP_QCAP_SMALL})
            if (ccap < 65534 && (q == FL_QCAP || allow_compact) && flp_smem_bytes(ccap, ncs, nt, 2, 0, q) <= c->smem_optin) {
                *out = ListFit{true, ccap, q, true};
                return EMDEE_OK;
            }
        return EMDEE_OK;      // (what the uncompacted attempt found, if anything)
    };
    auto finish = [&](int cap, int block, int lblock, int rowmax, ListFit lf) -> int {
        s->fc_cap = cap;
        s->fl_cap = lf.ok ? lf.cap : cap;
        s->fl_qcap = lf.ok ? lf.qcap : FL_QCAP;
        s->fl_compact = lf.ok && lf.compact;
        s->fc_rowmax = rowmax;
        // 32-atom groups per brick from the densest cell, with head-room so that density fluctuations between
        // re-binnings do not resize the pair list
        // split lists (two lanes per home atom) where a brick has too few warp tasks to occupy the consumer warps of two buffers
        s->fl_split = lf.ok && s->want_split && !s->want_n3 && !s->want_tma && c->nranks == 1 &&
                      std::ceil(1.03 * per_cell * g.bx * g.by * g.bz / 32.0) <= FLP_NCONS / 2;
        const int gmax = ((g.bx * g.by * g.bz * (std::max(maxpop, 1) + 8) << (s->fl_split ? 1 : 0)) + 31) / 32 + 1;
        if (gmax > s->fc_gmax || gmax * 2 < s->fc_gmax) s->fc_gmax = gmax;
        s->list_valid = false;
        s->fc_ncs = (g.bx + 2 * R) * (g.by + 2 * R) * (g.bz + 2 * R);
        s->fc_block = block;
        s->fc_smem = fc_smem_bytes(cap, s->fc_ncs, block, typed);
        s->fl_block = lblock;
        s->fl_smem = fl_smem_bytes(cap, s->fc_ncs, lblock, std::max(s->ntypes, 1));
        s->fl_persistent = lf.ok;
        // TMA staging: the x image of a staged atom is resolved against its segment's centre, so a segment must span less than
        // half the box; the raw ring (two groups of whole rows) must fit behind the stacks
        s->fl_tma = false;
        if (s->fl_persistent && !s->fl_compact && s->fl_qcap == FL_QCAP && s->want_tma && !s->want_n3 && 2 * (g.bx + 2 * R + 1) <= g.M && rowmax > 0) {
            const int nrows = (g.by + 2 * R) * (g.bz + 2 * R), rowpad = (rowmax + 5) & ~1;
            const size_t base = flp_smem_bytes(cap, s->fc_ncs, std::max(s->ntypes, 1), 2) + 2 * (size_t)(2 * nrows) * sizeof(int4) + 64;
            if (base < c->smem_optin) {
                const int r = (int)std::min<size_t>((c->smem_optin - base) / (2 * 3 * sizeof(double) * (size_t)rowpad), (size_t)nrows);
                if (r >= 1 && (int64_t)r * rowpad < 65536) {
                    s->fl_tma = true;
                    s->segcap = 2 * nrows; s->raw_rows = r; s->rawlen = r * rowpad;
                }
            }
        }
        s->fc_typed = typed;
        s->fc_nblocks = g.nbx * g.nby * g.nbz;
        return EMDEE_OK;
    };
    // keep the previous configuration while it fits and the cells hold about as many atoms as when it was chosen
    // (one tiny kernel + sync per re-binning)
    if (s->fc_shape[0] > 0 && !getenv("EMDEE_BRICK") && std::fabs(per_cell - s->fc_per_cell) <= 0.08 * s->fc_per_cell && R == s->fc_R) {
        set_brick_shape(s, s->fc_shape);
        int cap = s->pre_cap, rowmax = s->pre_rowmax;      // (a slab re-binning measured them for this shape before its one synchronisation)
        if (!s->pre_valid) EMDEE_TRY(brick_capacity(s, &cap, &rowmax));
        const int ncs = (g.bx + 2 * R) * (g.by + 2 * R) * (g.bz + 2 * R);
        const size_t need = listed ? fl_smem_bytes(cap, ncs, s->fl_block, std::max(s->ntypes, 1)) : fc_smem_bytes(cap, ncs, s->fc_block, typed);
        ListFit lf = {false, cap, FL_QCAP, false};
        if (listed) EMDEE_TRY(list_fit(cap, ncs, (s->fl_compact && s->want_compact < 2) || s->fl_qcap != FL_QCAP, &lf));
        if (cap <= 65534 && need <= s->fc_smem_budget && fc_smem_bytes(cap, ncs, s->fc_block, typed) <= c->smem_optin &&
            (!listed || (lf.ok == s->fl_persistent && lf.compact == s->fl_compact && lf.qcap == s->fl_qcap)))
            return finish(cap, s->fc_block, s->fl_block, rowmax, lf);
    }
    // shapes in units of ndiv cells (a cell edge is (rc + skin)/ndiv), so the candidates keep their physical size
    static const int base_shapes[][3] = {{8, 2, 2}, {4, 4, 2}, {7, 2, 2}, {5, 3, 2}, {6, 2, 2}, {4, 3, 2}, {5, 2, 2}, {3, 3, 2}, {8, 2, 1}, {4, 4, 1},
                                         {4, 2, 2}, {3, 2, 2}, {4, 2, 1}, {2, 2, 2}, {8, 1, 1}, {4, 1, 1}, {2, 2, 1}, {2, 1, 1}, {1, 1, 1}};
    constexpr int NBASE = sizeof(base_shapes) / sizeof(base_shapes[0]);
    int shapes[2 * NBASE][3];
    int nshapes_all = 0;
    for (int k = 0; k < NBASE; k++, nshapes_all++)
        for (int d = 0; d < 3; d++) shapes[nshapes_all][d] = base_shapes[k][d] * std::max(R, 1);
    if (R > 1)      // finer cells also allow the plain shapes
        for (int k = 0; k < NBASE; k++, nshapes_all++)
            for (int d = 0; d < 3; d++) shapes[nshapes_all][d] = base_shapes[k][d];
    int forced[3] = {0, 0, 0}, forced_block = 0, forced_lblock = 0;
    if (const char *e = getenv("EMDEE_BRICK")) sscanf(e, "%d,%d,%d", &forced[0], &forced[1], &forced[2]);
    if (const char *e = getenv("EMDEE_BLOCK")) forced_block = atoi(e);
    if (const char *e = getenv("EMDEE_LBLOCK")) forced_lblock = atoi(e);
    double best_score = -1;
    int best_shape[3] = {0, 0, 0}, best_block = 0, best_lblock = 192, best_cap = 0, best_rowmax = 0;
    size_t best_budget = 0;
    const int nshape = forced[0] > 0 ? 1 : nshapes_all;
    ListFit best_lf = {false, 0, FL_QCAP, false};
    // second pass (dense cells): no shape fits the persistent kernel as it is -- allow shallower stacks and compacted staging
    for (int pass = 0; pass < 2; pass++) {
    if (pass == 1 && (!listed || !s->want_persistent || best_lf.ok)) break;
    int prev[3] = {-1, -1, -1};
    for (int k = 0; k < nshape; k++) {
        const int *sh = forced[0] > 0 ? forced : shapes[k];
        set_brick_shape(s, sh);
        if (g.by * g.bz > FC_MAX_HOMEROWS) continue;
        if (std::max(g.bx, std::max(g.by, g.bz)) + 2 * R > 32) continue;       // ctab / ccoord hold 32 cells per dimension
        if (g.bx == prev[0] && g.by == prev[1] && g.bz == prev[2]) continue;   // clipped to the same shape as the previous one
        prev[0] = g.bx; prev[1] = g.by; prev[2] = g.bz;
        int cap = 0, rowmax = 0;
        EMDEE_TRY(brick_capacity(s, &cap, &rowmax));
        if (cap > 65534) continue;
        const int ncs = (g.bx + 2 * R) * (g.by + 2 * R) * (g.bz + 2 * R);
        const double home = per_cell * g.bx * g.by * g.bz;
        if (listed) {
            // the stepping kernel decides, as long as k_force_cells (single-point evaluations) fits too
            const int cblock = cells_block(cap, ncs, forced_block);
            if (!cblock) continue;
            ListFit lf;
            EMDEE_TRY(list_fit(cap, ncs, pass == 1, &lf));
            if (pass == 1 && !lf.ok) continue;
            if (lf.ok) {
                // persistent kernel (two staging buffers).  Cost model per home atom, fitted to B200 measurements
                // (4x2x2 / 3x2x2 / 4x3x2 / 5x2x2 bricks at skins 0.35-0.5): the consumers' time is 1 / fill, where a brick period
                // lasts ceil(groups / consumer warps) warp-task times (13 groups on 12 consumers: 1.55 ms per launch against
                // 1.31 ms for 11-12), plus ~2.5 % per staged cell per home cell for the producers' staging
                const double groups = std::ceil(1.03 * home / 32.0);
                const double fill = groups / (std::ceil(groups / FLP_NCONS) * FLP_NCONS);
                // (a factor for bricks that stick out of the grid -- a slab of 7 planes cut into layers of 2 -- was tried at 8 GPUs: it
                // picks 4x4x1 bricks there, 0.252 ms per launch against 0.240 ms for 4x2x2 with a half-filled last layer: not kept)
                const double score = 1e6 + 1000.0 / (0.025 * ncs / (g.bx * g.by * g.bz) * ((double)lf.cap / cap) + 1.0 / fill);
                if (score > best_score) {
                    best_lf = lf;
                    best_score = score; best_block = cblock; best_lblock = 192; best_cap = cap; best_rowmax = rowmax;
                    best_shape[0] = g.bx; best_shape[1] = g.by; best_shape[2] = g.bz;
                    best_budget = c->smem_optin;
                }
                continue;
            }
            const double groups = std::max(1.0, std::ceil(home / 32.0));
            for (int lblock : lblocks) {
                if (forced_lblock && lblock != forced_lblock) continue;
                const size_t smem = fl_smem_bytes(cap, ncs, lblock, std::max(s->ntypes, 1));
                if (smem > c->smem_optin) continue;
                const int by_smem = (int)(per_sm / (smem + 1024));
                const int by_regs = std::max(1, 65536 / (FL_REGS_ESTIMATE * lblock));
                const int resident = std::min(std::min(by_smem, by_regs), 32);
                if (resident < 1) continue;
                const int nw = lblock / 32;
                // rounds of tasks per warp: a block lives ceil(groups / warps) task times, idle slots are lost
                const double rounds = std::ceil(groups / nw);
                const double busy = groups / (rounds * nw);
                const double active = std::min(16.0, resident * nw * busy);
                const double score = active * 1000.0 + 100.0 * std::min(resident, 3) + 2.0 * (g.bx * g.by * g.bz);
                if (score > best_score) {
                    best_score = score; best_block = cblock; best_lblock = lblock; best_cap = cap; best_rowmax = rowmax;
                    best_shape[0] = g.bx; best_shape[1] = g.by; best_shape[2] = g.bz;
                    best_budget = per_sm / resident - 1024;
                }
            }
            continue;
        }
        const double tasks = g.by * g.bz * std::max(1.0, std::ceil(per_cell * g.bx / 32.0));
        for (int block : blocks) {
            if (forced_block && block != forced_block) continue;
            const size_t smem = fc_smem_bytes(cap, ncs, block, typed);
            if (smem > c->smem_optin) continue;
            const int by_smem = (int)(per_sm / (smem + 1024));
            const int by_regs = std::max(1, 65536 / (FC_REGS_ESTIMATE * block));
            const int resident = std::min(std::min(by_smem, by_regs), 32);
            if (resident < 1) continue;
            const double active = resident * std::min<double>(block / 32, tasks);
            // ties: larger bricks first (measured on B200: 8x2x2 @384 threads beats two resident 4x2x1 blocks by 1.3x --
            // the halo staged per home atom drops from 9x to 5x), then fewer idle warps
            const double score = active * 1000.0 + 2.0 * (g.bx * g.by * g.bz) - 0.1 * resident * (block / 32);
            if (score > best_score) {
                best_score = score; best_block = block; best_cap = cap; best_rowmax = rowmax;
                best_shape[0] = g.bx; best_shape[1] = g.by; best_shape[2] = g.bz;
                best_budget = per_sm / resident - 1024;
            }
        }
    }
    }
    if (best_score < 0) EMDEE_FAIL(EMDEE_ERR_CAPACITY, "emdee_bin: no brick shape fits shared memory (cells too populated); use a larger ndiv");
    set_brick_shape(s, best_shape);
    if (getenv("EMDEE_DEBUG"))
        fprintf(stderr, "[emdee] bricks %dx%dx%d block %d list-block %d cap %d (full search, per_cell %.1f, maxpop %d, listed %d; persistent %d: cap %d, stack %d, compacted %d)\n",
                best_shape[0], best_shape[1], best_shape[2], best_block, best_lblock, best_cap, per_cell, maxpop, (int)listed, (int)best_lf.ok, best_lf.cap, best_lf.qcap,
                (int)best_lf.compact);
    for (int k = 0; k < 3; k++) s->fc_shape[k] = best_shape[k];
    s->fc_per_cell = per_cell;
    s->fc_R = R;
    s->fc_smem_budget = std::min(best_budget, c->smem_optin);
    return finish(best_cap, best_block, best_lblock, best_rowmax, best_lf);
}

static int do_bin(emdee_system *s, int ndiv)
{
    emdee_ctx *c = s->ctx;
    s->inner_at_current = false;
    AtomArrays &A = s->A[s->cur];
    // cells_per_dimension(L, cutoff, ndiv) = floor(Int32, ndiv*L/cutoff), src/cells.jl:36 (cutoff + skin here)
    const double Mf = std::floor((double)ndiv * s->L / (s->cutoff + s->skin));
    if (!(Mf >= 1) || Mf > 2000) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_bin: M=%g cells per dimension out of range", Mf);
    const int M = (int)Mf;
    GridDesc &g = s->g;
    g.M = M;
    g.R = ndiv;
    g.zwrap = 1;
    g.nzt = M;
    g.zhome0 = 0;
    g.nzhome = M;
    g.zglob0 = 0;
    s->ndiv = ndiv;
    s->grid_ok = (M >= 2 * g.R + 1);   // otherwise a neighbourhood would hold a cell twice
    s->ncell = (int64_t)M * M * g.nzt;
    if (s->ncell + 1 > s->ncell_cap) {
        dev_free(s->count); dev_free(s->cell_start); dev_free(s->fill); dev_free(s->block_sum);
        s->ncell_cap = s->ncell + 1;
        EMDEE_TRY(dev_alloc(&s->count, (size_t)s->ncell_cap));
        EMDEE_TRY(dev_alloc(&s->cell_start, (size_t)s->ncell_cap));
        EMDEE_TRY(dev_alloc(&s->fill, (size_t)s->ncell_cap));
        EMDEE_TRY(dev_alloc(&s->block_sum, (size_t)ceil_div64(s->ncell_cap, SCAN_BLOCK * SCAN_ITEMS) + 1));
    }
    const int64_t n = s->nown;
    CUDA_TRY(cudaMemsetAsync(s->cell_start, 0, sizeof(int32_t) * (s->ncell + 1), c->stream));
    CUDA_TRY(cudaMemsetAsync(s->fill, 0, sizeof(int32_t) * (s->ncell + 1), c->stream));
    CUDA_TRY(cudaMemsetAsync(s->maxpop, 0, sizeof(int32_t), c->stream));
    LAUNCH_1D(c, k_cell_index, n, s->nlo, n, A.s[0], A.s[1], A.s[2], M, g.zglob0, g.nzt, g.zwrap, s->gcell[s->cur],
              s->lcell[s->cur], s->cell_start, s->err);
    CUDA_TRY(cudaMemcpyAsync(s->count, s->cell_start, sizeof(int32_t) * (s->ncell + 1), cudaMemcpyDeviceToDevice, c->stream));
    EMDEE_TRY(exclusive_scan(s, s->cell_start, s->ncell + 1, s->maxpop));
    LAUNCH_1D(c, k_scatter, n, s->nlo, n, s->lcell[s->cur], s->cell_start, s->fill, s->order);
    LAUNCH_1D(c, k_rank_in_cell, n, s->nlo, n, s->order, s->lcell[s->cur], s->cell_start, A.id, s->src_of_new);
    GatherArgs ga;
    ga.pfirst = s->nlo;
    ga.n = n;
    ga.src_of_new = s->src_of_new;
    ga.gcell_old = s->gcell[s->cur];
    ga.lcell_old = s->lcell[s->cur];
    ga.in = A;
    ga.out = s->A[1 - s->cur];
    ga.gcell_new = s->gcell[1 - s->cur];
    ga.lcell_new = s->lcell[1 - s->cur];
    ga.slot_of_id = s->slot_of_id;
    ga.has_vel = 1;
    ga.has_excl = 1;
    LAUNCH_1D(c, k_gather, n, ga);
    EMDEE_TRY(check_launch("binning"));
    s->cur = 1 - s->cur;
    s->binned = true;
    s->order_valid = true;
    s->steps_since_bin = 0;
    s->forces_valid = false;
    s->last_bitmask = 0;        // f / en / vir are in the previous slot order: the getters fail until the next compute
    CUDA_TRY(cudaMemsetAsync(s->maxd2, 0, sizeof(unsigned), c->stream));
    if (s->grid_ok) EMDEE_TRY(choose_bricks(s));
    return EMDEE_OK;
}


// ------------------------------------------------------------------------------------------------
// slab decomposition: re-binning with migration and ghost exchange, per-step halo
// ------------------------------------------------------------------------------------------------
// Message order inside every NCCL group: sends [to lower, to upper], receives [from upper, from lower].
// With two ranks both neighbours are the same peer and NCCL pairs the k-th send with the k-th receive,
// so the peer's "to lower" slice (its bottom planes) lands in my upper ghost range, as it must.
static int slab_exchange(emdee_ctx *c, cudaStream_t st, const void *send_lo, size_t n_send_lo, const void *send_hi,
                         size_t n_send_hi, void *recv_lo, size_t n_recv_lo, void *recv_hi, size_t n_recv_hi, size_t elem)
{
    const int lower = (c->rank + c->nranks - 1) % c->nranks, upper = (c->rank + 1) % c->nranks;
    NCCL_TRY(ncclSend(send_lo, n_send_lo * elem, ncclInt8, lower, c->comm, st));
    NCCL_TRY(ncclSend(send_hi, n_send_hi * elem, ncclInt8, upper, c->comm, st));
    NCCL_TRY(ncclRecv(recv_hi, n_recv_hi * elem, ncclInt8, upper, c->comm, st));
    NCCL_TRY(ncclRecv(recv_lo, n_recv_lo * elem, ncclInt8, lower, c->comm, st));
    return EMDEE_OK;
}

// Scaled positions of the boundary planes -> neighbours' ghost ranges (three contiguous array slices
// per side, no packing).  Runs on `st`; callers order it against the compute stream with events.
static int slab_halo_positions(emdee_system *s, cudaStream_t st)
{
    emdee_ctx *c = s->ctx;
    AtomArrays &A = s->A[s->cur];
    const int64_t own_end = s->nlo + s->nown;
    NCCL_TRY(ncclGroupStart());
    for (int k = 0; k < 3; k++)
        EMDEE_TRY(slab_exchange(c, st, A.s[k] + s->lo_send_a, s->lo_send_n, A.s[k] + s->hi_send_a, s->hi_send_n,
                                A.s[k], s->nlo, A.s[k] + own_end, s->nhi, sizeof(double)));
    NCCL_TRY(ncclGroupEnd());
    return EMDEE_OK;
}

// Collective over the slab ranks (first re-binning): every rank exports its position pool with one CUDA IPC handle, hands it to
// both ring neighbours through NCCL, and maps theirs.  All ranks use the peer-mapped halo or none does (min over ranks).
static int ensure_peer_mapping(emdee_system *s)
{
    emdee_ctx *c = s->ctx;
    if (s->peer_tried) return EMDEE_OK;
    s->peer_tried = true;
    const char *e = getenv("EMDEE_P2P");
    int ok = (e && atoi(e) == 0) ? 0 : 1;
    cudaIpcMemHandle_t mine, from[2];
    memset(&mine, 0, sizeof(mine));
    if (ok && cudaIpcGetMemHandle(&mine, s->spool) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    EMDEE_TRY(ensure_tmp(s, 512));
    char *d = reinterpret_cast<char *>(s->tmp);
    CUDA_TRY(cudaMemcpyAsync(d, &mine, 64, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(ncclGroupStart());
    EMDEE_TRY(slab_exchange(c, c->stream, d, 64, d, 64, d + 64, 64, d + 128, 64, 1));
    NCCL_TRY(ncclGroupEnd());
    CUDA_TRY(cudaMemcpyAsync(&from[0], d + 64, 64, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(&from[1], d + 128, 64, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const int lower = (c->rank + c->nranks - 1) % c->nranks, upper = (c->rank + 1) % c->nranks;
    if (ok && cudaIpcOpenMemHandle(&s->peer_pool[0], from[0], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; s->peer_pool[0] = nullptr; cudaGetLastError(); }
    if (ok) {
        if (lower == upper) s->peer_pool[1] = s->peer_pool[0];
        else if (cudaIpcOpenMemHandle(&s->peer_pool[1], from[1], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; s->peer_pool[1] = nullptr; cudaGetLastError(); }
    }
    int32_t *flag = reinterpret_cast<int32_t *>(d + 256);
    int32_t h = ok;
    CUDA_TRY(cudaMemcpyAsync(flag, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(g_nccl.AllReduce(flag, flag, 1, ncclInt32, ncclMin, c->comm, c->stream));
    CUDA_TRY(cudaMemcpyAsync(&h, flag, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    s->peer_ok = h != 0;
    if (getenv("EMDEE_DEBUG"))
        fprintf(stderr, "[emdee] rank %d: peer-mapped halos %s (mine %d, all ranks %d)\n", c->rank, s->peer_ok ? "on" : "off (ncclSend/ncclRecv halos)", ok, h);
    return EMDEE_OK;
}

// EMDEE_DEBUG=2: wall-clock anatomy of the slab re-binnings (a stream synchronisation at every phase boundary: perturbs the run)
static double g_rebin_phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};
static int g_rebin_count = 0;
static double wall_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
#define REBIN_PHASE(k)                                                   \
    if (phase_dbg) {                                                     \
        CUDA_TRY(cudaStreamSynchronize(c->stream));                      \
        const double now_ = wall_ms();                                   \
        g_rebin_phase[k] += now_ - phase_t0;                             \
        if (c->rank == 0) fprintf(stderr, "[emdee] re-binning %d phase %d: %.3f ms\n", g_rebin_count, k, now_ - phase_t0); \
        phase_t0 = now_;                                                 \
    }

static int do_bin_slab(emdee_system *s, int ndiv)
{
    emdee_ctx *c = s->ctx;
    const int G = c->nranks;
    EMDEE_TRY(ensure_peer_mapping(s));
    const char *dbg_ = getenv("EMDEE_DEBUG");
    const bool phase_dbg = dbg_ && atoi(dbg_) >= 2;
    if (phase_dbg) CUDA_TRY(cudaStreamSynchronize(c->stream));
    double phase_t0 = wall_ms();
    g_rebin_count++;
    const double Mf = std::floor((double)ndiv * s->L / (s->cutoff + s->skin));
    if (!(Mf >= 1) || Mf > 2000) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_bin: M=%g cells per dimension out of range", Mf);
    const int M = (int)Mf, R = ndiv;
    if (M < 2 * R + 1) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_bin: M=%d is too small for a cell grid; slab decomposition needs M >= %d", M, 2 * R + 1);
    if (M / G < std::max(R, 1)) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_bin: %d z planes over %d ranks leaves fewer than %d planes per slab", M, G, R);
    const int z0 = (int)((int64_t)c->rank * M / G), z1 = (int)((int64_t)(c->rank + 1) * M / G), nz = z1 - z0;
    GridDesc &g = s->g;
    g.M = M; g.R = R; g.zwrap = 0; g.nzt = nz + 2 * R; g.zhome0 = R; g.nzhome = nz; g.zglob0 = z0 - R;
    s->ndiv = ndiv; s->grid_ok = true; s->z0 = z0; s->nz = nz;
    const int64_t plane = (int64_t)M * M;
    s->ncell = plane * g.nzt;
    if (s->ncell + 2 > s->ncell_cap) {
        dev_free(s->count); dev_free(s->cell_start); dev_free(s->fill); dev_free(s->block_sum);
        s->ncell_cap = s->ncell + 2;
        EMDEE_TRY(dev_alloc(&s->count, (size_t)s->ncell_cap));
        EMDEE_TRY(dev_alloc(&s->cell_start, (size_t)s->ncell_cap));
        EMDEE_TRY(dev_alloc(&s->fill, (size_t)s->ncell_cap));
        EMDEE_TRY(dev_alloc(&s->block_sum, (size_t)ceil_div64(s->ncell_cap, SCAN_BLOCK * SCAN_ITEMS) + 1));
    }
    AtomArrays &A = s->A[s->cur];
    const bool first_time = !s->decomposed;
    int64_t n_in = s->nown;
    const int64_t first = s->nlo;
    CUDA_TRY(cudaMemsetAsync(s->cell_start, 0, sizeof(int32_t) * (s->ncell + 2), c->stream));
    CUDA_TRY(cudaMemsetAsync(s->fill, 0, sizeof(int32_t) * (s->ncell + 2), c->stream));
    CUDA_TRY(cudaMemsetAsync(s->sendcount, 0, sizeof(int32_t) * 2, c->stream));
    LAUNCH_1D(c, k_cell_index_slab, n_in, first, n_in, A.s[0], A.s[1], A.s[2], M, z0, nz, R, first_time ? 1 : 0,
              (int)s->ncell, s->gcell[s->cur], s->lcell[s->cur], s->cell_start, s->sendcount, s->list_lo, s->list_hi,
              (int)s->cap, s->err);
    if (!first_time) {
        // ---- migration: counts, then payload --------------------------------------------------
        int32_t nsend[2] = {0, 0}, nrecv[2] = {0, 0};
        NCCL_TRY(ncclGroupStart());
        EMDEE_TRY(slab_exchange(c, c->stream, s->sendcount, 1, s->sendcount + 1, 1, s->recvcount, 1, s->recvcount + 1, 1, sizeof(int32_t)));
        NCCL_TRY(ncclGroupEnd());
        CUDA_TRY(cudaMemcpyAsync(nsend, s->sendcount, sizeof(nsend), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(nrecv, s->recvcount, sizeof(nrecv), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        const int64_t need = std::max<int64_t>(std::max(nsend[0], nsend[1]), std::max(nrecv[0], nrecv[1]));
        if (need > s->migcap) {
            for (int k = 0; k < 4; k++) dev_free(s->migbuf[k]);
            s->migcap = need + need / 2 + 1024;
            for (int k = 0; k < 4; k++) EMDEE_TRY(dev_alloc(&s->migbuf[k], (size_t)MIG_FIELDS * s->migcap));
        }
        if (first + n_in + nrecv[0] + nrecv[1] > s->cap) EMDEE_FAIL(EMDEE_ERR_CAPACITY, "emdee_bin: slab capacity exceeded by migrating atoms");
        LAUNCH_1D(c, k_pack_migrants, (int64_t)nsend[0], nsend[0], s->list_lo, A, s->migbuf[0]);
        LAUNCH_1D(c, k_pack_migrants, (int64_t)nsend[1], nsend[1], s->list_hi, A, s->migbuf[1]);
        NCCL_TRY(ncclGroupStart());
        EMDEE_TRY(slab_exchange(c, c->stream, s->migbuf[0], (size_t)MIG_FIELDS * nsend[0], s->migbuf[1], (size_t)MIG_FIELDS * nsend[1],
                                s->migbuf[2], (size_t)MIG_FIELDS * nrecv[0], s->migbuf[3], (size_t)MIG_FIELDS * nrecv[1], sizeof(double)));
        NCCL_TRY(ncclGroupEnd());
        LAUNCH_1D(c, k_unpack_migrants, (int64_t)nrecv[0], nrecv[0], s->migbuf[2], first + n_in, A, s->L, M, z0, nz, R,
                  s->gcell[s->cur], s->lcell[s->cur], s->cell_start, s->err);
        LAUNCH_1D(c, k_unpack_migrants, (int64_t)nrecv[1], nrecv[1], s->migbuf[3], first + n_in + nrecv[0], A, s->L, M, z0, nz, R,
                  s->gcell[s->cur], s->lcell[s->cur], s->cell_start, s->err);
        n_in += nrecv[0] + nrecv[1];
    }
    REBIN_PHASE(0)       // cell index, migration (counts, payload, unpack)
    // ---- populations of the ghost planes come from the neighbours' boundary planes ----------------
    int32_t *cnt = s->cell_start;
    NCCL_TRY(ncclGroupStart());
    EMDEE_TRY(slab_exchange(c, c->stream, cnt + plane * R, (size_t)plane * R, cnt + plane * nz, (size_t)plane * R,
                            cnt, (size_t)plane * R, cnt + plane * (R + nz), (size_t)plane * R, sizeof(int32_t)));
    NCCL_TRY(ncclGroupEnd());
    CUDA_TRY(cudaMemcpyAsync(s->count, s->cell_start, sizeof(int32_t) * (s->ncell + 1), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(s->maxpop, 0, sizeof(int32_t), c->stream));
    EMDEE_TRY(exclusive_scan(s, s->cell_start, s->ncell + 1, s->maxpop));
    int32_t marks[5];
    const int64_t mark_idx[5] = {plane * R, plane * 2 * R, plane * nz, plane * (R + nz), s->ncell};
    for (int k = 0; k < 5; k++)
        CUDA_TRY(cudaMemcpyAsync(&marks[k], s->cell_start + mark_idx[k], sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    // read back in the same synchronisation: the densest cell, the staged-atom capacity of the previous brick shape (what
    // choose_bricks needs when it keeps that shape), and the largest displacement of the interval that ends here
    int pre_maxpop = 0, pre_brickmax[2] = {0, 0};
    unsigned dmax_bits = 0;
    const bool pre = s->fc_shape[0] > 0 && !getenv("EMDEE_BRICK");
    if (pre) {
        set_brick_shape(s, s->fc_shape);
        CUDA_TRY(cudaMemsetAsync(s->brick_max, 0, 2 * sizeof(int), c->stream));
        LAUNCH_1D(c, k_brick_max, (int64_t)(g.nbx * g.nby * g.nbz), g, s->cell_start, s->brick_max);
        CUDA_TRY(cudaMemcpyAsync(pre_brickmax, s->brick_max, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaMemcpyAsync(&pre_maxpop, s->maxpop, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (!first_time) {
        NCCL_TRY(g_nccl.AllReduce(s->maxd2, s->maxd2 + 1, 1, ncclUint32, ncclMax, c->comm, c->stream));
        CUDA_TRY(cudaMemcpyAsync(&dmax_bits, s->maxd2 + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    REBIN_PHASE(1)       // ghost-plane populations, scan, marks
    s->pre_valid = pre;
    s->pre_maxpop = pre_maxpop;
    s->pre_cap = std::max(64, (pre_brickmax[0] + 4) & ~3);
    s->pre_rowmax = pre_brickmax[1];
    if (!first_time && s->skin > 0) {
        float d2;
        memcpy(&d2, &dmax_bits, 4);
        s->last_dmax_ratio = std::sqrt((double)d2) / (0.5 * s->skin);
        // next interval: one step longer while a linear (ballistic) extrapolation of this interval's largest displacement stays
        // below 85 % of skin/2; shorter when this one came close.  The skin check in the integrator stays armed.
        const int k = std::max<int64_t>(1, s->steps_since_bin);
        if (s->last_dmax_ratio > 0 && k >= s->slab_interval) {
            if (s->last_dmax_ratio * (k + 1) / k < 0.85) s->slab_interval = k + 1;
            else if (s->last_dmax_ratio > 0.92) s->slab_interval = std::max(1, k - 1);
            else s->slab_interval = k;
        }
        if (getenv("EMDEE_DEBUG") && c->rank == 0)
            fprintf(stderr, "[emdee] slab re-binning after %d steps: max displacement %.3f of skin/2, next interval %d\n", k, s->last_dmax_ratio, s->slab_interval);
    }
    const int64_t nlo = marks[0], own_end = marks[3], ntot = marks[4];
    if (ntot > s->cap) EMDEE_FAIL(EMDEE_ERR_CAPACITY, "emdee_bin: slab holds %lld atoms incl. ghosts, capacity %lld", (long long)ntot, (long long)s->cap);
    const int64_t nown = own_end - nlo;
    LAUNCH_1D(c, k_scatter_slab, n_in, first, n_in, s->lcell[s->cur], (int)s->ncell, s->cell_start, s->fill, s->order);
    LAUNCH_1D(c, k_rank_in_cell, nown, nlo, nown, s->order, s->lcell[s->cur], s->cell_start, A.id, s->src_of_new);
    GatherArgs ga;
    ga.pfirst = nlo; ga.n = nown; ga.src_of_new = s->src_of_new;
    ga.gcell_old = s->gcell[s->cur]; ga.lcell_old = s->lcell[s->cur];
    ga.in = A; ga.out = s->A[1 - s->cur];
    ga.gcell_new = s->gcell[1 - s->cur]; ga.lcell_new = s->lcell[1 - s->cur];
    ga.slot_of_id = nullptr; ga.has_vel = 1; ga.has_excl = 1;
    LAUNCH_1D(c, k_gather, nown, ga);
    EMDEE_TRY(check_launch("slab binning"));
    REBIN_PHASE(2)       // scatter, rank in cell, gather
    s->cur = 1 - s->cur;
    s->nlo = nlo; s->nown = nown; s->nhi = ntot - own_end;
    s->lo_send_a = marks[0]; s->lo_send_n = marks[1] - marks[0];
    s->hi_send_a = marks[2]; s->hi_send_n = marks[3] - marks[2];
    s->decomposed = true;
    // ---- ghost atoms: static per-atom data and current scaled positions ----------------------------
    AtomArrays &B = s->A[s->cur];
    {   // one packed message per neighbour (pack -> ncclSend/ncclRecv -> unpack)
        const int64_t need = std::max(std::max(s->lo_send_n, s->hi_send_n), std::max(s->nlo, s->nhi));
        if (need > s->ghostcap) {
            for (int k = 0; k < 4; k++) dev_free(s->ghostbuf[k]);
            s->ghostcap = need + need / 4 + 1024;
            for (int k = 0; k < 4; k++) EMDEE_TRY(dev_alloc(&s->ghostbuf[k], (size_t)GHOST_FIELDS * s->ghostcap));
        }
        LAUNCH_1D(c, k_pack_ghosts, s->lo_send_n, s->lo_send_a, (int)s->lo_send_n, B, s->ghostbuf[0]);
        LAUNCH_1D(c, k_pack_ghosts, s->hi_send_n, s->hi_send_a, (int)s->hi_send_n, B, s->ghostbuf[1]);
    }
    {   // what the neighbours need to address my ghost slots: [0] my first upper-ghost slot, [1] which pool array holds the current
        // positions (the buffer rotation must be the same on every rank)
        const long long info[4] = {(long long)own_end, (long long)((B.s[0] - (s->spool + 64)) / (ptrdiff_t)s->spool_stride), 0, 0};
        CUDA_TRY(cudaMemcpyAsync(s->peerinfo, info, sizeof(info), cudaMemcpyHostToDevice, c->stream));
    }
    NCCL_TRY(ncclGroupStart());
    EMDEE_TRY(slab_exchange(c, c->stream, s->ghostbuf[0], (size_t)GHOST_FIELDS * s->lo_send_n, s->ghostbuf[1], (size_t)GHOST_FIELDS * s->hi_send_n,
                            s->ghostbuf[2], (size_t)GHOST_FIELDS * s->nlo, s->ghostbuf[3], (size_t)GHOST_FIELDS * s->nhi, sizeof(double)));
    EMDEE_TRY(slab_exchange(c, c->stream, s->peerinfo, 4, s->peerinfo, 4, s->peerinfo + 4, 4, s->peerinfo + 8, 4, sizeof(long long)));
    NCCL_TRY(ncclGroupEnd());
    LAUNCH_1D(c, k_unpack_ghosts, s->nlo, (int64_t)0, (int)s->nlo, s->ghostbuf[2], B);
    LAUNCH_1D(c, k_unpack_ghosts, s->nhi, own_end, (int)s->nhi, s->ghostbuf[3], B);
    EMDEE_TRY(check_launch("ghost atoms"));
    REBIN_PHASE(3)       // ghost atoms
    s->ghost_state = 0;
    s->binned = true;
    s->steps_since_bin = 0;
    s->forces_valid = false;
    s->last_bitmask = 0;
    CUDA_TRY(cudaMemsetAsync(s->maxd2, 0, sizeof(unsigned), c->stream));
    {
        const int rc_ = choose_bricks(s);
        s->pre_valid = false;
        EMDEE_TRY(rc_);
    }
    REBIN_PHASE(4)       // brick configuration
    // brick layers whose halo reaches ghost planes (they must wait for the halo exchange)
    const int bz = s->g.bz, nbz = s->g.nbz;
    s->brick_lo_end = std::min(nbz, (R + bz - 1) / bz);
    s->brick_hi_begin = std::max(s->brick_lo_end, (nz - R) / bz);
    s->hi_layer0 = std::max(0, (nz - R) / bz);
    if (getenv("EMDEE_DEBUG")) {     // the buffer rotation is the same on every rank (the peers address my arrays by it)
        long long h[12];
        CUDA_TRY(cudaMemcpyAsync(h, s->peerinfo, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (h[5] != h[1] || h[9] != h[1]) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_bin: position-buffer rotation differs between slab ranks (%lld, %lld, %lld)", h[5], h[1], h[9]);
    }
    return EMDEE_OK;
}

extern "C" int emdee_bin(emdee_system *s, int ndiv)
{
    SYS_ENTER(s, "emdee_bin");
    if (!s->has_model || !s->has_pos) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_bin: set the model (cutoff) and positions first");
    if (ndiv < 1 || ndiv > 4) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_bin: ndiv=%d must be in [1,4]", ndiv);
    if (s->kick_pending) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_bin: a velocity-Verlet step is half-finished");
    return c->nranks > 1 ? do_bin_slab(s, ndiv) : do_bin(s, ndiv);
}

// update_cells!(cells, r, L), src/cells.jl:196-222, after new positions were set: recompute every atom's cell; if no atom
// changed cell the sorted order, the cell table and (while no atom has moved more than skin/2 since the binning) the pair
// list stay as they are -- no sort, no gather, no list build; otherwise the movers are re-linked by a full (cell, id) re-sort
// (atoms are stored densely in cell order, so one mover shifts everything behind it).  *movers: atoms whose cell changed.
extern "C" int emdee_update_cells(emdee_system *s, int64_t *movers)
{
    SYS_ENTER(s, "emdee_update_cells");
    if (!s->has_model || !s->has_pos) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_update_cells: set the model (cutoff) and positions first");
    if (s->ndiv < 1) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_update_cells: call emdee_bin once first (it fixes ndiv)");
    if (s->kick_pending) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_update_cells: a velocity-Verlet step is half-finished");
    if (movers) *movers = -1;
    if (c->nranks > 1 || !s->order_valid) return emdee_bin(s, s->ndiv);          // slabs migrate atoms: always the full path
    EMDEE_TRY(ensure_tmp(s, 16));
    CUDA_TRY(cudaMemsetAsync(s->tmp, 0, 16, c->stream));
    unsigned long long *dm = reinterpret_cast<unsigned long long *>(s->tmp);
    unsigned *dd = reinterpret_cast<unsigned *>(dm + 1);
    LAUNCH_1D(c, k_count_movers, s->nown, s->nlo, s->nown, A.s[0], A.s[1], A.s[2], s->g.M, s->gcell[s->cur], A.r[0], A.r[1], A.r[2],
              A.rb[0], A.rb[1], A.rb[2], dm, dd);
    unsigned long long h[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(h, s->tmp, 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    EMDEE_TRY(check_launch("k_count_movers"));
    if (movers) *movers = (int64_t)h[0];
    if (h[0] != 0) return do_bin(s, s->ndiv);
    float d2;
    const unsigned bits = (unsigned)(h[1] & 0xffffffffu);
    memcpy(&d2, &bits, 4);
    if ((double)d2 > 0.25 * s->skin * s->skin) s->list_valid = false;      // a list built with a skin holds only while atoms stay within skin/2
    s->binned = true;
    s->last_bitmask = 0;
    return EMDEE_OK;
}

extern "C" int emdee_get_cells_per_dimension(emdee_system *s, int32_t *M)
{
    SYS_ENTER(s, "emdee_get_cells_per_dimension");
    if (!s->binned || !M) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_cells_per_dimension: call emdee_bin first");
    *M = s->g.M;
    return EMDEE_OK;
}

extern "C" int emdee_get_cell_index(emdee_system *s, int32_t *index)
{
    SYS_ENTER(s, "emdee_get_cell_index");
    if (!s->binned || !index) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_cell_index: call emdee_bin first");
    EMDEE_TRY(ensure_tmp(s, sizeof(int32_t) * s->N));
    CUDA_TRY(cudaMemsetAsync(s->tmp, 0, sizeof(int32_t) * s->N, c->stream));
    // the reference's index is 1-based (src/cells.jl:85)
    LAUNCH_1D(c, k_get1<int32_t>, s->nown, s->nlo, s->nown, A.id, s->gcell[s->cur], (int32_t *)s->tmp, (int32_t)1);
    CUDA_TRY(cudaMemcpyAsync(index, s->tmp, sizeof(int32_t) * s->N, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return check_launch("get_cell_index");
}

extern "C" int emdee_get_cell_population(emdee_system *s, int32_t *pop)
{
    SYS_ENTER(s, "emdee_get_cell_population");
    if (!s->binned || !pop) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_cell_population: call emdee_bin first");
    const int64_t plane = (int64_t)s->g.M * s->g.M;
    CUDA_TRY(cudaMemcpyAsync(pop, s->count + plane * s->g.zhome0, sizeof(int32_t) * plane * s->g.nzhome,
                             cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return EMDEE_OK;
}

extern "C" int emdee_get_cell_order(emdee_system *s, int32_t *perm, int32_t *cell_start)
{
    SYS_ENTER(s, "emdee_get_cell_order");
    if (!s->binned) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_cell_order: call emdee_bin first");
    if (perm) CUDA_TRY(cudaMemcpyAsync(perm, A.id + s->nlo, sizeof(int32_t) * s->nown, cudaMemcpyDeviceToHost, c->stream));
    if (cell_start) {
        const int64_t plane = (int64_t)s->g.M * s->g.M;
        CUDA_TRY(cudaMemcpyAsync(cell_start, s->cell_start + plane * s->g.zhome0, sizeof(int32_t) * (plane * s->g.nzhome + 1),
                                 cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return EMDEE_OK;
}

extern "C" int emdee_get_local_count(emdee_system *s, int64_t *nlocal, int64_t *nghost)
{
    SYS_ENTER(s, "emdee_get_local_count");
    if (nlocal) *nlocal = s->nown;
    if (nghost) *nghost = s->nlo + s->nhi;
    return EMDEE_OK;
}
extern "C" int emdee_get_local_ids(emdee_system *s, int32_t *ids)
{
    SYS_ENTER(s, "emdee_get_local_ids");
    if (!ids) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_get_local_ids: null array");
    CUDA_TRY(cudaMemcpyAsync(ids, A.id + s->nlo, sizeof(int32_t) * s->nown, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return EMDEE_OK;
}

// ------------------------------------------------------------------------------------------------
// force evaluation
// ------------------------------------------------------------------------------------------------
template <bool F, bool EW, bool EXCL, bool AUDIT, bool TYPED>
static int launch_cells_t(emdee_system *s, const CellArgs &a, int nblocks)
{
    if (nblocks <= 0) return EMDEE_OK;
    auto kern = k_force_cells<F, EW, EXCL, AUDIT, TYPED>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->fc_smem));
    kern<<<nblocks, s->fc_block, s->fc_smem, s->ctx->stream>>>(a);
    s->ctx->launches++;
    return check_launch("k_force_cells");
}
template <bool EXCL>
static int launch_build_t(emdee_system *s, const CellArgs &a, int nblocks)
{
    if (nblocks <= 0) return EMDEE_OK;
    auto kern = s->build_n3 ? k_list_build_plain<EXCL, true> : k_list_build_plain<EXCL, false>;
    if (a.split || a.compact) kern = k_list_build<EXCL, false, true>;      // (dense cells; never together with the half list)
    const size_t smem = lb_smem_bytes(a.cap, s->fc_ncs, LB_MAX_BLOCK, EXCL);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nblocks, LB_MAX_BLOCK, smem, s->ctx->stream>>>(a);
    s->ctx->launches++;
    return check_launch("k_list_build");
}
template <bool MULTI, bool COUNT, int ILP>
static int launch_list_i(emdee_system *s, const CellArgs &a, int nblocks)
{
    auto kern = k_force_list<MULTI, COUNT, ILP>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->fl_smem));
    kern<<<nblocks, s->fl_block, s->fl_smem, s->ctx->stream>>>(a);
    s->ctx->launches++;
    return check_launch("k_force_list");
}
template <bool MULTI, bool COUNT, bool EW>
static int launch_list_p(emdee_system *s, const CellArgs &a, int nblocks, bool store_f)
{
    auto kern = s->fl_fuse ? k_force_list_p<MULTI, COUNT, 2, EW, true> : k_force_list_p<MULTI, COUNT, 2, EW, false>;
    // dense cells (shallow stacks and / or split lists): the variants that read both from the launch arguments
    const bool dense = s->fl_split || s->fl_qcap != FL_QCAP;
    if (dense) kern = k_force_list_p<MULTI, COUNT, 2, EW, true, false, false, false, false, 0, true>;
    const bool n3 = !COUNT && !EW && s->vv_mode != 0 && s->list_n3;
    bool tma = s->fl_tma && a.seg != nullptr && !n3;      // the stepping variants and the single-point F/E/W variant have a TMA form
    if (!COUNT && !EW && s->vv_mode != 0) {
        if (n3) kern = s->p2p_launch ? k_force_list_p<MULTI, false, 2, false, true, true, true, true> : k_force_list_p<MULTI, false, 2, false, true, true, false, true>;
        else if (tma) kern = s->p2p_launch ? k_force_list_p<MULTI, false, 2, false, true, true, true, false, true> : k_force_list_p<MULTI, false, 2, false, true, true, false, false, true>;
        else if (dense) kern = k_force_list_p<MULTI, false, 2, false, true, true, false, false, false, 0, true>;
        else if (s->lm == 1 && !s->p2p_launch) kern = k_force_list_p<MULTI, false, 2, false, true, true, false, false, false, 1>;
        else if (s->lm == 2 && !s->p2p_launch) kern = k_force_list_p<MULTI, false, 2, false, true, true, false, false, false, 2>;
        else kern = s->p2p_launch ? k_force_list_p<MULTI, false, 2, false, true, true, true> : k_force_list_p<MULTI, false, 2, false, true, true, false>;
        s->counters[(n3 || tma || s->p2p_launch) ? 1 : 1 + s->lm]++;
    } else if (dense)
        tma = false;
    else if (COUNT && !EW && s->lm == 2)       // the audit counts through the inner list when the stepping kernel would replay it
        { kern = k_force_list_p<MULTI, true, 2, false, true, false, false, false, false, 2>; tma = false; }
    else if (tma && !COUNT && EW && s->fl_fuse)
        kern = k_force_list_p<MULTI, false, 2, true, true, false, false, false, true>;
    else
        tma = false;
    const size_t smem = flp_smem_bytes(a.cap, s->fc_ncs, std::max(s->ntypes, 1), 2, n3 ? s->fc_gmax : 0, s->fl_qcap) +
                        (tma ? flp_tma_bytes(s->segcap, s->rawlen) : 0);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // while a halo exchange is in flight the persistent blocks leave a few SMs to NCCL's kernel (each block holds all
    // registers of its SM, so NCCL could not start before the first block retires otherwise)
    const int sms = std::max(1, s->ctx->sm_count - s->reserve_sms);
    CUDA_TRY(cudaMemsetAsync(s->brick_counter, 0, 4 * sizeof(int), s->ctx->stream));
    CellArgs ac = a;
    ac.brick_counter = s->brick_counter;
    ac.timing = s->digest + 8;
    ac.vv_mode = 0;
    if (!COUNT && !EW && s->vv_mode != 0) {
        AtomArrays &A = s->A[s->cur];
        ac.vv_mode = s->vv_mode; ac.vv_dt = s->vv_dt; ac.vv_half_skin2 = 0.25 * s->skin * s->skin;
        for (int k = 0; k < 3; k++) { ac.vv_v[k] = A.v[k]; ac.vv_r[k] = A.r[k]; ac.vv_snew[k] = s->s_alt[k]; ac.vv_rb[k] = A.rb[k]; }
        ac.vv_mass = A.mass;
        ac.vv_maxd2 = s->vv_track ? s->maxd2 : nullptr;
        ac.vv_maxstep = s->vv_track ? s->maxd2 + 2 : nullptr;
        ac.vv_check_skin = s->vv_check_skin;
        if (s->p2p_launch) {
            emdee_ctx *c = s->ctx;
            const int layer = s->g.nbx * s->g.nby;
            ac.p2p = 1;
            ac.p2p_lo_layers = s->brick_lo_end; ac.p2p_hi_layer0 = s->hi_layer0;
            ac.p2p_nlo = s->brick_lo_end * layer; ac.p2p_nhi = (s->g.nbz - s->hi_layer0) * layer;
            ac.lo_send_a = (int)s->lo_send_a; ac.lo_send_n = (int)s->lo_send_n; ac.hi_send_a = (int)s->hi_send_a; ac.hi_send_n = (int)s->hi_send_n;
            for (int k = 0; k < 3; k++) {      // the neighbours rotate their position buffers as I do: same offset inside the pool
                const ptrdiff_t off = s->s_alt[k] - s->spool;
                ac.peer_lo[k] = reinterpret_cast<double *>(s->peer_pool[0]) + off;
                ac.peer_hi[k] = reinterpret_cast<double *>(s->peer_pool[1]) + off;
            }
            ac.peer_info = s->peerinfo + 4;
            ac.flag_lo_peer = reinterpret_cast<unsigned long long *>(s->peer_pool[0]) + 1;   // the lower neighbour's "from above" word
            ac.flag_hi_peer = reinterpret_cast<unsigned long long *>(s->peer_pool[1]) + 0;   // the upper neighbour's "from below" word
            ac.flag_from_lo = reinterpret_cast<unsigned long long *>(s->spool) + 0;
            ac.flag_from_hi = reinterpret_cast<unsigned long long *>(s->spool) + 1;
            ac.wait_epoch = s->p2p_wait; ac.publish_epoch = s->p2p_publish;
            ac.p2p_done = s->brick_counter + 1;
            (void)c;
        }
    }
    kern<<<std::min(nblocks, sms), FLP_THREADS, smem, s->ctx->stream>>>(ac, nblocks, store_f ? 1 : 0);
    s->ctx->launches++;
    return check_launch("k_force_list_p");
}
template <bool MULTI, bool COUNT>
static int launch_list_t(emdee_system *s, const CellArgs &a, int nblocks, bool F, bool EW)
{
    if (nblocks <= 0) return EMDEE_OK;
    if (s->fl_persistent) {     // one resident block per SM walks over the bricks (force_list_p.cuh)
        if (EW && !COUNT) return launch_list_p<MULTI, false, true>(s, a, nblocks, F);
        return launch_list_p<MULTI, COUNT, false>(s, a, nblocks, !COUNT);     // the counting pass leaves the forces alone
    }
    if (EW) EMDEE_FAIL(EMDEE_ERR_STATE, "launch_list: energies and virials need the persistent list kernel");
    // blocks of up to 192 threads run the 8-chain variant (more registers per thread), larger ones the 4-chain variant
    return s->fl_block <= 192 && s->fl_ilp8 ? launch_list_i<MULTI, COUNT, 8>(s, a, nblocks) : launch_list_i<MULTI, COUNT, 4>(s, a, nblocks);
}
template <bool TYPED>
static int launch_cells_ty(emdee_system *s, const CellArgs &a, int nb, bool F, bool EW, bool EXCL, bool AUDIT)
{
    if (AUDIT) return EXCL ? launch_cells_t<true, true, true, true, TYPED>(s, a, nb) : launch_cells_t<true, true, false, true, TYPED>(s, a, nb);
    if (EXCL) {
        if (F && EW) return launch_cells_t<true, true, true, false, TYPED>(s, a, nb);
        if (F) return launch_cells_t<true, false, true, false, TYPED>(s, a, nb);
        return launch_cells_t<false, true, true, false, TYPED>(s, a, nb);
    }
    if (F && EW) return launch_cells_t<true, true, false, false, TYPED>(s, a, nb);
    if (F) return launch_cells_t<true, false, false, false, TYPED>(s, a, nb);
    return launch_cells_t<false, true, false, false, TYPED>(s, a, nb);
}
// mode 0: window scan (k_force_cells); 1: pair-list build (k_list_build, no forces); 2: pair-list walk (k_force_list);
// 3: the same, counting the pairs inside the cutoff
static int launch_cells(emdee_system *s, const CellArgs &a, int nb, bool F, bool EW, bool EXCL, bool AUDIT, int mode)
{
    if (mode == 1) return EXCL ? launch_build_t<true>(s, a, nb) : launch_build_t<false>(s, a, nb);
    if (mode == 2) return s->ntypes > 1 ? launch_list_t<true, false>(s, a, nb, F, EW) : launch_list_t<false, false>(s, a, nb, F, EW);
    if (mode == 3) return s->ntypes > 1 ? launch_list_t<true, true>(s, a, nb, true, false) : launch_list_t<false, true>(s, a, nb, true, false);
    return s->ntypes > 0 ? launch_cells_ty<true>(s, a, nb, F, EW, EXCL, AUDIT) : launch_cells_ty<false>(s, a, nb, F, EW, EXCL, AUDIT);
}

// FP16 pre-culls (k_force_list*, k_list_build): a pair with r <= rcut must pass r2_fp16 <= threshold.
//   stored coordinates: |c_k| <= h_k (half extent of the staged box + skin/2), rounded to FP16 (through FP32):
//     half an ulp16(h_k) per atom; the separation rounds once more: ulp16(rcut)/2   ->  delta_k
//   r2 of the stored separations <= rcut^2 + 2 rcut |delta| + |delta|^2            (Cauchy-Schwarz)
//   at most four FP16 roundings in the squares and sums: 4.2 u (rcut + |delta|)^2, u = 2^-11
// (5 % head-room on the error terms; the device rounds the threshold up to FP16.)
static float fp16_threshold(const double hext[3], double rcut)
{
    auto ulp16 = [](double v) { return std::ldexp(1.0, std::max(-14, (int)std::floor(std::log2(v))) - 10); };
    const double u = std::ldexp(1.0, -11);
    double d2 = 0;
    for (int k = 0; k < 3; k++) {
        const double dk = 1.001 * ulp16(hext[k]) + 0.5 * ulp16(rcut);
        d2 += dk * dk;
    }
    const double dn = std::sqrt(d2);
    const double err = 2.0 * rcut * dn + d2 + 4.2 * u * (rcut + dn) * (rcut + dn);
    return (float)((rcut * rcut + 1.05 * err) * (1.0 + u));
}
// host-only helper behind the CPU test of that bound (no GPU involved)
extern "C" int emdee_fp16_threshold(const double half_extent[3], double rcut, float *threshold)
{
    if (!half_extent || !threshold || !(rcut > 0)) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_fp16_threshold: bad arguments");
    *threshold = fp16_threshold(half_extent, rcut);
    return EMDEE_OK;
}

// Newton's third law inside the brick is usable when the persistent kernel runs, the home index fits the recipe's 11 spare
// bits, and the accumulators fit next to the two staging buffers
static bool n3_usable(const emdee_system *s)
{
    return s->want_n3 && s->fl_persistent && !s->fl_compact && !s->fl_split && s->fl_qcap == FL_QCAP && s->fc_gmax * 32 < 2047 &&
           flp_smem_bytes(s->fc_cap, s->fc_ncs, std::max(s->ntypes, 1), 2, s->fc_gmax) <= s->ctx->smem_optin;
}

// halo: when true (slab decomposition, inside the step loop) the ghost positions are refreshed on the
// communication stream while the interior brick layers compute; the boundary layers wait for them.
static int run_cells(emdee_system *s, int bitmask, bool audit, int32_t *pairs, int64_t pair_cap, bool halo = false, int mode = 0)
{
    emdee_ctx *c = s->ctx;
    if (mode >= 2 && s->list_valid) {
        // the fused velocity-Verlet launch walks a half list (N3), every other list kernel the full one: rebuild on a change
        const bool need = mode == 2 && s->vv_mode != 0 && n3_usable(s);
        if (s->list_n3 != need) {
            const int vm = s->vv_mode;
            const bool pl = s->p2p_launch;
            s->vv_mode = 0; s->p2p_launch = false; s->build_n3 = need;
            const int rc_ = run_cells(s, EMDEE_FORCES, false, nullptr, 0, false, 1);
            s->vv_mode = vm; s->p2p_launch = pl;
            EMDEE_TRY(rc_);
        }
    }
    if ((s->ntypes > 0) != s->fc_typed) EMDEE_TRY(choose_bricks(s));   // LJ classes changed since binning: re-size shared memory
    AtomArrays &A = s->A[s->cur];
    CellArgs a = {};
    a.g = s->g;
    a.cell_start = s->cell_start;
    a.sx = A.s[0]; a.sy = A.s[1]; a.sz = A.s[2];
    a.hs = A.hs; a.ts = A.ts;
    a.id = A.id; a.type = A.type; a.xbase = A.xbase; a.xmask = A.xmask;
    a.ljtab = s->ljtab; a.ntypes = s->ntypes;
    a.fx = s->f[0]; a.fy = s->f[1]; a.fz = s->f[2];
    a.en = s->en; a.vir = s->vir;
    if (audit) {       // side-effect free: the audit's own forces / energies / virials go to scratch arrays
        EMDEE_TRY(ensure_audit_scratch(s));
        a.fx = s->aud[0]; a.fy = s->aud[1]; a.fz = s->aud[2]; a.en = s->aud[3]; a.vir = s->aud[4];
    }
    a.digest = s->digest;
    a.pairs = pairs;
    a.pair_cap = pair_cap;
    a.pair_n = s->digest + 3;
    a.L = s->L;
    a.cell_edge = s->L / s->g.M;
    a.model = s->model;
    {   // conservative FP32 pre-cull threshold: rc2 plus a bound on the rounding error of
        // |c|^2 - 2 c.p + |p|^2 for coordinates within the staged region (relative to the brick centre)
        const double hx = 0.5 * (s->g.bx + 2 * s->g.R) * a.cell_edge + 0.5 * s->skin;
        const double hy = 0.5 * (s->g.by + 2 * s->g.R) * a.cell_edge + 0.5 * s->skin;
        const double hz = 0.5 * (s->g.bz + 2 * s->g.R) * a.cell_edge + 0.5 * s->skin;
        const double cmax2 = hx * hx + hy * hy + hz * hz;
        a.rc2f = (float)(s->model.rc2 * (1.0 + 1e-4) + 64.0 * 1.2e-7 * cmax2);
        const double rl = s->cutoff + s->skin;
        a.rl2f = (float)(rl * rl * (1.0 + 1e-4) + 64.0 * 1.2e-7 * cmax2);
    }
    // staged-atom capacity: the list kernels of a persistent configuration may stage a compacted brick (fl_cap < fc_cap)
    a.cap = (mode != 0 && s->fl_persistent) ? s->fl_cap : s->fc_cap;
    a.qcap = s->fl_qcap;
    a.compact = (mode != 0 && s->fl_persistent && s->fl_compact) ? 1 : 0;
    a.split = (mode != 0 && s->fl_persistent && s->fl_split) ? 1 : 0;
    {
        const double rl = s->cutoff + s->skin;
        a.keep2 = rl * rl * (1.0 + 1e-6);
    }
    a.ncs_max = s->fc_ncs;
    a.err = s->err;
    a.list8 = s->list8; a.list_n = s->list_n; a.gmax = s->fc_gmax; a.lcap8 = s->lcap8;
    {
        uint64_t bits;
        memcpy(&bits, &s->model.rc2, 8);
        a.rc2hi = (int)(bits >> 32);
        a.fast.id2 = s->model.id2;
        a.fast.nrs2id2 = -s->model.rs2 * s->model.id2;
        a.fast.c60id2 = 60.0 * s->model.id2;
        const double hext[3] = {0.5 * (s->g.bx + 2 * s->g.R) * a.cell_edge + 0.5 * s->skin,
                                0.5 * (s->g.by + 2 * s->g.R) * a.cell_edge + 0.5 * s->skin,
                                0.5 * (s->g.bz + 2 * s->g.R) * a.cell_edge + 0.5 * s->skin};
        auto thr16 = [&](double rcut) { return fp16_threshold(hext, rcut); };
        a.rc2h = thr16(s->cutoff);
        a.rl2h = thr16(s->cutoff + s->skin);
        a.rp2h = thr16(s->cutoff + std::min(s->skin2, s->skin));
    }
    if (mode != 0) {
        if (audit) EMDEE_FAIL(EMDEE_ERR_STATE, "run_cells: the pair-list kernels do not audit");
        if (!(s->grid_ok && s->use_list && s->ntypes > 0)) EMDEE_FAIL(EMDEE_ERR_STATE, "run_cells: pair list requested for a system that cannot use one");
        const int64_t slots = (int64_t)s->fc_nblocks * s->fc_gmax;
        if (!s->lcap8_forced) {
            // chunks per atom from the mean density: entries inside rc + skin (+ FP16 margin), 40 % head-room for
            // density fluctuations, at least 24 chunks; a denser neighbourhood than that raises the overflow error
            const double rl = s->cutoff + s->skin + 0.05;
            const double mean = (double)s->N / (s->L * s->L * s->L) * 4.18879 * rl * rl * rl;
            const int want = std::max(24, (int)std::ceil(1.4 * mean / 8.0) + 2);
            if (want > s->lcap8) { s->lcap8 = want; s->list_slots = 0; }      // re-allocate below
            a.lcap8 = s->lcap8;
        }
        if (slots > s->list_slots) {
            dev_free(s->list8); dev_free(s->list_n); dev_free(s->homeidx);
            s->list_slots = slots + slots / 4;
            EMDEE_TRY(dev_alloc(&s->homeidx, (size_t)s->list_slots * 32));
            if (getenv("EMDEE_DEBUG")) fprintf(stderr, "[emdee] pair list: %lld groups x %d chunks (%.2f GB)\n", (long long)s->list_slots, s->lcap8, (double)s->list_slots * s->lcap8 * 512 / 1e9);
            EMDEE_TRY(dev_alloc(&s->list8, (size_t)s->list_slots * s->lcap8 * 32));
            EMDEE_TRY(dev_alloc(&s->list_n, (size_t)s->list_slots * 32));
            a.list8 = s->list8; a.list_n = s->list_n;
        }
        const int64_t rneed = (int64_t)s->fc_nblocks * (a.cap + 1);
        if (rneed > s->recipe_cap) {
            dev_free(s->recipe);
            s->recipe_cap = rneed + rneed / 8;
            EMDEE_TRY(dev_alloc(&s->recipe, (size_t)s->recipe_cap));
        }
        if (s->fc_nblocks > s->hdr_cap) {
            dev_free(s->brickhdr);
            s->hdr_cap = s->fc_nblocks + s->fc_nblocks / 4;
            EMDEE_TRY(dev_alloc(&s->brickhdr, (size_t)2 * s->hdr_cap));
        }
        a.recipe = s->recipe; a.homeidx = s->homeidx; a.brickhdr = s->brickhdr; a.rcap = a.cap + 1;
        if ((mode == 2 || mode == 3) && s->lm != 0) {
            if (s->inner_slots < s->list_slots || s->inner_lcap8 != s->lcap8) {
                if (s->lm == 2) EMDEE_FAIL(EMDEE_ERR_STATE, "run_cells: no inner list to replay");
                dev_free(s->inner8); dev_free(s->inner_n);
                s->inner_slots = s->list_slots; s->inner_lcap8 = s->lcap8;
                EMDEE_TRY(dev_alloc(&s->inner8, (size_t)s->inner_slots * s->lcap8 * 32));
                EMDEE_TRY(dev_alloc(&s->inner_n, (size_t)s->inner_slots));
            }
            a.inner8 = s->inner8; a.inner_n = s->inner_n;
        }
        if (s->fl_tma) {
            const int64_t sneed = (int64_t)s->fc_nblocks * s->segcap;
            if (sneed > s->seg_cap_total) {
                dev_free(s->seg);
                s->seg_cap_total = sneed + sneed / 8;
                EMDEE_TRY(dev_alloc(&s->seg, (size_t)s->seg_cap_total));
            }
            a.seg = s->seg; a.segcap = s->segcap; a.raw_rows = s->raw_rows; a.rawlen = s->rawlen;
            if (mode == 1 && getenv("EMDEE_DEBUG"))
                fprintf(stderr, "[emdee] TMA staging: %d segments per brick, raw ring of 2 x %d rows x %d atoms (longest row %d)\n", s->segcap,
                        s->raw_rows, s->rawlen / std::max(1, s->raw_rows), s->fc_rowmax);
        }
        // lanes of a group without a home atom must read "no entries"
        if (mode == 1 && (s->range_count == 0 || s->range_first == 0))
            CUDA_TRY(cudaMemsetAsync(s->list_n, 0, (size_t)slots * 32 * sizeof(uint16_t), c->stream));
    }
    const bool F = (bitmask & EMDEE_FORCES) != 0;
    const bool EW = (bitmask & (EMDEE_ENERGIES | EMDEE_VIRIALS)) != 0;
    if (audit) CUDA_TRY(cudaMemsetAsync(s->digest, 0, 4 * sizeof(unsigned long long), c->stream));
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (s->profiling && s->prof_used + 2 <= 16384) {
        if (s->prof_events.size() < s->prof_used + 2) {
            s->prof_events.resize(s->prof_used + 2);
            CUDA_TRY(cudaEventCreate(&s->prof_events[s->prof_used]));
            CUDA_TRY(cudaEventCreate(&s->prof_events[s->prof_used + 1]));
        }
        pe0 = s->prof_events[s->prof_used];
        pe1 = s->prof_events[s->prof_used + 1];
        if (s->prof_mode.size() < s->prof_used / 2 + 1) s->prof_mode.resize(s->prof_used / 2 + 1);
        s->prof_mode[s->prof_used / 2] = mode;
        s->prof_used += 2;
        CUDA_TRY(cudaEventRecord(pe0, c->stream));
    }
    const int layer = s->g.nbx * s->g.nby;
    // launch order: interior layers first, then (one launch over two ranges) the layers that read ghost planes
    int ranges[2][3] = {{0, s->fc_nblocks, 0}, {0, 0, 0}};       // {first brick, bricks of the first range, bricks of the second range}
    int second_first = 0;
    if (s->range_count > 0) { ranges[0][0] = s->range_first * layer; ranges[0][1] = s->range_count * layer; }
    if (halo && c->nranks > 1) {
        CUDA_TRY(cudaEventRecord(c->ev_compute, c->stream));
        CUDA_TRY(cudaStreamWaitEvent(c->comm_stream, c->ev_compute, 0));
        EMDEE_TRY(slab_halo_positions(s, c->comm_stream));
        CUDA_TRY(cudaEventRecord(c->ev_comm, c->comm_stream));
        ranges[0][0] = s->brick_lo_end * layer; ranges[0][1] = (s->brick_hi_begin - s->brick_lo_end) * layer;
        ranges[1][0] = 0; ranges[1][1] = s->brick_lo_end * layer;
        second_first = s->brick_hi_begin * layer; ranges[1][2] = s->fc_nblocks - s->brick_hi_begin * layer;
    }
    a.block_split2 = 0x7fffffff;
    a.block_first3 = 0;
    if (s->p2p_launch) {
        // ONE launch per step: the bricks whose halo reaches ghost planes come first (their atoms' new positions are what the
        // neighbours wait for), then the interior; the kernel itself waits for the neighbours' flags where it has to
        a.block_first = 0; a.block_split = s->brick_lo_end * layer;
        a.block_first2 = s->brick_hi_begin * layer; a.block_split2 = s->fc_nblocks - s->brick_hi_begin * layer;
        a.block_first3 = s->brick_lo_end * layer;
        EMDEE_TRY(launch_cells(s, a, s->fc_nblocks, F, EW, s->has_excl, audit, mode));
        if (pe1) CUDA_TRY(cudaEventRecord(pe1, c->stream));
        return EMDEE_OK;
    }
    for (int k = 0; k < 2; k++) {
        if (k == 1 && halo && c->nranks > 1) {
            cudaEvent_t m0 = nullptr, m1 = nullptr;
            if (pe0 && getenv("EMDEE_DEBUG")) {
                CUDA_TRY(cudaEventCreate(&m0)); CUDA_TRY(cudaEventCreate(&m1));
                s->prof_mid.push_back(m0); s->prof_mid.push_back(m1);
                s->prof_mid_of.push_back(s->prof_used - 2);
                CUDA_TRY(cudaEventRecord(m0, c->stream));
            }
            CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
            if (m1) CUDA_TRY(cudaEventRecord(m1, c->stream));
        }
        a.block_first = ranges[k][0];
        a.block_split = ranges[k][1];
        a.block_first2 = second_first;
        s->reserve_sms = (k == 0 && halo && c->nranks > 1) ? s->nccl_sms : 0;
        EMDEE_TRY(launch_cells(s, a, ranges[k][1] + ranges[k][2], F, EW, s->has_excl, audit, mode));
        s->reserve_sms = 0;
        if (getenv("EMDEE_DEBUG_SYNC")) {
            fprintf(stderr, "[emdee] force launch mode %d range %d (%d blocks) issued\n", mode, k, ranges[k][1] + ranges[k][2]);
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            fprintf(stderr, "[emdee] force launch mode %d range %d done\n", mode, k);
        }
    }
    if (pe1) CUDA_TRY(cudaEventRecord(pe1, c->stream));
    if (mode == 1) { s->list_n3 = s->build_n3; s->build_n3 = false; s->list_gen++; }
    return EMDEE_OK;
}

template <bool CULL, bool EXCL>
static int launch_tiles_t(emdee_system *s, const TileArgs &a, int bitmask)
{
    const unsigned grid = (unsigned)ceil_div64(a.ntiles, 4);
    const bool F = bitmask & EMDEE_FORCES, E = bitmask & EMDEE_ENERGIES, W = bitmask & EMDEE_VIRIALS;
#define TILE_CASE(f, e, w) \
    if (F == f && E == e && W == w) k_force_tiles<f, e, w, CULL, EXCL><<<grid, 128, 0, s->ctx->stream>>>(a);
    TILE_CASE(true, true, true) TILE_CASE(true, true, false) TILE_CASE(true, false, true) TILE_CASE(true, false, false)
    TILE_CASE(false, true, true) TILE_CASE(false, true, false) TILE_CASE(false, false, true)
#undef TILE_CASE
    s->ctx->launches++;
    return check_launch("k_force_tiles");
}

static int run_tiles(emdee_system *s, int bitmask, bool cull, int32_t *pairs = nullptr, int64_t pair_cap = 0, bool audit = false)
{
    emdee_ctx *c = s->ctx;
    AtomArrays &A = s->A[s->cur];
    if (s->tiles_default && s->ntiles == 0) EMDEE_TRY(default_tiles(s));
    const int64_t n = s->nown;
    double *of[3] = {s->f[0], s->f[1], s->f[2]}, *oe = s->en, *ow = s->vir;
    if (audit) {       // pair digest / pair set: outputs to scratch, the last compute's results stay
        EMDEE_TRY(ensure_audit_scratch(s));
        for (int k = 0; k < 3; k++) of[k] = s->aud[k];
        oe = s->aud[3]; ow = s->aud[4];
    }
    // compute_nonbonded! zeroes the selected outputs, then accumulates with atomics (src/nonbonded.jl:112-114)
    if (bitmask & EMDEE_FORCES) for (int k = 0; k < 3; k++) CUDA_TRY(cudaMemsetAsync(of[k], 0, sizeof(double) * n, c->stream));
    if (bitmask & EMDEE_ENERGIES) CUDA_TRY(cudaMemsetAsync(oe, 0, sizeof(double) * n, c->stream));
    if (bitmask & EMDEE_VIRIALS) CUDA_TRY(cudaMemsetAsync(ow, 0, sizeof(double) * n, c->stream));
    CUDA_TRY(cudaMemsetAsync(s->digest, 0, 4 * sizeof(unsigned long long), c->stream));
    TileArgs a;
    a.tiles = s->tiles; a.ntiles = s->ntiles; a.N = s->N;
    a.slot_of_id = s->slot_of_id;
    a.sx = A.s[0]; a.sy = A.s[1]; a.sz = A.s[2]; a.hs = A.hs; a.ts = A.ts;
    a.id = A.id; a.xbase = A.xbase; a.xmask = A.xmask;
    a.fx = of[0]; a.fy = of[1]; a.fz = of[2]; a.en = oe; a.vir = ow;
    a.digest = s->digest;
    a.pairs = pairs;
    a.pair_cap = pair_cap;
    a.L = s->L;
    a.model = s->model;
    if (cull) return s->has_excl ? launch_tiles_t<true, true>(s, a, bitmask) : launch_tiles_t<true, false>(s, a, bitmask);
    return s->has_excl ? launch_tiles_t<false, true>(s, a, bitmask) : launch_tiles_t<false, false>(s, a, bitmask);
}

static int check_device_flag(emdee_system *s, const char *where)
{
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, s->err, sizeof(int), cudaMemcpyDeviceToHost, s->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(s->ctx->stream));
    if (flag) {
        CUDA_TRY(cudaMemsetAsync(s->err, 0, sizeof(int), s->ctx->stream));
        if (flag == 3) EMDEE_FAIL(EMDEE_ERR_SKIN, "%s: an atom moved more than skin/2 since the last binning; re-bin more often or raise the skin", where);
        if (flag == 2) EMDEE_FAIL(EMDEE_ERR_CAPACITY, "%s: a brick overflowed its shared-memory staging area", where);
        if (flag == 5) EMDEE_FAIL(EMDEE_ERR_CAPACITY, "%s: the pair list overflowed its capacity (set EMDEE_LIST_CHUNKS higher or EMDEE_LIST=0)", where);
        if (flag == 4) EMDEE_FAIL(EMDEE_ERR_CAPACITY, "%s: migration list overflow", where);
        if (flag == 8) EMDEE_FAIL(EMDEE_ERR_STATE, "%s: an internal bounds check of the list kernels failed (library built with -DEMDEE_CHECKS=1)", where);
        if (flag == 9) EMDEE_FAIL(EMDEE_ERR_CUDA, "%s: a bulk copy of the staging path never completed", where);
        if (flag == 6) EMDEE_FAIL(EMDEE_ERR_NCCL, "%s: a neighbouring rank never published its boundary atoms (peer-mapped halo timed out after ~2 s)", where);
        if (flag == 7) EMDEE_FAIL(EMDEE_ERR_INVALID, "%s: a position window did not cover every atom this rank owns", where);
        EMDEE_FAIL(EMDEE_ERR_STATE, "%s: an atom left the slab's cell range", where);
    }
    return EMDEE_OK;
}

extern "C" int emdee_compute_nonbonded(emdee_system *s, int mode, int bitmask)
{
    SYS_ENTER(s, "emdee_compute_nonbonded");
    if (!s->has_model || !s->has_atoms || !s->has_pos)
        EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_compute_nonbonded: set model, LJ atoms and positions first");
    if ((bitmask & 7) == 0 || (bitmask & ~7)) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_compute_nonbonded: bitmask %d must combine FORCES|ENERGIES|VIRIALS", bitmask);
    if (mode == EMDEE_ALLPAIRS_REFERENCE) {
        if (c->nranks > 1) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_compute_nonbonded: ALLPAIRS_REFERENCE is single-GPU (the reference's O(N^2) mode is for small N)");
        EMDEE_TRY(run_tiles(s, bitmask, false));
    } else if (mode == EMDEE_CUTOFF) {
        if (!s->binned) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_compute_nonbonded: EMDEE_CUTOFF needs emdee_bin after the last emdee_set_positions");
        if (s->grid_ok) {
            if (single_point_list(s)) {
                // list kernels: build at the current binning if there is no valid list (k_list_build), then one walk
                // (k_force_list_p); a later evaluation within the skin re-uses the list
                if (!s->list_valid) {
                    EMDEE_TRY(run_cells(s, EMDEE_FORCES, false, nullptr, 0, false, 1));
                    s->list_valid = true;
                }
                EMDEE_TRY(run_cells(s, bitmask, false, nullptr, 0, false, 2));
            } else
                EMDEE_TRY(run_cells(s, bitmask, false, nullptr, 0));
        } else {
            if (!s->tiles_default) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_compute_nonbonded: box too small for a cell grid and a custom tile list is set");
            EMDEE_TRY(run_tiles(s, bitmask, true));
        }
        EMDEE_TRY(apply_pairs14(s, bitmask));
    } else
        EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_compute_nonbonded: unknown mode %d", mode);
    s->last_mode = mode;
    s->last_bitmask = bitmask;
    s->forces_valid = (bitmask & EMDEE_FORCES) != 0;
    return EMDEE_OK;
}

static int get3(emdee_system *s, double *const src[3], double *out, const char *what)
{
    emdee_ctx *c = s->ctx;
    AtomArrays &A = s->A[s->cur];
    if (!out) EMDEE_FAIL(EMDEE_ERR_INVALID, "%s: null array", what);
    EMDEE_TRY(ensure_tmp(s, sizeof(double) * 3 * s->N));
    if (s->nown != s->N) CUDA_TRY(cudaMemsetAsync(s->tmp, 0, sizeof(double) * 3 * s->N, c->stream));
    LAUNCH_1D(c, k_get3, s->nown, s->nlo, s->nown, A.id, src[0], src[1], src[2], s->tmp);
    CUDA_TRY(cudaMemcpyAsync(out, s->tmp, sizeof(double) * 3 * s->N, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return check_launch(what);
}
static int get1(emdee_system *s, const double *src, double *out, const char *what)
{
    emdee_ctx *c = s->ctx;
    AtomArrays &A = s->A[s->cur];
    if (!out) EMDEE_FAIL(EMDEE_ERR_INVALID, "%s: null array", what);
    EMDEE_TRY(ensure_tmp(s, sizeof(double) * s->N));
    if (s->nown != s->N) CUDA_TRY(cudaMemsetAsync(s->tmp, 0, sizeof(double) * s->N, c->stream));
    LAUNCH_1D(c, k_get1<double>, s->nown, s->nlo, s->nown, A.id, src, s->tmp, 0.0);
    CUDA_TRY(cudaMemcpyAsync(out, s->tmp, sizeof(double) * s->N, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return check_launch(what);
}

extern "C" int emdee_get_positions(emdee_system *s, double *out)
{
    SYS_ENTER(s, "emdee_get_positions");
    if (!s->has_pos) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_positions: positions were never set");
    EMDEE_TRY(check_device_flag(s, "emdee_get_positions"));      // a skin violation during emdee_vv_step means pairs were missed
    return get3(s, A.r, out, "emdee_get_positions");
}
extern "C" int emdee_get_velocities(emdee_system *s, double *out)
{
    SYS_ENTER(s, "emdee_get_velocities");
    EMDEE_TRY(check_device_flag(s, "emdee_get_velocities"));
    return get3(s, A.v, out, "emdee_get_velocities");
}
extern "C" int emdee_get_forces(emdee_system *s, double *out)
{
    SYS_ENTER(s, "emdee_get_forces");
    if (!(s->last_bitmask & EMDEE_FORCES)) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_forces: the last compute did not select FORCES");
    EMDEE_TRY(check_device_flag(s, "emdee_get_forces"));
    return get3(s, s->f, out, "emdee_get_forces");
}
extern "C" int emdee_get_energies(emdee_system *s, double *out)
{
    SYS_ENTER(s, "emdee_get_energies");
    if (!(s->last_bitmask & EMDEE_ENERGIES)) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_energies: the last compute did not select ENERGIES");
    EMDEE_TRY(check_device_flag(s, "emdee_get_energies"));
    return get1(s, s->en, out, "emdee_get_energies");
}
extern "C" int emdee_get_virials(emdee_system *s, double *out)
{
    SYS_ENTER(s, "emdee_get_virials");
    if (!(s->last_bitmask & EMDEE_VIRIALS)) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_virials: the last compute did not select VIRIALS");
    EMDEE_TRY(check_device_flag(s, "emdee_get_virials"));
    return get1(s, s->vir, out, "emdee_get_virials");
}

// compute_nonbonded!(forces, energies, virials, ...) with HOST output arrays (src/nonbonded.jl:122-155): the evaluation of
// emdee_compute_nonbonded followed by the getters of the selected outputs, as one call.  On one GPU with a cell grid and the
// list kernels the box is evaluated in chunks of brick layers (z planes): list build and force kernel of chunk k+1 run while the
// rows of chunk k travel to the host on a second stream.  The host arrays are id-ordered, a chunk's atoms are whatever the
// binning put into its planes: their rows are found as occupied id buckets (runs of buckets = one copy each).  Where ids run
// with z (lattices, most structure files) a chunk is one or two runs; if the buckets of all chunks add up to much more than N
// rows (ids uncorrelated with position) the call falls back to one evaluation and three whole-array copies.
#define INTO_BUCKETS 2048
extern "C" int emdee_compute_nonbonded_into(emdee_system *s, int mode, int bitmask, double *forces, double *energies, double *virials)
{
    SYS_ENTER(s, "emdee_compute_nonbonded_into");
    if ((bitmask & 7) == 0 || (bitmask & ~7)) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_compute_nonbonded_into: bitmask %d must combine FORCES|ENERGIES|VIRIALS", bitmask);
    if (((bitmask & EMDEE_FORCES) && !forces) || ((bitmask & EMDEE_ENERGIES) && !energies) || ((bitmask & EMDEE_VIRIALS) && !virials))
        EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_compute_nonbonded_into: null array for a selected output");
    const GridDesc &g = s->g;
    const char *pe = getenv("EMDEE_PIPE");
    int want_chunks = pe ? atoi(pe) : 8;
    const bool fast = want_chunks > 1 && mode == EMDEE_CUTOFF && c->nranks == 1 && s->has_model && s->has_atoms && s->has_pos && s->binned &&
                      s->grid_ok && single_point_list(s) && (s->n14 == 0 || s->scale14 == 1.0) && g.zwrap && g.nbz >= 4 &&
                      (s->ntypes > 0) == s->fc_typed && s->N >= 65536;
    std::vector<std::pair<int64_t, int64_t>> runs[CHUNKS_MAX];      // per chunk: {first row, rows}
    ChunkPlan plan = {};
    int layer_bound[CHUNKS_MAX + 1];
    bool piped = fast;
    if (piped) {
        const int nchunks = std::min(std::min(want_chunks, CHUNKS_MAX), g.nbz / 2);
        plan.nchunks = nchunks;
        for (int k = 0; k <= nchunks; k++) {
            layer_bound[k] = (int)((int64_t)g.nbz * k / nchunks);
            plan.cell_bound[k] = std::min(layer_bound[k] * g.bz, g.M) * g.M * g.M;
        }
        // rows of every chunk as runs of occupied id buckets (one small kernel, one read-back)
        EMDEE_TRY(ensure_tmp(s, sizeof(double) * 5 * s->N + (size_t)CHUNKS_MAX * INTO_BUCKETS));
        unsigned char *occ_d = reinterpret_cast<unsigned char *>(s->tmp + 5 * s->N);
        CUDA_TRY(cudaMemsetAsync(occ_d, 0, (size_t)nchunks * INTO_BUCKETS, c->stream));
        LAUNCH_1D(c, k_chunk_id_buckets, s->N, s->N, plan, s->cell_start, A.id, s->N, INTO_BUCKETS, occ_d);
        std::vector<unsigned char> occ((size_t)nchunks * INTO_BUCKETS);
        CUDA_TRY(cudaMemcpyAsync(occ.data(), occ_d, occ.size(), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        EMDEE_TRY(check_launch("k_chunk_id_buckets"));
        auto lo_of = [&](int64_t b) { return (b * s->N + INTO_BUCKETS - 1) / INTO_BUCKETS; };     // bucket b holds ids [lo_of(b), lo_of(b + 1))
        int64_t total = 0;
        for (int k = 0; k < nchunks; k++) {
            const unsigned char *o = occ.data() + (size_t)k * INTO_BUCKETS;
            for (int b = 0; b < INTO_BUCKETS;) {
                if (!o[b]) { b++; continue; }
                int e = b;
                while (e < INTO_BUCKETS && o[e]) e++;
                runs[k].push_back({lo_of(b), lo_of(e) - lo_of(b)});
                total += lo_of(e) - lo_of(b);
                b = e;
            }
        }
        if (total > s->N + s->N / 2) piped = false;      // ids do not run with z: every chunk would copy most of the arrays
    }
    if (!piped) {
        EMDEE_TRY(emdee_compute_nonbonded(s, mode, bitmask));
        if (bitmask & EMDEE_FORCES) EMDEE_TRY(emdee_get_forces(s, forces));
        if (bitmask & EMDEE_ENERGIES) EMDEE_TRY(emdee_get_energies(s, energies));
        if (bitmask & EMDEE_VIRIALS) EMDEE_TRY(emdee_get_virials(s, virials));
        return EMDEE_OK;
    }
    if (!c->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    while ((int)c->chunk_events.size() < plan.nchunks) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->chunk_events.push_back(e);
    }
    double *tF = (bitmask & EMDEE_FORCES) ? s->tmp : nullptr, *tE = (bitmask & EMDEE_ENERGIES) ? s->tmp + 3 * s->N : nullptr,
           *tW = (bitmask & EMDEE_VIRIALS) ? s->tmp + 4 * s->N : nullptr;
    const bool build = !s->list_valid || s->list_n3;      // (a half list left by Newton's-third-law stepping is rebuilt in full form)
    s->build_n3 = false;
    int rc_ = EMDEE_OK;
    for (int k = 0; k < plan.nchunks && rc_ == EMDEE_OK; k++) {
        s->range_first = layer_bound[k];
        s->range_count = layer_bound[k + 1] - layer_bound[k];
        if (build) rc_ = run_cells(s, EMDEE_FORCES, false, nullptr, 0, false, 1);
        if (rc_ == EMDEE_OK) rc_ = run_cells(s, bitmask, false, nullptr, 0, false, 2);
        if (rc_ != EMDEE_OK) break;
        const int64_t est = std::max<int64_t>(1024, s->N / plan.nchunks);
        k_get_chunk<<<(unsigned)ceil_div64(est, 256), 256, 0, c->stream>>>(s->cell_start, plan.cell_bound[k], plan.cell_bound[k + 1], A.id, s->f[0], s->f[1],
                                                                           s->f[2], s->en, s->vir, tF, tE, tW);
        c->launches++;
        CUDA_TRY(cudaEventRecord(c->chunk_events[k], c->stream));
        CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->chunk_events[k], 0));
        for (const auto &r : runs[k]) {
            if (tF) CUDA_TRY(cudaMemcpyAsync(forces + 3 * r.first, tF + 3 * r.first, sizeof(double) * 3 * r.second, cudaMemcpyDeviceToHost, c->copy_stream));
            if (tE) CUDA_TRY(cudaMemcpyAsync(energies + r.first, tE + r.first, sizeof(double) * r.second, cudaMemcpyDeviceToHost, c->copy_stream));
            if (tW) CUDA_TRY(cudaMemcpyAsync(virials + r.first, tW + r.first, sizeof(double) * r.second, cudaMemcpyDeviceToHost, c->copy_stream));
        }
    }
    s->range_first = 0; s->range_count = 0;
    if (rc_ != EMDEE_OK) { cudaStreamSynchronize(c->copy_stream); return rc_; }
    if (build) s->list_valid = true;
    s->last_mode = mode;
    s->last_bitmask = bitmask;
    s->forces_valid = (bitmask & EMDEE_FORCES) != 0;
    CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    EMDEE_TRY(check_launch("emdee_compute_nonbonded_into"));
    return check_device_flag(s, "emdee_compute_nonbonded_into");      // (a list or brick overflow invalidates what was copied)
}

extern "C" int emdee_get_forces_range(emdee_system *s, int64_t id_first, int64_t count, double *out)
{
    SYS_ENTER(s, "emdee_get_forces_range");
    if (!(s->last_bitmask & EMDEE_FORCES)) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_forces_range: the last compute did not select FORCES");
    EMDEE_TRY(check_device_flag(s, "emdee_get_forces_range"));
    return get3_range(s, s->f, id_first, count, out, "emdee_get_forces_range");
}
extern "C" int emdee_get_energies_range(emdee_system *s, int64_t id_first, int64_t count, double *out)
{
    SYS_ENTER(s, "emdee_get_energies_range");
    if (!(s->last_bitmask & EMDEE_ENERGIES)) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_energies_range: the last compute did not select ENERGIES");
    EMDEE_TRY(check_device_flag(s, "emdee_get_energies_range"));
    return get1_range(s, s->en, id_first, count, out, "emdee_get_energies_range");
}
extern "C" int emdee_get_virials_range(emdee_system *s, int64_t id_first, int64_t count, double *out)
{
    SYS_ENTER(s, "emdee_get_virials_range");
    if (!(s->last_bitmask & EMDEE_VIRIALS)) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_virials_range: the last compute did not select VIRIALS");
    EMDEE_TRY(check_device_flag(s, "emdee_get_virials_range"));
    return get1_range(s, s->vir, id_first, count, out, "emdee_get_virials_range");
}

__global__ void k_sum_ew(int64_t first, int64_t n, const double *__restrict__ en, const double *__restrict__ vir,
                         int do_e, int do_w, double *__restrict__ partial)
{
    __shared__ double sE[256], sW[256];
    double E = 0, W = 0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        if (do_e) E += en[first + k];
        if (do_w) W += vir[first + k];
    }
    sE[threadIdx.x] = E; sW[threadIdx.x] = W;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sE[threadIdx.x] += sE[threadIdx.x + o]; sW[threadIdx.x] += sW[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = sE[0]; partial[2 * blockIdx.x + 1] = sW[0]; }
}

extern "C" int emdee_get_totals(emdee_system *s, double *E, double *W, int64_t *npairs)
{
    SYS_ENTER(s, "emdee_get_totals");
    if (s->last_mode < 0) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_totals: nothing has been computed");
    EMDEE_TRY(check_device_flag(s, "emdee_get_totals"));
    // per-atom arrays summed in slot order with a fixed tree: deterministic
    const int nb = 64;
    EMDEE_TRY(ensure_tmp(s, sizeof(double) * 2 * nb));
    k_sum_ew<<<nb, 256, 0, c->stream>>>(s->nlo, s->nown, s->en, s->vir, (s->last_bitmask & EMDEE_ENERGIES) != 0,
                                        (s->last_bitmask & EMDEE_VIRIALS) != 0, s->tmp);
    c->launches++;
    double h[2 * 64];
    CUDA_TRY(cudaMemcpyAsync(h, s->tmp, sizeof(double) * 2 * nb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    double e = 0, w = 0;
    for (int k = 0; k < nb; k++) { e += h[2 * k]; w += h[2 * k + 1]; }
    if (E) *E = e;
    if (W) *W = w;
    if (npairs) {
        *npairs = -1;
        if (s->last_mode == EMDEE_CUTOFF) {
            uint64_t d[3];
            EMDEE_TRY(emdee_pair_set_digest(s, d));
            *npairs = (int64_t)d[0];
        } else
            *npairs = s->N * (s->N - 1) / 2;
    }
    return check_launch("emdee_get_totals");
}

extern "C" int emdee_pair_set_digest(emdee_system *s, uint64_t out[3])
{
    SYS_ENTER(s, "emdee_pair_set_digest");
    if (!out) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_pair_set_digest: null output");
    if (!s->binned || !s->has_atoms) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_pair_set_digest: needs LJ atoms, positions and emdee_bin");
    // the audit pass re-evaluates every pair; its forces / energies / virials go to scratch arrays, so this call changes
    // neither the results of the last compute nor which getters are valid
    if (s->grid_ok)
        EMDEE_TRY(run_cells(s, EMDEE_FORCES | EMDEE_ENERGIES | EMDEE_VIRIALS, true, nullptr, 0));
    else {
        if (!s->tiles_default) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_pair_set_digest: box too small for a cell grid and a custom tile list is set");
        EMDEE_TRY(run_tiles(s, EMDEE_FORCES | EMDEE_ENERGIES | EMDEE_VIRIALS, true, nullptr, 0, true));
    }
    unsigned long long h[4];
    CUDA_TRY(cudaMemcpyAsync(h, s->digest, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    EMDEE_TRY(check_device_flag(s, "emdee_pair_set_digest"));
    out[0] = h[0]; out[1] = h[1]; out[2] = h[2];
    return EMDEE_OK;
}

extern "C" int emdee_pair_set(emdee_system *s, int32_t *ij, int64_t cap, int64_t *n)
{
    SYS_ENTER(s, "emdee_pair_set");
    if (!ij || !n || cap < 0) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_pair_set: bad arguments");
    if (!s->binned || !s->has_atoms) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_pair_set: needs LJ atoms, positions and emdee_bin");
    if (!s->grid_ok && !s->tiles_default) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_pair_set: box too small for a cell grid and a custom tile list is set");
    int32_t *d = nullptr;
    EMDEE_TRY(dev_alloc(&d, (size_t)2 * std::max<int64_t>(cap, 1)));
    int st = s->grid_ok ? run_cells(s, 7, true, d, cap) : run_tiles(s, 7, true, d, cap, true);
    unsigned long long h[4] = {0, 0, 0, 0};
    if (st == EMDEE_OK && cudaMemcpyAsync(h, s->digest, sizeof(h), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) st = EMDEE_ERR_CUDA;
    if (st == EMDEE_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) st = EMDEE_ERR_CUDA;
    if (st == EMDEE_OK) {
        *n = (int64_t)h[3];
        const int64_t m = std::min<int64_t>(*n, cap);
        if (cudaMemcpy(ij, d, sizeof(int32_t) * 2 * m, cudaMemcpyDeviceToHost) != cudaSuccess) st = EMDEE_ERR_CUDA;
        // lexicographic order (i, j): the device emits pairs in arrival order
        struct P { int32_t i, j; };
        P *p = reinterpret_cast<P *>(ij);
        std::sort(p, p + m, [](const P &x, const P &y) { return x.i != y.i ? x.i < y.i : x.j < y.j; });
    }
    dev_free(d);
    if (st != EMDEE_OK) EMDEE_FAIL(st, "emdee_pair_set: device failure (%s)", cudaGetErrorString(cudaGetLastError()));
    EMDEE_TRY(check_device_flag(s, "emdee_pair_set"));
    if (*n > cap) EMDEE_FAIL(EMDEE_ERR_CAPACITY, "emdee_pair_set: %lld pairs exceed the buffer capacity %lld", (long long)*n, (long long)cap);
    return EMDEE_OK;
}

extern "C" int emdee_list_pair_count(emdee_system *s, int64_t *npairs)
{
    SYS_ENTER(s, "emdee_list_pair_count");
    if (!npairs) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_list_pair_count: null output");
    *npairs = -1;
    if (!list_capable(s)) return EMDEE_OK;
    if (!s->list_valid) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_list_pair_count: no valid pair list (call emdee_vv_step first)");
    CUDA_TRY(cudaMemsetAsync(s->digest, 0, 4 * sizeof(unsigned long long), c->stream));
    // (what the stepping kernel walked last: the inner list if the last launch pruned into it or replayed it at these positions)
    s->lm = s->inner_at_current && s->inner_gen != 0 && s->inner_gen == s->list_gen && s->fl_persistent ? 2 : 0;
    const int rc_ = run_cells(s, EMDEE_FORCES, false, nullptr, 0, false, 3);
    s->lm = 0;
    EMDEE_TRY(rc_);
    unsigned long long n = 0;
    CUDA_TRY(cudaMemcpyAsync(&n, s->digest, sizeof(n), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    EMDEE_TRY(check_device_flag(s, "emdee_list_pair_count"));
    *npairs = (int64_t)(n / 2);      // the full-neighbour kernel sees every pair from both sides
    return EMDEE_OK;
}

// ------------------------------------------------------------------------------------------------
// velocity-Verlet
// ------------------------------------------------------------------------------------------------
static int launch_vv(emdee_system *s, double dt, int drift, int check_skin, bool track = false)
{
    emdee_ctx *c = s->ctx;
    AtomArrays &A = s->A[s->cur];
    VVArgs a;
    a.first = s->nlo;
    a.n = s->nown;
    for (int k = 0; k < 3; k++) { a.r[k] = A.r[k]; a.s[k] = A.s[k]; a.v[k] = A.v[k]; a.rb[k] = A.rb[k]; a.f[k] = s->f[k]; }
    a.mass = A.mass;
    a.dt = dt;
    a.L = s->L;
    a.half_skin2 = 0.25 * s->skin * s->skin;
    a.pending_kick = s->kick_pending ? 1 : 0;
    a.drift = drift;
    a.check_skin = check_skin;
    a.err = s->err;
    a.maxd2 = track ? s->maxd2 : nullptr;
    LAUNCH_1D(c, k_vv, a.n, a);
    return check_launch("k_vv");
}

extern "C" int emdee_vv_step(emdee_system *s, double dt, int64_t nsteps, int rebin_every)
{
    SYS_ENTER(s, "emdee_vv_step");
    if (!s->has_model || !s->has_atoms || !s->has_pos) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_vv_step: set model, LJ atoms and positions first");
    if (!s->binned) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_vv_step: call emdee_bin first");
    if (!s->forces_valid) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_vv_step: call emdee_compute_nonbonded(EMDEE_CUTOFF, FORCES|...) first");
    if (s->last_mode != EMDEE_CUTOFF) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_vv_step: forces must come from EMDEE_CUTOFF mode");
    if (!(dt > 0) || nsteps < 0) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_vv_step: dt=%g, nsteps=%lld", dt, (long long)nsteps);
    const bool adaptive = rebin_every < 0 && s->skin > 0;      // rebin_every < 0: re-bin when the skin is used up
    int64_t resume_at = -1;          // >= 0: the fused loop handed this step over after its kick, drift and re-binning
    // slab decomposition: the fused loop needs the neighbours' position pools mapped (peer halos) and a re-binning decision
    // every rank takes alike without a read-back (fixed cadence, or never)
    const bool slab_fused = c->nranks > 1 && s->decomposed && s->peer_ok;
    // slab runs with rebin_every < 0 re-bin at an interval chosen at the previous re-binning (do_bin_slab) instead of
    // reading the displacement back on every step
    const bool slab_adaptive = slab_fused && adaptive;
    const bool adaptive1 = adaptive && !slab_adaptive;      // per-step read-back (single GPU)
    if (s->fuse_vv && s->n14 == 0 && (c->nranks == 1 || slab_fused) && list_capable(s) && s->fl_persistent && nsteps > 0) {
        // One kernel per step: the stepping kernel's epilogue finishes step n (second half-kick) and starts step n+1
        // (first half-kick, drift, s = r/L into the second buffer) for every atom as soon as its force is known.
        // k_vv only starts the first step of the call.
        for (int k = 0; k < 3; k++)
            if (!s->s_alt[k]) EMDEE_TRY(dev_alloc(&s->s_alt[k], (size_t)s->cap + 2));
        bool drifted = false;
        // two-level list: with the per-step read-back of the adaptive re-binning the host also knows the largest step any atom has
        // taken since the last prune step, so it can let the kernel replay the inner list while no atom can have moved skin2 / 2
        // (steps since the prune x largest step), and ask for a new prune step otherwise.  The first step of a call drifts in
        // k_vv (not tracked): it prunes.
        const bool inner_ok = adaptive1 && c->nranks == 1 && s->skin2 > 0 && s->skin2 < s->skin && !s->want_n3 && !s->want_tma;
        s->inner_gen = 0;
        s->inner_at_current = false;
        float step2max = 0.0f;
        const char *ue = getenv("EMDEE_DEBUG_UNFUSE_AT");
        const int64_t unfuse_at = ue ? atoll(ue) : -1;
        for (int64_t st = 0; st < nsteps; st++) {
            // steps between re-binnings: the caller's, or (slab runs with rebin_every < 0) the interval the last re-binning chose
            auto every = [&]() { return rebin_every > 0 ? rebin_every : (slab_adaptive ? s->slab_interval : 0); };
            bool rebin = !s->list_valid || (every() > 0 && s->steps_since_bin + 1 >= every());
            if (!drifted) {
                EMDEE_TRY(launch_vv(s, dt, 1, rebin || adaptive1 ? 0 : 1, (adaptive1 && !rebin) || slab_adaptive));
                s->ghost_state = 1;        // slab: the neighbours' atoms moved too
            }
            if (adaptive1 && !rebin) {
                unsigned bits[3] = {0, 0, 0};
                CUDA_TRY(cudaMemcpyAsync(bits, s->maxd2, sizeof(bits), cudaMemcpyDeviceToHost, c->stream));
                CUDA_TRY(cudaStreamSynchronize(c->stream));
                float d2max;
                memcpy(&d2max, &bits[0], 4);
                memcpy(&step2max, &bits[2], 4);
                rebin = (double)d2max > 0.25 * s->skin * s->skin;
            }
            s->kick_pending = false;
            s->steps_since_bin++;
            if (rebin) { EMDEE_TRY(c->nranks > 1 ? do_bin_slab(s, s->ndiv) : do_bin(s, s->ndiv)); s->counters[0]++; }      // (a slab re-binning also refreshes the ghosts)
            if (!(list_capable(s) && s->fl_persistent) || (rebin && st > 0 && st == unfuse_at)) {
                // the re-binning chose bricks whose two staging buffers no longer fit (denser cells): this step's atoms are
                // already kicked, drifted and re-binned, so evaluate its forces with the generic path and carry on there
                if (st == unfuse_at) s->fl_persistent = false;      // EMDEE_DEBUG_UNFUSE_AT=<step>: exercise this branch (tests)
                resume_at = st;
                break;
            }
            if (!s->list_valid) {
                s->build_n3 = n3_usable(s);
                EMDEE_TRY(run_cells(s, EMDEE_FORCES, false, nullptr, 0, false, 1));
                s->list_valid = true;
            }
            const bool last = st == nsteps - 1;
            const bool next_rebin = every() > 0 && s->steps_since_bin + 1 >= every();
            s->vv_mode = last ? 1 : 2;
            s->vv_dt = dt;
            s->vv_check_skin = (!adaptive1 && !next_rebin) ? 1 : 0;
            s->vv_track = adaptive1 || slab_adaptive;
            if (slab_fused) {
                // ghost positions: exchanged by NCCL after k_vv started this call (state 1), complete after a re-binning (0), or
                // being written by the neighbours' previous launch, whose id their flags will carry (2)
                if (s->ghost_state == 1) { EMDEE_TRY(slab_halo_positions(s, c->stream)); s->ghost_state = 0; }
                s->p2p_launch = true;
                s->p2p_wait = s->ghost_state == 2 ? s->epoch : 0;
                s->epoch++;
                s->p2p_publish = (!last && !next_rebin) ? s->epoch : 0;     // a re-binning exchanges the ghosts itself
            }
            s->lm = 0;
            if (inner_ok && !s->list_n3 && !s->fl_tma && !s->fl_split && s->fl_qcap == FL_QCAP) {      // (the dense-cell kernels have no prune / replay form)
                const bool have = s->inner_gen != 0 && s->inner_gen == s->list_gen && s->steps_since_prune >= 1;
                const double moved = (double)s->steps_since_prune * std::sqrt((double)step2max) * (1.0 + 1e-6);
                s->lm = have && moved <= 0.5 * s->skin2 ? 2 : 1;
                if (const char *e = getenv("EMDEE_DEBUG_LM")) {      // timing experiments only: 2 replays a stale inner list (wrong forces), 1 prunes on every step
                    if (atoi(e) == 2 && have) s->lm = 2;
                    if (atoi(e) == 1) s->lm = 1;
                }
                if (s->lm == 1) CUDA_TRY(cudaMemsetAsync(s->maxd2 + 2, 0, sizeof(unsigned), c->stream));
            }
            const int lm_ = s->lm;
            const int rc_ = run_cells(s, EMDEE_FORCES, false, nullptr, 0, false, 2);
            s->vv_mode = 0;
            s->p2p_launch = false;
            s->lm = 0;
            EMDEE_TRY(rc_);
            if (lm_ == 1) { s->inner_gen = s->list_gen; s->steps_since_prune = 0; }
            if (lm_ != 0 && !last) s->steps_since_prune++;
            s->inner_at_current = lm_ != 0 && last;
            if (slab_fused) s->ghost_state = s->p2p_publish ? 2 : (last ? 0 : 1);
            if (!last) {
                AtomArrays &Ac = s->A[s->cur];
                for (int k = 0; k < 3; k++) std::swap(Ac.s[k], s->s_alt[k]);      // the drifted positions become current
                drifted = true;
            }
        }
        if (resume_at < 0) {
            s->kick_pending = false;
            s->last_bitmask = EMDEE_FORCES;
            s->forces_valid = true;
            return EMDEE_OK;
        }
    }
    for (int64_t st = std::max<int64_t>(resume_at, 0); st < nsteps; st++) {
        // A pair list must be built at the positions the cells were binned at (both rely on "no atom moved
        // more than skin/2 since the binning"), so a missing list forces a re-binning on this step.
        const bool need_list = list_capable(s) && !s->list_valid;
        bool rebin = need_list || (rebin_every > 0 && s->steps_since_bin + 1 >= rebin_every);
        if (st == resume_at) {
            rebin = true;            // taken over from the fused loop: kicked, drifted and re-binned already
        } else if (adaptive && !rebin) {
            // re-bin exactly when an atom has moved more than skin/2 since the last binning: one 4-byte read-back per
            // step (all ranks must take the same decision: max over ranks)
            EMDEE_TRY(launch_vv(s, dt, 1, 0, true));
            if (c->nranks > 1) NCCL_TRY(g_nccl.AllReduce(s->maxd2, s->maxd2, 1, ncclUint32, ncclMax, c->comm, c->stream));
            unsigned bits = 0;
            CUDA_TRY(cudaMemcpyAsync(&bits, s->maxd2, sizeof(bits), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            float d2max;
            memcpy(&d2max, &bits, 4);
            rebin = (double)d2max > 0.25 * s->skin * s->skin;
        } else
            EMDEE_TRY(launch_vv(s, dt, 1, rebin || adaptive ? 0 : 1));     // [kick2 of the previous step] + kick1 + drift
        if (st != resume_at) {
            s->kick_pending = false;
            s->steps_since_bin++;
            if (rebin) EMDEE_TRY(c->nranks > 1 ? do_bin_slab(s, s->ndiv) : do_bin(s, s->ndiv));
        }
        if (s->grid_ok) {
            // pair list: built (k_list_build, a filter) right after a (re-)binning, walked (k_force_list) on every step
            if (list_capable(s)) {
                if (!s->list_valid) {
                    EMDEE_TRY(run_cells(s, EMDEE_FORCES, false, nullptr, 0, false, 1));
                    s->list_valid = true;
                }
                EMDEE_TRY(run_cells(s, EMDEE_FORCES, false, nullptr, 0, !rebin, 2));   // a re-bin already refreshed the ghosts
            } else {
                EMDEE_TRY(run_cells(s, EMDEE_FORCES, false, nullptr, 0, !rebin, 0));
                s->list_valid = false;      // a list built for a single-point evaluation does not survive a drift without skin
            }
        }
        else
            EMDEE_TRY(run_tiles(s, EMDEE_FORCES, true));
        EMDEE_TRY(apply_pairs14(s, EMDEE_FORCES));
        s->kick_pending = true;
    }
    if (s->kick_pending) {
        EMDEE_TRY(launch_vv(s, dt, 0, 0));                  // final second half-kick
        s->kick_pending = false;
    }
    s->last_bitmask = EMDEE_FORCES;
    s->forces_valid = true;
    return EMDEE_OK;
}

extern "C" int emdee_scale_velocities(emdee_system *s, double factor)
{
    SYS_ENTER(s, "emdee_scale_velocities");
    if (!s->has_vel) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_scale_velocities: velocities were never set");
    if (!std::isfinite(factor)) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_scale_velocities: factor=%g", factor);
    if (s->kick_pending) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_scale_velocities: a half-kick is pending (call between emdee_vv_step calls)");
    LAUNCH_1D(c, k_scale3, s->nown, s->nlo, s->nown, factor, A.v[0], A.v[1], A.v[2]);
    return check_launch("k_scale3");
}

extern "C" int emdee_get_step_counters(emdee_system *s, int64_t out[4])
{
    SYS_ENTER(s, "emdee_get_step_counters");
    if (!out) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_get_step_counters: null output");
    for (int k = 0; k < 4; k++) out[k] = s->counters[k];
    return EMDEE_OK;
}

extern "C" int emdee_get_step_config(emdee_system *s, int32_t out[8])
{
    SYS_ENTER(s, "emdee_get_step_config");
    if (!out) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_get_step_config: null output");
    if (!s->binned) EMDEE_FAIL(EMDEE_ERR_STATE, "emdee_get_step_config: call emdee_bin first");
    const bool listed = list_capable(s);
    out[0] = s->grid_ok ? s->g.bx : 0;
    out[1] = s->grid_ok ? s->g.by : 0;
    out[2] = s->grid_ok ? s->g.bz : 0;
    out[3] = s->grid_ok ? s->fc_cap : 0;      // (full staging; a compacted persistent configuration stages fl_cap atoms)
    out[4] = listed ? 1 : 0;
    // bit 1: staging by bulk asynchronous copies (EMDEE_TMA=1); bit 2: compacted staging (dense cells); bit 3: shallow stacks; bit 4: split lists
    out[5] = listed && s->fl_persistent ? (1 | (s->fl_tma ? 2 : 0) | (s->fl_compact ? 4 : 0) | (s->fl_qcap != FL_QCAP ? 8 : 0) | (s->fl_split ? 16 : 0)) : 0;
    if (listed && s->fl_persistent) out[3] = s->fl_cap;
    out[6] = listed && s->fl_persistent && s->fuse_vv && s->n14 == 0 && (c->nranks == 1 || s->peer_ok) ? 1 : 0;
    out[7] = s->lcap8;
    return EMDEE_OK;
}

extern "C" int emdee_kinetic_energy(emdee_system *s, double *K)
{
    SYS_ENTER(s, "emdee_kinetic_energy");
    if (!K) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_kinetic_energy: null output");
    EMDEE_TRY(check_device_flag(s, "emdee_kinetic_energy"));
    const int nb = 128;
    EMDEE_TRY(ensure_tmp(s, sizeof(double) * nb));
    k_kinetic<<<nb, 256, 0, c->stream>>>(s->nlo, s->nown, A.v[0], A.v[1], A.v[2], A.mass, s->tmp);
    c->launches++;
    double h[128];
    CUDA_TRY(cudaMemcpyAsync(h, s->tmp, sizeof(double) * nb, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    double k = 0;
    for (int i = 0; i < nb; i++) k += h[i];
    *K = k;
    return check_launch("k_kinetic");
}

extern "C" int emdee_profile_begin(emdee_system *s)
{
    SYS_ENTER(s, "emdee_profile_begin");
    s->profiling = true;
    s->prof_used = 0;
    return EMDEE_OK;
}
extern "C" int emdee_profile_end(emdee_system *s, double *ms, int64_t *launches)
{
    SYS_ENTER(s, "emdee_profile_end");
    s->profiling = false;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    double total = 0, per_mode[4] = {0, 0, 0, 0};
    int n_mode[4] = {0, 0, 0, 0};
    for (size_t k = 0; k + 1 < s->prof_used; k += 2) {
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, s->prof_events[k], s->prof_events[k + 1]));
        total += t;
        const int m = s->prof_mode[k / 2] & 3;
        per_mode[m] += t; n_mode[m]++;
    }
    for (int m = 0; m < 4; m++) { s->prof_ms[m] = per_mode[m]; s->prof_n[m] = n_mode[m]; }
    if (!s->prof_mid.empty()) {       // slab runs: interior launch / wait for the halo / boundary launch
        double ti = 0, tw = 0, tb = 0;
        for (size_t k = 0; k < s->prof_mid_of.size(); k++) {
            float x = 0;
            const size_t p = s->prof_mid_of[k];
            CUDA_TRY(cudaEventElapsedTime(&x, s->prof_events[p], s->prof_mid[2 * k])); ti += x;
            CUDA_TRY(cudaEventElapsedTime(&x, s->prof_mid[2 * k], s->prof_mid[2 * k + 1])); tw += x;
            CUDA_TRY(cudaEventElapsedTime(&x, s->prof_mid[2 * k + 1], s->prof_events[p + 1])); tb += x;
        }
        const double n = (double)s->prof_mid_of.size();
        fprintf(stderr, "[emdee] rank %d slab step: interior %.4f ms, waiting for the halo %.4f ms, boundary %.4f ms (%d steps)\n", c->rank,
                ti / n, tw / n, tb / n, (int)n);
        for (cudaEvent_t e : s->prof_mid) cudaEventDestroy(e);
        s->prof_mid.clear(); s->prof_mid_of.clear();
    }
    if (g_rebin_count && getenv("EMDEE_DEBUG") && atoi(getenv("EMDEE_DEBUG")) >= 2) {
        fprintf(stderr, "[emdee] rank %d: %d slab re-binnings, wall ms each: migration %.3f, populations+scan+marks %.3f, sort+gather %.3f, ghosts %.3f, bricks %.3f\n",
                c->rank, g_rebin_count, g_rebin_phase[0] / g_rebin_count, g_rebin_phase[1] / g_rebin_count, g_rebin_phase[2] / g_rebin_count,
                g_rebin_phase[3] / g_rebin_count, g_rebin_phase[4] / g_rebin_count);
        for (int k = 0; k < 8; k++) g_rebin_phase[k] = 0;
        g_rebin_count = 0;
    }
#if FLP_TIMING
    {
        unsigned long long t[8];
        CUDA_TRY(cudaMemcpy(t, s->digest + 8, sizeof(t), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemset(s->digest + 8, 0, sizeof(t)));
        const double P = (double)t[3], Cn = (double)t[5];
        fprintf(stderr, "[emdee] role timers: producers: wait-empty %.1f%% stage %.1f%% integrate %.1f%% other %.1f%% of their time (%.0f cycles per brick); "
                        "consumers: wait-full %.1f%% of their time; %llu brick periods\n", 100 * t[0] / P, 100 * t[1] / P, 100 * t[2] / P,
                100 * (P - t[0] - t[1] - t[2]) / P, P / std::max(1.0, (double)t[6]), 100 * t[4] / Cn, t[6]);
    }
#endif
    if (getenv("EMDEE_DEBUG"))
        for (int m = 0; m < 4; m++)
            if (n_mode[m]) fprintf(stderr, "[emdee] force kernel mode %d: %d launches, %.4f ms each\n", m, n_mode[m], per_mode[m] / n_mode[m]);
    if (ms) *ms = total;
    if (launches) *launches = (int64_t)(s->prof_used / 2);
    return EMDEE_OK;
}

extern "C" int emdee_profile_kind(emdee_system *s, int kind, double *ms, int64_t *launches)
{
    SYS_ENTER(s, "emdee_profile_kind");
    if (kind < 0 || kind > 2) EMDEE_FAIL(EMDEE_ERR_INVALID, "emdee_profile_kind: kind %d (0 window scan, 1 list build, 2 list walk)", kind);
    if (ms) *ms = s->prof_ms[kind];
    if (launches) *launches = s->prof_n[kind];
    return EMDEE_OK;
}

extern "C" int emdee_synchronize(emdee_system *s)
{
    SYS_ENTER(s, "emdee_synchronize");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return check_device_flag(s, "emdee_synchronize");
}

// ------------------------------------------------------------------------------------------------
// one-shot mirror of compute_nonbonded! on host arrays (src/nonbonded.jl:109-120)
// ------------------------------------------------------------------------------------------------
extern "C" int emdee_compute_nonbonded_host(int64_t N, const double *pos, double L, double cutoff, double sw,
                                            const double *atoms, const int32_t *tiles, int64_t ntiles, int mode, int ndiv,
                                            int bitmask, double *forces, double *energies, double *virials)
{
    emdee_ctx *c = nullptr;
    emdee_system *s = nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    int st = emdee_create(&c, dev);
    if (st == EMDEE_OK) st = emdee_system_create(c, N, L, &s);
    if (st == EMDEE_OK) st = emdee_set_model(s, cutoff, sw);
    if (st == EMDEE_OK) st = emdee_set_lj_atoms(s, atoms);
    if (st == EMDEE_OK) st = emdee_set_positions(s, pos);
    if (st == EMDEE_OK && tiles) st = emdee_set_tiles(s, tiles, ntiles);
    if (st == EMDEE_OK && mode == EMDEE_CUTOFF) st = emdee_bin(s, ndiv);
    if (st == EMDEE_OK) st = emdee_compute_nonbonded_into(s, mode, bitmask, forces, energies, virials);
    emdee_system_destroy(s);
    emdee_destroy(c);
    return st;
}
