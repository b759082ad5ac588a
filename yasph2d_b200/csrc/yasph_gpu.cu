// yasph_gpu.cu -- context, step orchestration and the C ABI (include/yasph_gpu.h) of libyasph_gpu.so.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false (see __graft_entry__.build()).
// There is no CPU fallback anywhere in this library: without a CUDA device yasph_create fails with
// YASPH_ERR_NO_DEVICE.
#include "../../include/yasph_gpu.h"

#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <atomic>
#include <string>
#include <vector>

#include "scan.cuh"
#include "slab.cuh"
#include "sweeps.cuh"

using namespace yasph;

static thread_local std::string g_create_error;

constexpr uint32_t UNSTAGED_GRID_MAX = 1024;
enum SlabField { SF_POS = 0, SF_VEL, SF_VSTAR, SF_DENS, SF_ALPHA, SF_KFAC, SF_KAPPA, SF_STIFF, SF_ACCEL, SLAB_NUM_FIELDS };  // CTAs of k_sweep_unstaged (their partial reductions follow k_sweep's in yasph_ctx::partials)

struct PassEvent {
    int pass;
    cudaEvent_t a, b;
};

struct yasph_ctx {
    yasph_config cfg;
    std::string err;
    int device = 0, num_sms = YASPH_NUM_SMS_B200;
    cudaStream_t stream = nullptr;
    uint32_t n = 0, m = 0;
    uint32_t cap_n = 0, cap_m = 0, max_tiles = 0;
    uint32_t lim_dyn = 0, lim_stat = 0;      // configured upper limits for a tile's staged candidates (0 = what shared memory allows)
    uint32_t cap_dyn = 0, cap_stat = 0;      // staging capacity of the current neighbourhood structure (largest tile, rounded up)
    uint32_t cap_pc = 0;                     // particles of the largest tile, rounded up
    uint32_t num_tiles = 0;                  // host copy of Control::num_tiles for the current structure == grid of the tile kernels
    size_t smem_optin = 0;
    float mass = 0, radius = 0;
    GridParams grid;
    KernelConsts kc;
    TimeParams tp;
    // particle state (ping-pong pairs are swapped by the gather)
    float2* pos_adv = nullptr;  // advected positions written by the density solver's last Jacobi B (OpJacobiBAdvect); swapped with pos afterwards
    float2 *pos = nullptr, *pos_alt = nullptr, *vel = nullptr, *vel_alt = nullptr, *vstar = nullptr, *vstar_alt = nullptr, *accel = nullptr;
    float *dens = nullptr, *alpha = nullptr, *kappa = nullptr, *stiff = nullptr, *err_buf = nullptr, *f_alt0 = nullptr, *f_alt1 = nullptr;
    uint32_t *keys[2] = {nullptr, nullptr}, *idx[2] = {nullptr, nullptr};
    uint32_t *cell_key = nullptr, *cell_start = nullptr, *tile_key = nullptr, *tile_pstart = nullptr, *tile_cstart = nullptr;
    // boundary
    float2 *bpos = nullptr, *bpos_alt = nullptr;
    uint32_t *scell_key = nullptr, *scell_start = nullptr, *stile_key = nullptr, *stile_cstart = nullptr;
    // tiles and lists
    TileRuns* tile_runs = nullptr;
    uint32_t *cslot_d = nullptr, *cslot_s = nullptr;
    unsigned long long* lists = nullptr;
    uint32_t* counts = nullptr;   // per particle: count_dynamic | count_total << 8
    uint32_t* tile_nk = nullptr;  // per tile: most list words of any of its particles
    uint32_t* apron_idx = nullptr;  // per tile: global index of its first APRON_TABLE apron slots (list build -> sweeps)
    // scratch
    // Radix sort scratch (tile tickets, digit histograms, look-back status): two areas used alternately.  The area of the NEXT sort is
    // zeroed on the side stream while the cell / tile tables of the current update are built (neighborhood_update), so no memset
    // sits between the kernels of the step's chain; radix_prepare falls back to a memset in the main stream when that has not happened.
    uint32_t* radix_base[2] = {nullptr, nullptr};
    uint32_t* radix_scratch = nullptr;   // the area of the running / next sort (radix_prepare)
    int radix_cur = 0;
    uint32_t radix_clean_n[2] = {0u, 0u};  // pairs the area is known to be zeroed for (0: dirty)
    unsigned long long *scan_chunks = nullptr, *scan_total = nullptr, *scan_status = nullptr;
    double* partials = nullptr;
#if defined(YASPH_SWEEP_TIMING) || defined(YASPH_LIST_TIMING)
    unsigned long long* sweep_dbg = nullptr;
#endif
    Control* ctl = nullptr;
    Control* h_ctl = nullptr;  // pinned mirror
    // read-back channel: a one-warp kernel copies the control block into mapped host memory and bumps `seq`; the host polls it
    struct Published {
        Control ctl;
        unsigned int seq;
    };
    Published* h_pub = nullptr;  // mapped pinned
    Published* d_pub = nullptr;  // its device address
    Control* ctl_snap = nullptr; // device copy of the control block taken by k_begin_step (yasph_step_n)
    unsigned int pub_seq = 0;
    // export scratch (allocated on demand)
    uint16_t *exp_cd = nullptr, *exp_ct = nullptr;
    uint32_t* exp_lists = nullptr;
    // pinned staging for yasph_step_host
    float *h_stage = nullptr;
    size_t h_stage_bytes = 0;
    // yasph_step_host: results that are final before the step ends (sorted positions after the gather, densities after the
    // density sweep) leave on a second stream while the rest of the step computes; only the velocities wait for the last pass
    cudaStream_t copy_stream = nullptr;
    cudaStream_t ctl_stream = nullptr;  // publishes the control block while the early list build runs (never carries copies)
    cudaStream_t aux_stream = nullptr;  // the gather of a neighbourhood update, next to the cell / tile tables
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_early[2] = {nullptr, nullptr};
    float *early_pos_out = nullptr, *early_dens_out = nullptr, *early_vel_out = nullptr;  // pinned host destinations of the current yasph_step_host call
    bool early_vel_stale = false;  // the velocities changed after their speculative download (the solve needed more iterations)
    struct PendingDownload {
        float* host = nullptr;
        const void* dev = nullptr;
        size_t bytes = 0;
    } pending[2];  // positions, densities: event recorded, copy not submitted yet
    cudaEvent_t ev_host[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // yasph_host_step_times (YASPH_FLAG_PROFILE_PASSES)
    float host_us[6] = {0, 0, 0, 0, 0, 0};
    // particle ids (YASPH_FLAG_TRACK_IDS) and the slab decomposition (multi-GPU)
    uint32_t *ids = nullptr, *ids_alt = nullptr;
    struct Slab {
        bool active = false;
        int rank = 0, world = 1;
        uint32_t col_lo = 0, col_hi = 65536, id_base = 0;
        uint64_t n_global = 0;
        void* comm = nullptr;            // ncclComm_t
        struct Fabric* fabric = nullptr; // loopback transport (all ranks in one process) instead of NCCL
        cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
        uint64_t link_seq[2] = {0, 0};
        uint8_t* pflag = nullptr;        // [cap_n] ghost flag of the current structure / classification during an update
        uint32_t* sel[2] = {nullptr, nullptr};    // index lists of an update's four-way selection: migrants to the left | right rank [max_halo]
        uint32_t* sel_g[2] = {nullptr, nullptr};  // ... and the stayers in the first | last W owned columns (ghost layer of the left | right rank)
        bool halo_lists_valid = false;            // send_idx / ghost_idx describe the current structure (built on demand: ensure_halo_lists)
        uint32_t* send_idx[2] = {nullptr, nullptr};   // per-pass halo send lists (left, right), sorted order [max_halo]
        uint32_t* ghost_idx[2] = {nullptr, nullptr};  // ghosts from the left / right rank, sorted order [max_halo]
        uint32_t* own_idx = nullptr;     // owned particles, sorted order [cap_n] (built on demand)
        unsigned char* sbuf[2] = {nullptr, nullptr};  // send buffers (left, right) [max_halo * RECORD_MAX]
        unsigned char* rbuf[2] = {nullptr, nullptr};  // receive buffers
        uint32_t *d_cnt = nullptr, *h_cnt = nullptr;  // [4]: send left/right, recv left/right
        uint32_t max_halo = 0;
        uint32_t ghost_cols = 1;         // width W of the ghost layer in cell columns (yasph_config.ghost_columns)
        // Validity of the ghosts' copy of every per-particle field: the number of ghost columns, counted from the slab's edge, whose
        // values are the bits their owner holds.  A refresh (ghost records, halo exchange) sets W; a pass that gathers a field from
        // the neighbours produces results that are right one column less far out than its inputs (the ghosts recompute what their
        // owners compute: same arithmetic, same neighbour order).  A pass runs a halo exchange for an input only when that input is
        // stale already in the first ghost column.
        int valid[SLAB_NUM_FIELDS] = {};
        uint32_t n_own = 0, n_ghost[2] = {0, 0}, n_send[2] = {0, 0};
        uint32_t mig_out[2] = {0, 0}, mig_in = 0;
        bool own_idx_valid = false;
        bool own_count_unchecked = false;  // own_idx was built in mid-step without reading its count back (checked with the step's last control block)
        uint64_t halo_exchanges = 0, allreduces = 0;
        // peer-memory transport (slab.cuh): the own mailbox, the peers' mailboxes as mapped into this process, sequence numbers
        bool peer = false, peer_tried = false;
        void* box = nullptr;
        void* peer_box[PEER_MAX_RANKS] = {};
        PeerBoxHeader** d_boxes = nullptr;  // device copy of peer_box (all-reduce kernel)
        unsigned int* d_ticket = nullptr;
        uint64_t halo_seq = 0, ar_seq = 0;
        PeerCounts* h_pcounts = nullptr;  // mapped pinned: counts of the last record exchange
        PeerCounts* d_pcounts = nullptr;  // its device address
        uint32_t pcounts_seq = 0;
    } slab;
    // state flags
    bool have_particles = false, lists_valid = false, dfsph_ready = false;
    uint32_t dfsph_n = 0;  // length of the DFSPH solver arrays (alpha / kappa / stiffness) as of their last resize (dfsph.rs:419-423)
    int list_margin_pct = 12;
    bool fuse_advect = true;         // YASPH_DEBUG_NO_FUSED_ADVECT=1: the separate k_advect_keygen pass (A/B timing)
    bool advect_fused_now = false;   // the running density solve advects in its last Jacobi B
    bool scan_status_clean = false;  // scan_status is all zero (k_scan_fused cleans up after itself)
    bool pdl = true;  // programmatic dependent launch of the step's kernel chain (YASPH_DEBUG_NO_PDL=1 turns it off for A/B timing)
    bool spec_advect = false, spec_advect_done = false;  // advect + sort enqueued ahead of the density solver's read-back (dfsph_step)
    // yasph_step_n: the head of the NEXT step (k_begin_step + its first pass) is enqueued ahead of the read-back that ends the
    // running step, so the GPU does not idle through that round trip
    bool spec_head = false;      // armed by yasph_step_n for every step but the last
    bool head_enqueued = false;  // the next step's head is in the stream (and, DFSPH, the device let it run)
    uint32_t step_token = 0;
    uint64_t list_rebuilds = 0;     // early list builds that had to be repeated
    bool unstaged_tiles = false;    // the current structure has tiles beyond the (clipped) staging capacities: every tile kernel is followed / preceded by its unstaged twin
    bool lists_valid_once = false;  // cap_dyn / cap_stat / num_tiles describe an earlier structure of this particle set
    cudaEvent_t ev_tables = nullptr;
    uint64_t launches = 0;
    // profiling
    std::vector<PassEvent> events;
    std::vector<cudaEvent_t> event_pool;
    float pass_us[YASPH_NUM_PASSES];
    int cur_pass = -1;
    cudaEvent_t cur_start = nullptr;
};

// ---------------------------------------------------------------------------------------------------------------------
// error helpers
// ---------------------------------------------------------------------------------------------------------------------
static void fabric_mark_failed(yasph_ctx* c);
static void peer_teardown(yasph_ctx* c);
static int32_t fail(yasph_ctx* c, int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) {
        c->err = buf;
        fabric_mark_failed(c);  // loopback transport: wake the peer ranks of this process instead of leaving them waiting
    } else {
        g_create_error = buf;
    }
    return code;
}
#define CU(call)                                                                                                     \
    do {                                                                                                             \
        cudaError_t e_ = (call);                                                                                     \
        if (e_ != cudaSuccess) return fail(c, YASPH_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CHECK_LAUNCH()                                                                                               \
    do {                                                                                                             \
        c->launches++;                                                                                               \
        cudaError_t e_ = cudaGetLastError();                                                                         \
        if (e_ != cudaSuccess) return fail(c, YASPH_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define TRY(expr)                    \
    do {                             \
        int32_t r_ = (expr);         \
        if (r_ != YASPH_OK) return r_; \
    } while (0)

template <typename T>
static cudaError_t dmalloc(T** p, size_t count) {
    // 64 bytes of slack: the sweeps' 16-byte aligned bulk copies may read up to three elements past the last particle
    return cudaMalloc((void**)p, (count ? count : 1) * sizeof(T) + 64);
}

// ---------------------------------------------------------------------------------------------------------------------
// Loopback transport: every rank is a context of THIS process (one host thread per rank, any mix of devices).  It carries
// the same messages as the NCCL transport through device-to-device copies and exists so that the whole slab logic
// (migration, ghosts, halo exchange, all-reduce) can be exercised on a single GPU.
// ---------------------------------------------------------------------------------------------------------------------
struct Fabric {
    int world = 0;
    std::mutex m;
    std::condition_variable cv;
    struct Post {
        const void* buf = nullptr;
        size_t bytes = 0;
        cudaEvent_t ready = nullptr;
        uint64_t seq = 0;
    };
    struct Ack {
        cudaEvent_t done = nullptr;
        uint64_t seq = 0;
    };
    std::vector<Post> post;  // [rank * 2 + side]: what `rank` offers to its neighbour on `side`
    std::vector<Ack> ack;    // [rank * 2 + side]: `rank` has enqueued the copy of what its neighbour on `side` offered
    // all-reduce rendezvous
    std::vector<double> ar_val;
    uint64_t ar_gen = 0;
    int ar_count = 0;
    double ar_result = 0.0;
    bool failed = false;
};
static void fabric_mark_failed(yasph_ctx* c) {
    Fabric* f = c->slab.fabric;
    if (!f || !c->slab.active) return;
    {
        std::lock_guard<std::mutex> lk(f->m);
        f->failed = true;
    }
    f->cv.notify_all();
}

// ---------------------------------------------------------------------------------------------------------------------
// NCCL, bound at run time (dlopen) so that single-GPU use has no dependency on it.  If the process already holds a
// libnccl.so.2 (e.g. the one PyTorch ships) the dynamic loader hands back that instance.
// ---------------------------------------------------------------------------------------------------------------------
#include <nccl.h>
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    std::string error;
};
static NcclApi g_nccl;
static bool nccl_load() {
    if (g_nccl.handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        g_nccl.error = std::string("dlopen(libnccl.so.2) failed: ") + (dlerror() ? dlerror() : "?");
        return false;
    }
#define NCCL_SYM(field, name)                                                   \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));    \
    if (!g_nccl.field) {                                                        \
        g_nccl.error = std::string("NCCL symbol missing: ") + name;             \
        dlclose(h);                                                             \
        return false;                                                           \
    }
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
    NCCL_SYM(GroupStart, "ncclGroupStart")
    NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NCCL_SYM(Send, "ncclSend")
    NCCL_SYM(Recv, "ncclRecv")
    NCCL_SYM(AllReduce, "ncclAllReduce")
    NCCL_SYM(AllGather, "ncclAllGather")
#undef NCCL_SYM
    g_nccl.handle = h;
    return true;
}
#define NC(call)                                                                                                          \
    do {                                                                                                                  \
        ncclResult_t r_ = (call);                                                                                         \
        if (r_ != ncclSuccess) return fail(c, YASPH_ERR_COMM, "%s failed: %s (%s:%d)", #call, g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)
static void comm_destroy(yasph_ctx* c) {
    if (c->slab.comm && g_nccl.handle) g_nccl.CommDestroy((ncclComm_t)c->slab.comm);
    c->slab.comm = nullptr;
}

// ---------------------------------------------------------------------------------------------------------------------
// pass timing (YASPH_FLAG_PROFILE_PASSES)
// ---------------------------------------------------------------------------------------------------------------------
static cudaEvent_t get_event(yasph_ctx* c) {
    if (!c->event_pool.empty()) {
        cudaEvent_t e = c->event_pool.back();
        c->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
static const bool g_no_pass_events = getenv("YASPH_DEBUG_NO_PASS_EVENTS") != nullptr;  // host timeline without the per-pass events
static void pass_begin(yasph_ctx* c, int pass) {
    if (!(c->cfg.flags & YASPH_FLAG_PROFILE_PASSES) || g_no_pass_events) return;
    c->cur_pass = pass;
    c->cur_start = get_event(c);
    cudaEventRecord(c->cur_start, c->stream);
}
static void pass_end(yasph_ctx* c) {
    if (!(c->cfg.flags & YASPH_FLAG_PROFILE_PASSES) || g_no_pass_events || c->cur_pass < 0) return;
    cudaEvent_t b = get_event(c);
    cudaEventRecord(b, c->stream);
    c->events.push_back(PassEvent{c->cur_pass, c->cur_start, b});
    c->cur_pass = -1;
}
static void pass_resolve(yasph_ctx* c) {
    if (!(c->cfg.flags & YASPH_FLAG_PROFILE_PASSES)) return;
    for (int i = 0; i < YASPH_NUM_PASSES; ++i) c->pass_us[i] = 0.f;
    for (auto& e : c->events) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        c->pass_us[e.pass] += ms * 1000.f;
        c->event_pool.push_back(e.a);
        c->event_pool.push_back(e.b);
    }
    c->events.clear();
    float tot = 0.f;
    for (int i = 0; i < YASPH_PASS_TOTAL; ++i) tot += c->pass_us[i];
    c->pass_us[YASPH_PASS_TOTAL] = tot;
}

// ---------------------------------------------------------------------------------------------------------------------
// config
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int32_t yasph_config_default(yasph_config* cfg, float smoothing_factor, float particle_density, float fluid_density, int32_t solver) {
    if (!cfg) return YASPH_ERR_INVALID_ARGUMENT;
    memset(cfg, 0, sizeof(*cfg));
    cfg->abi_version = YASPH_ABI_VERSION;
    cfg->device = 0;
    cfg->max_particles = 1u << 20;
    cfg->max_boundary = 1u << 18;
    const float radius = 0.5f / sqrtf(particle_density);          // fluidparticleworld.rs:82-85
    cfg->smoothing_length = 2.0f * radius * smoothing_factor;    // fluidparticleworld.rs:58
    cfg->particle_density = particle_density;
    cfg->fluid_density = fluid_density;
    cfg->gravity[0] = 0.0f;
    cfg->gravity[1] = -9.81f;                                     // fluidparticleworld.rs:123
    cfg->grid_min[0] = -100.0f;
    cfg->grid_min[1] = -100.0f;                                   // neighborhood_search.rs:478
    cfg->solver = solver;
    cfg->viscosity = YASPH_VISCOSITY_XSPH;
    cfg->viscosity_param = 0.05f;                                 // xsph.rs:14
    cfg->dfsph_max_avg_density_error = 0.01f / 100.0f;            // dfsph.rs:49
    cfg->dfsph_max_density_iters = 200;
    cfg->dfsph_max_divergence_error = 0.1f / 100.0f;              // dfsph.rs:53
    cfg->dfsph_max_divergence_iters = 400;
    const float speed_of_sound = 1.0f / sqrtf(0.01f);             // wscsph.rs:47 with the defaults of wscsph.rs:39
    cfg->wcsph_stiffness = fluid_density * speed_of_sound * speed_of_sound / 7.0f;  // wscsph.rs:48
    cfg->wcsph_boundary_force_factor = 1.0f;                      // wscsph.rs:34
    cfg->adaptive_timestep = 1;
    cfg->timestep_fixed_ns = 0;
    cfg->timestep_min_ns = duration_from_secs_f32(1.0f / 60.0f / 400.0f);  // main.rs:124
    cfg->timestep_max_ns = duration_from_secs_f32(1.0f / 120.0f / 3.0f);   // main.rs:123
    cfg->cfl_factor = solver == YASPH_SOLVER_WCSPH ? 0.2f : 1.5f;          // main.rs:115-118
    return YASPH_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// create / destroy
// ---------------------------------------------------------------------------------------------------------------------
// every tile kernel may use up to the device's opt-in shared memory; the actual size is chosen per launch
template <class K>
static cudaError_t allow_max_smem(yasph_ctx* c, K kernel) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->smem_optin + 1024 - fa.sharedSizeBytes));
}
template <class Op>
static cudaError_t prepare_sweep(yasph_ctx* c) {
    return allow_max_smem(c, k_sweep<Op>);
}
// the most shared-memory-hungry tile kernels for given capacities: the list build and the WCSPH sweep (24 B per candidate)
static size_t worst_smem_bytes(uint32_t cap_dyn, uint32_t cap_stat, uint32_t cap_pc) {
    // the sweep's list staging is optional (sweep_nk_stage shrinks it to what fits), so its floor is the size without it
    const size_t a = list_smem_bytes(cap_dyn, cap_stat), b = SweepLayout<OpWcsphAccel>::total_bytes(cap_dyn, cap_stat, cap_pc, 0, 1);
    return a > b ? a : b;
}

static void free_all(yasph_ctx* c) {
    void* ptrs[] = {c->pos_adv, c->pos, c->pos_alt, c->vel, c->vel_alt, c->vstar, c->vstar_alt, c->accel, c->dens, c->alpha, c->kappa, c->stiff, c->err_buf,
                    c->f_alt0, c->f_alt1, c->keys[0], c->keys[1], c->idx[0], c->idx[1], c->cell_key, c->cell_start, c->tile_key, c->tile_pstart,
                    c->tile_cstart, c->bpos, c->bpos_alt, c->scell_key, c->scell_start, c->stile_key, c->stile_cstart, c->tile_runs, c->cslot_d,
                    c->cslot_s, c->lists, c->counts, c->tile_nk, c->apron_idx,
                    c->radix_base[0], c->scan_chunks, c->scan_total, c->scan_status, c->partials, c->ctl, c->ctl_snap, c->exp_cd, c->exp_ct, c->exp_lists,
                    c->ids, c->ids_alt, c->slab.pflag, c->slab.sel[0], c->slab.sel[1], c->slab.sel_g[0], c->slab.sel_g[1], c->slab.send_idx[0], c->slab.send_idx[1],
                    c->slab.ghost_idx[0], c->slab.ghost_idx[1], c->slab.own_idx, c->slab.sbuf[0], c->slab.sbuf[1], c->slab.rbuf[0],
                    c->slab.rbuf[1], c->slab.d_cnt};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (c->h_ctl) cudaFreeHost(c->h_ctl);
    if (c->h_pub) cudaFreeHost(c->h_pub);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->slab.h_cnt) cudaFreeHost(c->slab.h_cnt);
    if (c->slab.h_pcounts) cudaFreeHost(c->slab.h_pcounts);
    for (int sd = 0; sd < 2; ++sd) {
        if (c->slab.ev_ready[sd]) cudaEventDestroy(c->slab.ev_ready[sd]);
        if (c->slab.ev_done[sd]) cudaEventDestroy(c->slab.ev_done[sd]);
    }
    peer_teardown(c);
    comm_destroy(c);
    for (auto e : c->event_pool) cudaEventDestroy(e);
    for (auto& e : c->events) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ctl_stream) cudaStreamDestroy(c->ctl_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (int i = 0; i < 2; ++i)
        if (c->ev_early[i]) cudaEventDestroy(c->ev_early[i]);
    for (int i = 0; i < 6; ++i)
        if (c->ev_host[i]) cudaEventDestroy(c->ev_host[i]);
    if (c->ev_tables) cudaEventDestroy(c->ev_tables);
    if (c->stream) cudaStreamDestroy(c->stream);
}

extern "C" int32_t yasph_create(const yasph_config* cfg, yasph_ctx** out) {
    yasph_ctx* c = nullptr;
    if (!cfg || !out) return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_create: null argument");
    if (cfg->abi_version != YASPH_ABI_VERSION) return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_create: abi_version %u != %u", cfg->abi_version, YASPH_ABI_VERSION);
    if (!(cfg->smoothing_length > 0.f) || !(cfg->particle_density > 0.f) || !(cfg->fluid_density > 0.f))
        return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_create: smoothing_length, particle_density and fluid_density must be > 0");
    if (cfg->max_particles == 0) return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_create: max_particles must be > 0");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(c, YASPH_ERR_NO_DEVICE, "yasph_create: no CUDA device (%s); libyasph_gpu has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_create: device %d out of range (%d devices)", cfg->device, ndev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(c, YASPH_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(c, YASPH_ERR_NO_DEVICE, "yasph_create: device %d is sm_%d%d; this library contains sm_100a code only", cfg->device, prop.major, prop.minor);

    yasph_ctx* ctx = new yasph_ctx();
    c = ctx;
    c->cfg = *cfg;
    c->device = cfg->device;
    c->num_sms = prop.multiProcessorCount;
    memset(c->pass_us, 0, sizeof(c->pass_us));
#define CREATE_FAIL(code, ...)                   \
    do {                                         \
        int32_t rc_ = fail(nullptr, code, __VA_ARGS__); \
        free_all(ctx);                           \
        delete ctx;                              \
        return rc_;                              \
    } while (0)
#define CUC(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) CREATE_FAIL(YASPH_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
    CUC(cudaSetDevice(c->device));
    CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    {  // the control block is published from this stream while persistent kernels fill the SMs: highest priority
        int prio_lo = 0, prio_hi = 0;
        CUC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUC(cudaStreamCreateWithPriority(&c->ctl_stream, cudaStreamNonBlocking, prio_hi));
    }
    CUC(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    CUC(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CUC(cudaEventCreateWithFlags(&c->ev_early[i], cudaEventDisableTiming));
    for (int i = 0; i < 6; ++i) CUC(cudaEventCreate(&c->ev_host[i]));
    CUC(cudaEventCreateWithFlags(&c->ev_tables, cudaEventDisableTiming));
    if (const char* e = getenv("YASPH_DEBUG_LIST_MARGIN_PCT")) c->list_margin_pct = atoi(e);
    if (const char* e = getenv("YASPH_DEBUG_NO_PDL")) c->pdl = atoi(e) == 0;
    if (const char* e = getenv("YASPH_DEBUG_NO_FUSED_ADVECT")) c->fuse_advect = atoi(e) == 0;

    c->cap_n = cfg->max_particles;
    c->cap_m = cfg->max_boundary;
    c->max_tiles = cfg->max_tiles ? cfg->max_tiles : cfg->max_particles / 32 + 4096;
    c->smem_optin = (size_t)prop.sharedMemPerBlockOptin - 1024;  // dynamic part; 1 KB is kept for the kernels' static shared memory
    c->lim_dyn = cfg->tile_dynamic_capacity;
    c->lim_stat = cfg->tile_static_capacity;
    if (c->lim_dyn > 65535u || c->lim_stat > 65535u) CREATE_FAIL(YASPH_ERR_INVALID_ARGUMENT, "tile capacities must be <= 65535 slots (16-bit slot indices)");
    if ((c->lim_dyn || c->lim_stat) && worst_smem_bytes(c->lim_dyn, c->lim_stat, 0) > c->smem_optin)
        CREATE_FAIL(YASPH_ERR_CAPACITY, "tile capacities %u/%u need %zu bytes of shared memory per CTA, the device allows %zu", c->lim_dyn, c->lim_stat,
                    worst_smem_bytes(c->lim_dyn, c->lim_stat, 0), c->smem_optin);
    if (c->cfg.speculative_iterations == 0) c->cfg.speculative_iterations = 2;
    c->cfg.max_tiles = c->max_tiles;
    if (c->cfg.max_halo == 0) c->cfg.max_halo = cfg->max_particles / 8 > 65536u ? cfg->max_particles / 8 : 65536u;
    if (c->cfg.max_halo > cfg->max_particles) c->cfg.max_halo = cfg->max_particles;

    // ConstantFluidProperties
    c->mass = cfg->fluid_density / cfg->particle_density;   // fluidparticleworld.rs:74-76
    c->radius = 0.5f / sqrtf(cfg->particle_density);        // fluidparticleworld.rs:82-85
    c->grid.radius = cfg->smoothing_length;                 // neighborhood_search.rs:474
    c->grid.radius_sq = cfg->smoothing_length * cfg->smoothing_length;  // neighborhood_search.rs:331
    c->grid.cell_size_inv = 1.0f / cfg->smoothing_length;   // neighborhood_search.rs:475
    c->grid.grid_min = make_float2(cfg->grid_min[0], cfg->grid_min[1]);
    c->kc = make_kernel_consts(cfg->smoothing_length);
    c->tp.adaptive = cfg->adaptive_timestep;
    c->tp.fixed_ns = cfg->timestep_fixed_ns;
    c->tp.min_ns = cfg->timestep_min_ns;
    c->tp.max_ns = cfg->timestep_max_ns;
    c->tp.target_ns = cfg->timestep_target_frame_ns;
    c->tp.cfl_factor = cfg->cfl_factor;

    const size_t N = c->cap_n, M = c->cap_m, NM = N > M ? N : M;
    CUC(dmalloc(&c->pos, N));
    CUC(dmalloc(&c->pos_alt, N));
    CUC(dmalloc(&c->pos_adv, N));
    CUC(dmalloc(&c->vel, N));
    CUC(dmalloc(&c->vel_alt, N));
    CUC(dmalloc(&c->vstar, N));
    CUC(dmalloc(&c->vstar_alt, N));
    CUC(dmalloc(&c->accel, N));
    CUC(dmalloc(&c->dens, N));
    CUC(dmalloc(&c->alpha, N));
    CUC(dmalloc(&c->kappa, N));
    CUC(dmalloc(&c->stiff, N));
    CUC(dmalloc(&c->err_buf, N));
    CUC(dmalloc(&c->f_alt0, N));
    CUC(dmalloc(&c->f_alt1, N));
    for (int b = 0; b < 2; ++b) {
        CUC(dmalloc(&c->keys[b], NM));
        CUC(dmalloc(&c->idx[b], NM));
    }
    CUC(dmalloc(&c->cell_key, N + 1));
    CUC(dmalloc(&c->cell_start, N + 2));
    CUC(dmalloc(&c->tile_key, (size_t)c->max_tiles + 1));
    CUC(dmalloc(&c->tile_pstart, (size_t)c->max_tiles + 2));
    CUC(dmalloc(&c->tile_cstart, (size_t)c->max_tiles + 2));
    CUC(dmalloc(&c->bpos, M));
    CUC(dmalloc(&c->bpos_alt, M));
    CUC(dmalloc(&c->scell_key, M + 1));
    CUC(dmalloc(&c->scell_start, M + 2));
    CUC(dmalloc(&c->stile_key, M + 1));
    CUC(dmalloc(&c->stile_cstart, M + 2));
    CUC(dmalloc(&c->tile_runs, (size_t)c->max_tiles));
    CUC(dmalloc(&c->cslot_d, (size_t)c->max_tiles * REGION_CELLS));
    CUC(dmalloc(&c->cslot_s, (size_t)c->max_tiles * REGION_CELLS));
    CUC(dmalloc(&c->lists, N * LIST_WORDS));
    CUC(dmalloc(&c->counts, N));
    CUC(dmalloc(&c->tile_nk, (size_t)c->max_tiles + 1));
    CUC(dmalloc(&c->apron_idx, ((size_t)c->max_tiles + 1) * APRON_TABLE));
    CUC(dmalloc(&c->radix_base[0], 2 * radix_scratch_words((uint32_t)NM)));
    c->radix_base[1] = c->radix_base[0] + radix_scratch_words((uint32_t)NM);
    c->radix_scratch = c->radix_base[0];
    CUC(dmalloc(&c->scan_chunks, (size_t)scan_num_chunks((uint32_t)NM) + 1));
    CUC(dmalloc(&c->scan_total, 1));
    CUC(dmalloc(&c->scan_status, (size_t)scan_num_chunks((uint32_t)NM) + 3));
    CUC(cudaMemset(c->scan_status, 0, ((size_t)scan_num_chunks((uint32_t)NM) + 3) * sizeof(unsigned long long)));
    c->scan_status_clean = true;  // and its users leave it zeroed
    CUC(dmalloc(&c->ctl, 1));
    CUC(dmalloc(&c->ctl_snap, 1));
    CUC(cudaMallocHost((void**)&c->h_ctl, sizeof(Control)));
    CUC(cudaHostAlloc((void**)&c->h_pub, sizeof(yasph_ctx::Published), cudaHostAllocMapped));
    memset(c->h_pub, 0, sizeof(yasph_ctx::Published));
    CUC(cudaHostGetDevicePointer((void**)&c->d_pub, c->h_pub, 0));
    CUC(cudaMemsetAsync(c->ctl, 0, sizeof(Control), c->stream));
    CUC(cudaMemsetAsync(c->accel, 0, N * sizeof(float2), c->stream));  // WCSPHSolver: accellerations start at zero (wscsph.rs:128)
    CUC(cudaMemsetAsync(c->scell_key, 0xFF, sizeof(uint32_t), c->stream));  // empty static grid: sentinel only
    CUC(cudaMemsetAsync(c->scell_start, 0, 2 * sizeof(uint32_t), c->stream));
    CUC(cudaMemsetAsync(c->stile_cstart, 0, 2 * sizeof(uint32_t), c->stream));

    // shared-memory opt-in for every tile kernel and the persistent grid size
    CUC((prepare_sweep<OpDensityAlpha<0, true>>(c)));
    CUC((prepare_sweep<OpDensityAlpha<0, false>>(c)));
    CUC((prepare_sweep<OpDensityAlpha<1, false>>(c)));
    CUC((prepare_sweep<OpDensityAlpha<1, false, true>>(c)));
    CUC((prepare_sweep<OpDensityAlpha<2, false>>(c)));
    CUC((prepare_sweep<OpDensityAlpha<3, false>>(c)));
    CUC((prepare_sweep<OpAlphaOnly>(c)));
    CUC((prepare_sweep<OpDensityAlphaDiv>(c)));
    CUC((prepare_sweep<OpViscosity>(c)));
    CUC((prepare_sweep<OpJacobiA<0>>(c)));
    CUC((prepare_sweep<OpJacobiA<1>>(c)));
    CUC((prepare_sweep<OpJacobiB<0, false>>(c)));
    CUC((prepare_sweep<OpJacobiBAdvect>(c)));
    CUC((prepare_sweep<OpJacobiB<0, true>>(c)));
    CUC((prepare_sweep<OpJacobiB<1, false>>(c)));
    CUC((prepare_sweep<OpJacobiB<1, true>>(c)));
    CUC((prepare_sweep<OpWcsphAccel>(c)));
    CUC(allow_max_smem(c, k_build_lists<false>));
    CUC(allow_max_smem(c, k_build_lists<true>));
    CUC(allow_max_smem(c, k_radix_pass<RS_ITEMS_CHOICES[0]>));
    CUC(allow_max_smem(c, k_radix_pass<RS_ITEMS_CHOICES[1]>));
    CUC(dmalloc(&c->partials, (size_t)c->max_tiles + 1 + UNSTAGED_GRID_MAX));  // per-CTA partials of k_sweep, then of k_sweep_unstaged
#if defined(YASPH_SWEEP_TIMING) || defined(YASPH_LIST_TIMING)
    CUC(dmalloc(&c->sweep_dbg, 16));
    CUC(cudaMemset(c->sweep_dbg, 0, 128));
    CUC(cudaMemset(c->sweep_dbg + 8, 0xFF, 8));
#endif

    // TimeManager::new: initial step = timestep_min / fixed (timemanager.rs:106-109); DFSPHSolver::new iteration counts (dfsph.rs:51,55)
    memset(c->h_ctl, 0, sizeof(Control));
    c->h_ctl->step_ns = cfg->adaptive_timestep ? cfg->timestep_min_ns : cfg->timestep_fixed_ns;
    c->h_ctl->iters[0] = 1;
    c->h_ctl->iters[1] = 0;
    CUC(cudaMemcpyAsync(c->ctl, c->h_ctl, sizeof(Control), cudaMemcpyHostToDevice, c->stream));
    CUC(cudaStreamSynchronize(c->stream));
    *out = ctx;
    return YASPH_OK;
#undef CUC
#undef CREATE_FAIL
}

extern "C" int32_t yasph_destroy(yasph_ctx* c) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_all(c);
    delete c;
    return YASPH_OK;
}
extern "C" const char* yasph_last_error(const yasph_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }
extern "C" int32_t yasph_get_config(const yasph_ctx* c, yasph_config* out) {
    if (!c || !out) return YASPH_ERR_INVALID_ARGUMENT;
    *out = c->cfg;
    return YASPH_OK;
}
extern "C" int32_t yasph_set_flags(yasph_ctx* c, uint32_t flags) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    c->cfg.flags = flags;
    return YASPH_OK;
}
extern "C" int32_t yasph_get_properties(const yasph_ctx* c, float* out2) {
    if (!c || !out2) return YASPH_ERR_INVALID_ARGUMENT;
    out2[0] = c->mass;
    out2[1] = c->radius;
    return YASPH_OK;
}
extern "C" int32_t yasph_num_particles(const yasph_ctx* c, uint32_t* n, uint32_t* m) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    if (n) *n = c->slab.active ? c->slab.n_own : c->n;
    if (m) *m = c->m;
    return YASPH_OK;
}
extern "C" int32_t yasph_launch_count(const yasph_ctx* c, uint64_t* launches) {
    if (!c || !launches) return YASPH_ERR_INVALID_ARGUMENT;
    *launches = c->launches;
    return YASPH_OK;
}
extern "C" int32_t yasph_stream(const yasph_ctx* c, void** stream) {
    if (!c || !stream) return YASPH_ERR_INVALID_ARGUMENT;
    *stream = (void*)c->stream;
    return YASPH_OK;
}
extern "C" int32_t yasph_pass_times(yasph_ctx* c, float* out_us) {
    if (!c || !out_us) return YASPH_ERR_INVALID_ARGUMENT;
    memcpy(out_us, c->pass_us, sizeof(c->pass_us));
    return YASPH_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// building blocks
// ---------------------------------------------------------------------------------------------------------------------
static inline uint32_t blocks_for(uint32_t n, uint32_t threads) { return (n + threads - 1) / threads; }

// Launch of a kernel of the step's chain.  With programmatic stream serialization the kernel may become resident while its
// predecessor in the stream drains; it must (and every kernel launched through here does) call pdl_enter() before it touches
// global memory (common.cuh).
template <class... P, class... A>
static inline void launch_chain(const yasph_ctx* c, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = c->pdl ? 1u : 0u;
    cudaLaunchKernelEx(&cfg, kernel, P(std::forward<A>(args))...);  // errors surface through CHECK_LAUNCH (cudaGetLastError)
}

// Stable LSD radix sort of (keys[0], idx[0]) over n elements; result back in buffer 0 (4 passes).  radix_prepare must be
// enqueued BEFORE the kernel that generates the keys, because that kernel accumulates the digit histograms.
static uint32_t radix_bound(const yasph_ctx* c, uint32_t n) {
    if (c->slab.active) n = (uint32_t)std::min<uint64_t>(c->cap_n, (uint64_t)n + 4ull * c->slab.max_halo);
    return n;
}
static int32_t radix_prepare(yasph_ctx* c, uint32_t n) {
    // Slab mode: migrants and ghosts are appended between key generation and the sort (slab_exchange_particles), so the sort runs
    // over up to n + 4 * max_halo pairs and its status area [pass][tiles(n_sort)][bins] is larger than tiles(n) suggests: zero
    // the area of the largest count the exchange can produce (the layout is addressed with the tile count of the sort's launch).
    n = radix_bound(c, n);
    c->radix_cur ^= 1;
    c->radix_scratch = c->radix_base[c->radix_cur];
    if (n && c->radix_clean_n[c->radix_cur] < n) CU(cudaMemsetAsync(c->radix_scratch, 0, radix_scratch_words(n) * sizeof(uint32_t), c->stream));
    c->radix_clean_n[c->radix_cur] = 0u;  // about to be used
    return YASPH_OK;
}
// zeroes the area of the NEXT sort in `stream` (the caller orders that stream ahead of the next key generation)
static int32_t radix_clean_next(yasph_ctx* c, uint32_t n, cudaStream_t stream) {
    const int nxt = c->radix_cur ^ 1;
    n = radix_bound(c, n);
    // a few tiles of margin: the particle count of the next sort may differ slightly (slab mode: migration)
    n = (uint32_t)std::min<uint64_t>(std::max(c->cap_n, c->cap_m), (uint64_t)n + 4ull * RS_TILE);
    if (n == 0) return YASPH_OK;
    CU(cudaMemsetAsync(c->radix_base[nxt], 0, radix_scratch_words(n) * sizeof(uint32_t), stream));
    c->radix_clean_n[nxt] = n;
    return YASPH_OK;
}
template <int ITEMS>
static int32_t radix_sort_with(yasph_ctx* c, uint32_t n) {
    const uint32_t ntiles = (n + RS_THREADS * ITEMS - 1) / (RS_THREADS * ITEMS);
    int src = 0;
    for (int pass = 0; pass < RS_PASSES; ++pass) {
        // the first pass reads no index array: the key-generating kernels do not write one, index i belongs to key i
        launch_chain(c, k_radix_pass<ITEMS>, ntiles, RS_THREADS, sizeof(RadixPassSmem<ITEMS>), c->stream, c->keys[src], pass == 0 ? (uint32_t*)nullptr : c->idx[src],
                     c->keys[src ^ 1], c->idx[src ^ 1], n, pass, c->radix_scratch, ntiles);
        CHECK_LAUNCH();
        src ^= 1;
    }
    return YASPH_OK;
}
static int32_t radix_sort(yasph_ctx* c, uint32_t n) {
    if (n == 0) return YASPH_OK;
    // the smallest tile shape whose tiles are all resident at once (two CTAs per SM); the default when none is (sort.cuh)
    const uint64_t wave = 2ull * (uint64_t)c->num_sms * RS_THREADS;
    if (n > wave * RS_ITEMS_CHOICES[0] && n <= wave * RS_ITEMS_CHOICES[1]) return radix_sort_with<RS_ITEMS_CHOICES[1]>(c, n);
    return radix_sort_with<RS_ITEMS_CHOICES[0]>(c, n);
}

// cells (+ tiles for the dynamic grid) from the sorted keys in keys[0]
static int32_t build_cells(yasph_ctx* c, uint32_t n, bool is_static) {
    uint32_t* ck = is_static ? c->scell_key : c->cell_key;
    uint32_t* cs = is_static ? c->scell_start : c->cell_start;
    const FinishCells fin{n, ck, cs, c->tile_pstart, is_static ? c->stile_cstart : c->tile_cstart, is_static ? c->cap_m : c->max_tiles, c->ctl, is_static ? 1 : 0};
    if (n) {
        // head flags -> prefix -> compaction -> sentinels and counts in ONE kernel (scan.cuh: k_scan_fused)
        const uint32_t nch = scan_num_chunks(n);
        HeadFlagsIn in{c->keys[0]};
        HeadCompactOut out{c->keys[0], ck, cs, is_static ? c->stile_key : c->tile_key, is_static ? nullptr : c->tile_pstart,
                           is_static ? c->stile_cstart : c->tile_cstart, is_static ? c->cap_m : c->max_tiles};
        // k_scan_fused and k_select4_fused leave the status words zeroed (the chunk that finishes last cleans up)
        if (!c->scan_status_clean) CU(cudaMemsetAsync(c->scan_status, 0, ((size_t)nch + 2) * sizeof(unsigned long long), c->stream));
        c->scan_status_clean = true;
        launch_chain(c, k_scan_fused<HeadFlagsIn, HeadCompactOut, FinishCells>, nch, SCAN_THREADS, 0, c->stream, in, n, c->scan_status, out, fin);
        CHECK_LAUNCH();
    } else {
        k_finish_cells<<<1, 32, 0, c->stream>>>(fin);
        CHECK_LAUNCH();
    }
    return YASPH_OK;
}

static TileTables tile_tables(const yasph_ctx* c) { return TileTables{c->tile_runs, c->cslot_d, c->cslot_s}; }

static SweepCommon sweep_common(const yasph_ctx* c) {
    SweepCommon s;
    s.tt = tile_tables(c);
    s.lists = c->lists;
    s.counts = c->counts;
    s.tile_nk = c->tile_nk;
    s.apron_idx = c->apron_idx;
    s.cap_pc = c->cap_pc;
    s.nk_stage = 0;  // launch_sweep
    s.pos = c->pos;
    s.bpos = c->bpos;
    s.ctl = c->ctl;
    s.kc = c->kc;
    s.cap_dyn = c->cap_dyn;
    s.cap_stat = c->cap_stat;
    s.n = c->n;
    s.mass = c->mass;
    s.rho0 = c->cfg.fluid_density;
    s.partials = c->partials;
    s.partials_unstaged = c->partials + c->max_tiles + 1;
    s.n_partials_unstaged = 0;
    s.ghost = c->slab.active ? c->slab.pflag : nullptr;
#ifdef YASPH_SWEEP_TIMING
    s.dbg = c->sweep_dbg;
#endif
    s.n_avg = c->slab.active ? (float)c->slab.n_global : (float)c->n;
    return s;
}
// persistent grid of a tile kernel: every SM filled to the kernel's occupancy at this shared-memory size, at most one CTA per tile
template <class K>
static uint32_t persistent_grid(const yasph_ctx* c, K kernel, size_t smem_bytes, int threads = TILE_THREADS) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem_bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
    const uint32_t g = (uint32_t)(c->num_sms * per_sm);
    return g < c->num_tiles ? g : (c->num_tiles ? c->num_tiles : 1u);
}
template <class Op>
static int32_t launch_sweep(yasph_ctx* c, Op op) {
    if (c->num_tiles == 0) return YASPH_OK;
    SweepCommon sc = sweep_common(c);
    if (c->unstaged_tiles) {
        // tiles too large to stage go first, from global memory; their partial reductions are picked up by k_sweep's last CTA
        const uint32_t gu = std::min<uint32_t>(std::min<uint32_t>(c->num_tiles, (uint32_t)c->num_sms * 4u), UNSTAGED_GRID_MAX);
        k_sweep_unstaged<Op><<<gu, SWU_THREADS, 0, c->stream>>>(sc, op);
        CHECK_LAUNCH();
        sc.n_partials_unstaged = gu;
    }
    // List words staged per particle: what the lists needed at the last read-back (+1: they change slowly), bounded by what
    // lets three CTAs share an SM.  Any value is correct -- words beyond it are read from global memory.
    const uint32_t hint = c->h_ctl->max_nk ? c->h_ctl->max_nk + 1u : 4u;
    const size_t soft = std::min<size_t>(c->smem_optin, YASPH_SWEEP_MIN_CTAS == 3 ? 74 * 1024 : (size_t)(227 * 1024) / YASPH_SWEEP_MIN_CTAS - 1536);
    typedef SweepLayout<Op> L;
    sc.nstages = SW_STAGES;
    if (L::total_bytes(c->cap_dyn, c->cap_stat, c->cap_pc, 0, SW_STAGES) <= soft) {
        sc.nk_stage = sweep_nk_stage<Op>(c->cap_dyn, c->cap_stat, c->cap_pc, hint, SW_STAGES, soft);
    } else {  // very large tiles: the deepest ring that fits the SM at all (neighborhood_update checked that one stage does)
        while (sc.nstages > 1 && L::total_bytes(c->cap_dyn, c->cap_stat, c->cap_pc, 0, sc.nstages) > c->smem_optin) --sc.nstages;
        sc.nk_stage = sweep_nk_stage<Op>(c->cap_dyn, c->cap_stat, c->cap_pc, hint, sc.nstages, c->smem_optin);
    }
    const size_t bytes = L::total_bytes(c->cap_dyn, c->cap_stat, c->cap_pc, sc.nk_stage, sc.nstages);
    launch_chain(c, k_sweep<Op>, persistent_grid(c, k_sweep<Op>, bytes, SW_THREADS), SW_THREADS, bytes, c->stream, sc, op);
    CHECK_LAUNCH();
    return YASPH_OK;
}

// Copies the control block to the host and waits for it (everything enqueued before it has then completed).  The block
// travels as zero-copy stores of a one-warp kernel into mapped host memory, published by a sequence number the host polls:
// a few microseconds less per read-back than a DMA copy plus a stream synchronisation, four or more times per step.
__global__ void k_publish_control(const Control* __restrict__ ctl, uint32_t* __restrict__ dst_words, unsigned int* seq_out, unsigned int seq) {
    pdl_enter();
    const uint32_t* src = reinterpret_cast<const uint32_t*>(ctl);
    for (uint32_t q = threadIdx.x; q < sizeof(Control) / 4; q += blockDim.x) dst_words[q] = src[q];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *reinterpret_cast<volatile unsigned int*>(seq_out) = seq;
        __threadfence_system();
    }
}
// polls a sequence number a kernel writes into mapped host memory
// `pub_stream`: the stream the publishing kernel was launched on (the main stream, or ctl_stream while the main stream is busy)
static int32_t wait_published(yasph_ctx* c, volatile unsigned int* vs, unsigned int seq, cudaStream_t pub_stream) {
    for (uint32_t spins = 0; *vs != seq; ++spins) {
        if ((spins & 0xFFFu) == 0xFFFu) {  // now and then: has the publishing stream died (or drained without publishing)?
            const cudaError_t q = cudaStreamQuery(pub_stream);
            if (q != cudaErrorNotReady) {
                if (q != cudaSuccess) return fail(c, YASPH_ERR_CUDA, "%s while waiting for a device-published value", cudaGetErrorString(q));
                std::atomic_thread_fence(std::memory_order_acquire);
                if (*vs != seq) return fail(c, YASPH_ERR_CUDA, "device-published value did not arrive");
            }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return YASPH_OK;
}
// the two halves of read_control: enqueue the snapshot / wait for it
static int32_t publish_control(yasph_ctx* c, cudaStream_t stream, unsigned int* seq_out, const Control* src = nullptr) {
    static_assert(sizeof(Control) % 4 == 0, "Control is published word by word");
    const unsigned int seq = ++c->pub_seq;
    launch_chain(c, k_publish_control, 1, 64, 0, stream, src ? src : c->ctl, reinterpret_cast<uint32_t*>(&c->d_pub->ctl), &c->d_pub->seq, seq);
    CHECK_LAUNCH();
    *seq_out = seq;
    return YASPH_OK;
}
static int32_t await_control(yasph_ctx* c, cudaStream_t stream, unsigned int seq) {
    TRY(wait_published(c, &c->h_pub->seq, seq, stream));
    memcpy(c->h_ctl, const_cast<const Control*>(&c->h_pub->ctl), sizeof(Control));
    return YASPH_OK;
}
static int32_t read_control(yasph_ctx* c, cudaStream_t stream = nullptr) {
    unsigned int seq = 0;
    TRY(publish_control(c, stream ? stream : c->stream, &seq));
    return await_control(c, stream ? stream : c->stream, seq);
}
static int32_t check_capacity_flags(yasph_ctx* c) {
    if (c->h_ctl->err_comm)
        return fail(c, YASPH_ERR_COMM, "peer-memory transport: rank %d waited more than %.0f s for a neighbour (mask %u)", c->slab.rank, PEER_TIMEOUT_NS * 1e-9,
                    c->h_ctl->err_comm);
    if (c->h_ctl->err_tile_count)
        return fail(c, YASPH_ERR_CAPACITY, "%u non-empty tiles exceed max_tiles=%u (particles too sparse for the configured capacity)", c->h_ctl->err_tile_count, c->max_tiles);
    if (c->h_ctl->err_tile_capacity)
        return fail(c, YASPH_ERR_CAPACITY, "a tile stages %u candidates; slots are 16-bit (<= 65535)", c->h_ctl->err_tile_capacity);
    return YASPH_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// slab mode (multi-GPU) plumbing
// ---------------------------------------------------------------------------------------------------------------------
static SlabParams slab_params(const yasph_ctx* c) { return SlabParams{c->slab.col_lo, c->slab.col_hi, c->slab.pflag, c->slab.active ? 1 : 0}; }
static bool has_left(const yasph_ctx* c) { return c->slab.active && c->slab.rank > 0; }
static bool has_right(const yasph_ctx* c) { return c->slab.active && c->slab.rank + 1 < c->slab.world; }
constexpr size_t RECORD_MAX_BYTES = 3 * 8 + 3 * 4;

// ordered selection of two index lists over [0, n): *total_out (device) receives count A | count B << 32
template <class In>
static int32_t select_pair(yasph_ctx* c, In in, uint32_t n, uint32_t* out_a, uint32_t* out_b, uint8_t* pflag_out, uint32_t cap, unsigned long long* total_out) {
    if (n == 0) {
        CU(cudaMemsetAsync(total_out, 0, sizeof(unsigned long long), c->stream));
        return YASPH_OK;
    }
    const uint32_t nch = scan_num_chunks(n);
    IndexPairOut out{out_a, out_b, pflag_out, cap};
    k_scan_reduce<unsigned long long, In><<<nch, SCAN_THREADS, 0, c->stream>>>(in, n, c->scan_chunks);
    CHECK_LAUNCH();
    k_scan_chunks<unsigned long long><<<1, SCAN_THREADS, 0, c->stream>>>(c->scan_chunks, nch, total_out);
    CHECK_LAUNCH();
    k_scan_apply<unsigned long long, In, IndexPairOut><<<nch, SCAN_THREADS, 0, c->stream>>>(in, n, c->scan_chunks, out);
    CHECK_LAUNCH();
    return YASPH_OK;
}

// one NCCL group: send `sbytes[side]` bytes of sbuf[side] to / receive `rbytes[side]` bytes into rbuf[side] from the left
// (side 0) and the right (side 1) neighbour rank
static int32_t loopback_sendrecv(yasph_ctx* c, const void* const sbuf[2], const size_t sbytes[2], void* const rbuf[2], const size_t rbytes[2]) {
    auto& sl = c->slab;
    Fabric* f = sl.fabric;
    // like the NCCL transport, a link with nothing to move in either direction is skipped (both ends know both sizes)
    const bool side_on[2] = {sl.rank > 0 && (sbytes[0] || rbytes[0]), sl.rank + 1 < sl.world && (sbytes[1] || rbytes[1])};
    const int peer[2] = {sl.rank - 1, sl.rank + 1};
    uint64_t seq[2] = {0, 0};
    // offer my buffers
    for (int sd = 0; sd < 2; ++sd)
        if (side_on[sd]) {
            seq[sd] = ++sl.link_seq[sd];
            CU(cudaEventRecord(sl.ev_ready[sd], c->stream));
        }
    {
        std::lock_guard<std::mutex> lk(f->m);
        for (int sd = 0; sd < 2; ++sd)
            if (side_on[sd]) f->post[sl.rank * 2 + sd] = Fabric::Post{sbuf[sd], sbytes[sd], sl.ev_ready[sd], seq[sd]};
    }
    f->cv.notify_all();
    // take what the neighbours offer (the neighbour on my side sd offers on its side 1 - sd)
    for (int sd = 0; sd < 2; ++sd) {
        if (!side_on[sd]) continue;
        Fabric::Post p;
        bool failed;
        {
            std::unique_lock<std::mutex> lk(f->m);
            f->cv.wait(lk, [&] { return f->failed || f->post[peer[sd] * 2 + (1 - sd)].seq >= seq[sd]; });
            failed = f->failed;
            p = f->post[peer[sd] * 2 + (1 - sd)];
        }
        if (failed) return fail(c, YASPH_ERR_COMM, "loopback transport: a peer rank failed");
        if (p.seq != seq[sd] || p.bytes != rbytes[sd])
            return fail(c, YASPH_ERR_COMM, "loopback transport: rank %d expected %zu bytes (seq %llu) from rank %d, which offers %zu (seq %llu)", sl.rank,
                        rbytes[sd], (unsigned long long)seq[sd], peer[sd], p.bytes, (unsigned long long)p.seq);
        CU(cudaStreamWaitEvent(c->stream, p.ready, 0));
        if (p.bytes) CU(cudaMemcpyAsync(rbuf[sd], p.buf, p.bytes, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaEventRecord(sl.ev_done[sd], c->stream));
        {
            std::lock_guard<std::mutex> lk(f->m);
            f->ack[sl.rank * 2 + sd] = Fabric::Ack{sl.ev_done[sd], seq[sd]};
        }
        f->cv.notify_all();
    }
    // my send buffers may be rewritten once the neighbours' copies have run
    for (int sd = 0; sd < 2; ++sd) {
        if (!side_on[sd]) continue;
        Fabric::Ack a;
        bool failed;
        {
            std::unique_lock<std::mutex> lk(f->m);
            f->cv.wait(lk, [&] { return f->failed || f->ack[peer[sd] * 2 + (1 - sd)].seq >= seq[sd]; });
            failed = f->failed;
            a = f->ack[peer[sd] * 2 + (1 - sd)];
        }
        if (failed) return fail(c, YASPH_ERR_COMM, "loopback transport: a peer rank failed");
        CU(cudaStreamWaitEvent(c->stream, a.done, 0));
    }
    sl.halo_exchanges++;
    return YASPH_OK;
}
static int32_t loopback_allreduce(yasph_ctx* c, void* dev_ptr, bool is_double_sum) {
    auto& sl = c->slab;
    Fabric* f = sl.fabric;
    double v = 0.0;
    if (is_double_sum) {
        CU(cudaMemcpyAsync(&v, dev_ptr, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    } else {
        float x = 0.f;
        CU(cudaMemcpyAsync(&x, dev_ptr, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        v = (double)x;
    }
    double result;
    bool peer_failed;
    {
        std::unique_lock<std::mutex> lk(f->m);
        const uint64_t gen = f->ar_gen;
        f->ar_val[sl.rank] = v;
        if (++f->ar_count == f->world) {
            double r = is_double_sum ? 0.0 : f->ar_val[0];
            for (int k = 0; k < f->world; ++k) r = is_double_sum ? r + f->ar_val[k] : (f->ar_val[k] > r ? f->ar_val[k] : r);
            f->ar_result = r;
            f->ar_count = 0;
            f->ar_gen++;
            f->cv.notify_all();
        } else {
            f->cv.wait(lk, [&] { return f->failed || f->ar_gen != gen; });
        }
        result = f->ar_result;
        peer_failed = f->failed;
    }
    if (peer_failed) return fail(c, YASPH_ERR_COMM, "loopback transport: a peer rank failed");
    if (is_double_sum) {
        CU(cudaMemcpyAsync(dev_ptr, &result, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    } else {
        const float x = (float)result;
        CU(cudaMemcpyAsync(dev_ptr, &x, sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));  // the source is a stack variable
    sl.allreduces++;
    return YASPH_OK;
}

static int32_t neighbor_sendrecv(yasph_ctx* c, const void* const sbuf[2], const size_t sbytes[2], void* const rbuf[2], const size_t rbytes[2]) {
    if (c->slab.fabric) return loopback_sendrecv(c, sbuf, sbytes, rbuf, rbytes);
    ncclComm_t comm = (ncclComm_t)c->slab.comm;
    const bool side_on[2] = {has_left(c), has_right(c)};
    const int peer[2] = {c->slab.rank - 1, c->slab.rank + 1};
    NC(g_nccl.GroupStart());
    for (int sd = 0; sd < 2; ++sd) {
        if (!side_on[sd]) continue;
        if (sbytes[sd]) NC(g_nccl.Send(sbuf[sd], sbytes[sd], ncclChar, peer[sd], comm, c->stream));
        if (rbytes[sd]) NC(g_nccl.Recv(rbuf[sd], rbytes[sd], ncclChar, peer[sd], comm, c->stream));
    }
    NC(g_nccl.GroupEnd());
    c->slab.halo_exchanges++;
    return YASPH_OK;
}

// Per-pass halo lists of the current structure, in sorted order: what this rank sends (its first / last W owned columns) and where
// its ghosts sit.  Entry k of a send list is entry k of the neighbour's ghost list (slab.cuh "ordering contract").  Built on demand:
// with ghost layers a pass rarely needs an exchange (yasph_ctx::Slab::valid).
static int32_t ensure_halo_lists(yasph_ctx* c) {
    auto& sl = c->slab;
    if (sl.halo_lists_valid) return YASPH_OK;
    const uint32_t W = sl.ghost_cols, n = c->n;
    const uint32_t a_lo = sl.col_lo, a_hi = has_left(c) ? std::min(sl.col_lo + W, sl.col_hi) : sl.col_lo;
    const uint32_t b_hi = sl.col_hi, b_lo = has_right(c) ? (sl.col_hi - sl.col_lo > W ? sl.col_hi - W : sl.col_lo) : sl.col_hi;
    TRY(select_pair(c, ColumnSelIn{c->keys[0], nullptr, a_lo, a_hi, b_lo, b_hi}, n, sl.send_idx[0], sl.send_idx[1], nullptr, sl.max_halo, &c->ctl->slab_send));
    TRY(select_pair(c, GhostSelIn{c->keys[0], sl.col_lo, sl.col_hi}, n, sl.ghost_idx[0], sl.ghost_idx[1], nullptr, sl.max_halo, &c->ctl->slab_own));
    unsigned long long cnt[2] = {0, 0};
    CU(cudaMemcpyAsync(&cnt[0], &c->ctl->slab_send, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(&cnt[1], &c->ctl->slab_own, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    sl.n_send[0] = (uint32_t)(cnt[0] & 0xFFFFFFFFull);
    sl.n_send[1] = (uint32_t)(cnt[0] >> 32);
    if ((uint32_t)(cnt[1] & 0xFFFFFFFFull) != sl.n_ghost[0] || (uint32_t)(cnt[1] >> 32) != sl.n_ghost[1])
        return fail(c, YASPH_ERR_STATE, "rank %d: ghost lists (%llu | %llu) disagree with the exchanged ghosts (%u | %u)", sl.rank, cnt[1] & 0xFFFFFFFFull, cnt[1] >> 32,
                    sl.n_ghost[0], sl.n_ghost[1]);
    if (sl.n_send[0] > sl.max_halo || sl.n_send[1] > sl.max_halo)
        return fail(c, YASPH_ERR_CAPACITY, "rank %d: %u | %u particles in the halo send lists exceed max_halo=%u", sl.rank, sl.n_send[0], sl.n_send[1], sl.max_halo);
    sl.halo_lists_valid = true;
    sl.own_idx_valid = false;  // the selection above borrowed Control::slab_own
    return YASPH_OK;
}

// halo exchange of one per-particle field of the current structure: owners' values overwrite the ghosts'
template <typename T>
static int32_t halo_exchange(yasph_ctx* c, T* field) {
    if (!c->slab.active || c->slab.world < 2) return YASPH_OK;
    TRY(ensure_halo_lists(c));
    pass_begin(c, YASPH_PASS_HALO);
    auto& sl = c->slab;
    const uint32_t ns = sl.n_send[0] + sl.n_send[1], ng = sl.n_ghost[0] + sl.n_ghost[1];
    if (sl.peer) {
        // stores into the neighbours' mailboxes, then the scatter of what the neighbours stored here (slab.cuh)
        const uint64_t seq = ++sl.halo_seq;
        const unsigned par = (unsigned)(seq & 1u);
        const bool hl = has_left(c), hr = has_right(c);
        void* bl = hl ? sl.peer_box[sl.rank - 1] : nullptr;
        void* br = hr ? sl.peer_box[sl.rank + 1] : nullptr;
        // my data arrives on the RIGHT side of my left neighbour and on the LEFT side of my right neighbour
        PeerBoxHeader* me = reinterpret_cast<PeerBoxHeader*>(sl.box);
        const uint32_t nsl = hl ? sl.n_send[0] : 0u, nsr = hr ? sl.n_send[1] : 0u, ngl = hl ? sl.n_ghost[0] : 0u, ngr = hr ? sl.n_ghost[1] : 0u;
        const uint32_t work = std::max(nsl + nsr, ngl + ngr);
        const uint32_t grid = std::max(1u, std::min(blocks_for(work, 256), (uint32_t)c->num_sms));  // all CTAs resident (k_halo_exchange)
        k_halo_exchange<T><<<grid, 256, 0, c->stream>>>(
            field, sl.send_idx[0], nsl, sl.send_idx[1], nsr, hl ? reinterpret_cast<T*>(peer_payload(bl, sl.max_halo, par, 1)) : nullptr,
            hr ? reinterpret_cast<T*>(peer_payload(br, sl.max_halo, par, 0)) : nullptr, hl ? &reinterpret_cast<PeerBoxHeader*>(bl)->halo_flag[1] : nullptr,
            hr ? &reinterpret_cast<PeerBoxHeader*>(br)->halo_flag[0] : nullptr, sl.ghost_idx[0], ngl, sl.ghost_idx[1], ngr,
            reinterpret_cast<const T*>(peer_payload(sl.box, sl.max_halo, par, 0)), reinterpret_cast<const T*>(peer_payload(sl.box, sl.max_halo, par, 1)),
            hl ? &me->halo_flag[0] : nullptr, hr ? &me->halo_flag[1] : nullptr, seq, sl.d_ticket, c->ctl);
        CHECK_LAUNCH();
        sl.halo_exchanges++;
        pass_end(c);
        return YASPH_OK;
    }
    if (ns) {
        k_halo_pack<T><<<blocks_for(ns, 256), 256, 0, c->stream>>>(field, sl.send_idx[0], sl.n_send[0], sl.send_idx[1], sl.n_send[1],
                                                                   reinterpret_cast<T*>(sl.sbuf[0]), reinterpret_cast<T*>(sl.sbuf[1]));
        CHECK_LAUNCH();
    }
    const void* sb[2] = {sl.sbuf[0], sl.sbuf[1]};
    void* rb[2] = {sl.rbuf[0], sl.rbuf[1]};
    const size_t sby[2] = {sl.n_send[0] * sizeof(T), sl.n_send[1] * sizeof(T)};
    const size_t rby[2] = {sl.n_ghost[0] * sizeof(T), sl.n_ghost[1] * sizeof(T)};
    TRY(neighbor_sendrecv(c, sb, sby, rb, rby));
    if (ng) {
        k_halo_unpack<T><<<blocks_for(ng, 256), 256, 0, c->stream>>>(field, sl.ghost_idx[0], sl.n_ghost[0], sl.ghost_idx[1], sl.n_ghost[1],
                                                                     reinterpret_cast<const T*>(sl.rbuf[0]), reinterpret_cast<const T*>(sl.rbuf[1]));
        CHECK_LAUNCH();
    }
    pass_end(c);
    return YASPH_OK;
}

// Slab mode bookkeeping around a pass (see yasph_ctx::Slab::valid).  `gathered`: fields the pass reads from NEIGHBOURS (they must be valid in
// the first ghost column: refreshed by a halo exchange if not); `own`: fields it reads of the particle itself; returns the validity of
// what the pass writes.  Elementwise passes have no gathered inputs.
template <typename T>
static int32_t slab_refresh(yasph_ctx* c, SlabField f, T* array) {
    if (!c->slab.active || c->slab.world < 2) return YASPH_OK;
    if (c->slab.valid[f] >= 1) return YASPH_OK;
    TRY(halo_exchange(c, array));
    c->slab.valid[f] = (int)c->slab.ghost_cols;
    return YASPH_OK;
}
static int slab_out_valid(const yasph_ctx* c, std::initializer_list<SlabField> gathered, std::initializer_list<SlabField> own) {
    int v = std::min((int)c->slab.ghost_cols, c->slab.valid[SF_POS]);  // positions are gathered by every neighbour pass (and bound the lists themselves)
    for (SlabField f : gathered) v = std::min(v, c->slab.valid[f]);
    v -= 1;
    for (SlabField f : own) v = std::min(v, c->slab.valid[f]);
    return std::max(v, 0);
}
static int slab_own_valid(const yasph_ctx* c, std::initializer_list<SlabField> own) {
    int v = (int)c->slab.ghost_cols;
    for (SlabField f : own) v = std::min(v, c->slab.valid[f]);
    return v;
}

// in-place all-reduce of one scalar of the control block over the ranks
static int32_t allreduce_scalar(yasph_ctx* c, void* dev_ptr, ncclDataType_t type, ncclRedOp_t op) {
    if (!c->slab.active || c->slab.world < 2) return YASPH_OK;
    if (c->slab.fabric) return loopback_allreduce(c, dev_ptr, type == ncclDouble && op == ncclSum);
    if (c->slab.peer) {
        launch_chain(c, k_allreduce_peer<NoAfter>, 1, 32, 0, c->stream, dev_ptr, type == ncclDouble && op == ncclSum ? 1 : 0, c->slab.d_boxes, c->slab.rank, c->slab.world,
                     ++c->slab.ar_seq, c->ctl, NoAfter());
        CHECK_LAUNCH();
        c->slab.allreduces++;
        return YASPH_OK;
    }
    NC(g_nccl.AllReduce(dev_ptr, dev_ptr, 1, type, op, (ncclComm_t)c->slab.comm, c->stream));
    c->slab.allreduces++;
    return YASPH_OK;
}

// CompactMortonCellGrid::update for the dynamic particles + NeighborLists::update.
// `keys_ready`: keys[0]/idx[0] were already produced by a fused advect / kick kernel.
// gather2: float2 arrays permuted with the particles (pointer to the ctx member pair), gather1 likewise for 4-byte arrays.
struct GatherPlan {
    float2** a2[3];
    float2** alt2[3];
    int n2 = 0;
    float** a1[3];
    float** alt1[3];
    int n1 = 0;
};
static RecordArrays record_arrays(const GatherPlan& gp) {
    RecordArrays r;
    memset(&r, 0, sizeof(r));
    r.n2 = gp.n2;
    r.n1 = gp.n1;
    for (int q = 0; q < gp.n2; ++q) r.a2[q] = *gp.a2[q];
    for (int q = 0; q < gp.n1; ++q) r.a1[q] = *gp.a1[q];
    return r;
}

// Slab mode, between key generation and the sort: particles that left the slab migrate to the adjacent rank, the ghosts of the
// previous structure are dropped, and the first / last W owned columns are sent as the neighbours' new ghost layers -- ONE ordered
// four-way selection, ONE message per neighbour, one wait of the host (slab.cuh).  On return *n_sort particles (old local set +
// arrivals) carry sort keys (dropped ones YASPH_KEY_DROPPED) and *n_keep of them survive the sort.
static int32_t slab_exchange_particles(yasph_ctx* c, const GatherPlan& gp, uint32_t n_old, uint32_t* n_sort, uint32_t* n_keep) {
    auto& sl = c->slab;
    const RecordArrays ra = record_arrays(gp);
    const size_t rec = record_bytes(gp.n2, gp.n1);
    const SlabParams sp = slab_params(c);
    const uint32_t ghosts_old = sl.n_ghost[0] + sl.n_ghost[1];
    const bool hl = has_left(c), hr = has_right(c);
    SlabCounts* dcnt = reinterpret_cast<SlabCounts*>(c->ctl->slab_cnt);
    static_assert(sizeof(SlabCounts) == sizeof(((Control*)nullptr)->slab_cnt), "SlabCounts lives in Control::slab_cnt");
    // (1) the four lists: migrants left | right, stayers of the first | last W owned columns (W = ghost_cols; an empty range disables a side)
    const uint32_t W = sl.ghost_cols;
    const uint32_t a_lo = sl.col_lo, a_hi = hl ? std::min(sl.col_lo + W, sl.col_hi) : sl.col_lo;
    const uint32_t b_hi = sl.col_hi, b_lo = hr ? (sl.col_hi - sl.col_lo > W ? sl.col_hi - W : sl.col_lo) : sl.col_hi;
    if (n_old) {
        const uint32_t nch = scan_num_chunks(n_old);
        const Select4In in{c->keys[0], sl.pflag, a_lo, a_hi, b_lo, b_hi};
        if (!c->scan_status_clean) CU(cudaMemsetAsync(c->scan_status, 0, ((size_t)nch + 2) * sizeof(unsigned long long), c->stream));
        c->scan_status_clean = true;  // both users of the status words leave them zeroed
        launch_chain(c, k_select4_fused, nch, SCAN_THREADS, 0, c->stream, in, n_old, c->scan_status, sl.sel[0], sl.sel[1], sl.sel_g[0], sl.sel_g[1], sl.max_halo, c->ctl->slab_sel4);
        CHECK_LAUNCH();
    } else {
        CU(cudaMemsetAsync(c->ctl->slab_sel4, 0, 4 * sizeof(uint32_t), c->stream));
    }
    // (2) the exchange
    const uint32_t hseq = ++sl.pcounts_seq;
    if (sl.peer) {
        const uint64_t seq = ++sl.halo_seq;
        const unsigned par = (unsigned)(seq & 1u);
        void* bl = hl ? sl.peer_box[sl.rank - 1] : nullptr;
        void* br = hr ? sl.peer_box[sl.rank + 1] : nullptr;
        PeerBoxHeader* me = reinterpret_cast<PeerBoxHeader*>(sl.box);
        RecordExchangeArgs xa;
        xa.arr = ra;
        xa.idx_ml = sl.sel[0], xa.idx_mr = sl.sel[1], xa.idx_gl = sl.sel_g[0], xa.idx_gr = sl.sel_g[1];
        xa.counts4 = c->ctl->slab_sel4;
        xa.cap_halo = sl.max_halo, xa.cap_n = c->cap_n, xa.n_old = n_old;
        xa.dst_l = hl ? peer_payload(bl, sl.max_halo, par, 1) : nullptr;
        xa.dst_r = hr ? peer_payload(br, sl.max_halo, par, 0) : nullptr;
        xa.pflag_l = hl ? &reinterpret_cast<PeerBoxHeader*>(bl)->halo_flag[1] : nullptr;
        xa.pflag_r = hr ? &reinterpret_cast<PeerBoxHeader*>(br)->halo_flag[0] : nullptr;
        xa.src_l = peer_payload(sl.box, sl.max_halo, par, 0);
        xa.src_r = peer_payload(sl.box, sl.max_halo, par, 1);
        xa.flag_l = hl ? &me->halo_flag[0] : nullptr;
        xa.flag_r = hr ? &me->halo_flag[1] : nullptr;
        xa.seq = seq, xa.ticket = sl.d_ticket, xa.pflag = sl.pflag, xa.dc = dcnt, xa.ctl = c->ctl, xa.g = c->grid;
        xa.col_lo = sl.col_lo, xa.col_hi = sl.col_hi, xa.W = W;
        xa.host_counts = sl.d_pcounts, xa.host_seq = hseq;
        launch_chain(c, k_records_exchange, 64, 256, 0, c->stream, xa);  // grid-stride (the counts are known on the device only); all CTAs resident
        CHECK_LAUNCH();
    } else {
        // host-mediated transports (NCCL send / recv, loopback): counts first, then the records
        uint32_t mine[4];
        CU(cudaMemcpyAsync(mine, c->ctl->slab_sel4, sizeof(mine), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        uint32_t* d = sl.d_cnt;  // [0..1] to the left: migrants, ghosts; [2..3] to the right; [4..5] from the left; [6..7] from the right
        const uint32_t snd[4] = {hl ? mine[0] : 0u, hl ? mine[2] : 0u, hr ? mine[1] : 0u, hr ? mine[3] : 0u};
        memcpy(sl.h_cnt, snd, sizeof(snd));
        CU(cudaMemcpyAsync(d, sl.h_cnt, 4 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemsetAsync(d + 4, 0, 4 * sizeof(uint32_t), c->stream));
        {
            const void* sb[2] = {d, d + 2};
            void* rb[2] = {d + 4, d + 6};
            const size_t by[2] = {2 * sizeof(uint32_t), 2 * sizeof(uint32_t)};
            TRY(neighbor_sendrecv(c, sb, by, rb, by));
        }
        CU(cudaMemcpyAsync(sl.h_cnt + 4, d + 4, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const uint32_t in_m[2] = {sl.h_cnt[4], sl.h_cnt[6]}, in_g[2] = {sl.h_cnt[5], sl.h_cnt[7]};
        const uint32_t out_m[2] = {snd[0], snd[2]}, out_g[2] = {snd[1], snd[3]};
        for (int sd = 0; sd < 2; ++sd)
            if (out_m[sd] + out_g[sd] > sl.max_halo || in_m[sd] + in_g[sd] > sl.max_halo)
                return fail(c, YASPH_ERR_CAPACITY, "rank %d: %u + %u particles out / %u + %u in on side %d exceed max_halo=%u", sl.rank, out_m[sd], out_g[sd], in_m[sd],
                            in_g[sd], sd, sl.max_halo);
        if ((uint64_t)n_old + in_m[0] + in_m[1] + in_g[0] + in_g[1] > c->cap_n)
            return fail(c, YASPH_ERR_CAPACITY, "rank %d: %u local particles with arrivals > max_particles=%u", sl.rank, n_old + in_m[0] + in_m[1] + in_g[0] + in_g[1], c->cap_n);
        for (int sd = 0; sd < 2; ++sd)
            if (out_m[sd] + out_g[sd]) {
                k_pack_records<<<blocks_for(out_m[sd] + out_g[sd], 256), 256, 0, c->stream>>>(ra, sl.sel[sd], out_m[sd], sl.sel_g[sd], out_g[sd], sl.sbuf[sd]);
                CHECK_LAUNCH();
            }
        {
            const void* sb[2] = {sl.sbuf[0], sl.sbuf[1]};
            void* rb[2] = {sl.rbuf[0], sl.rbuf[1]};
            const size_t sby[2] = {(out_m[0] + out_g[0]) * rec, (out_m[1] + out_g[1]) * rec}, rby[2] = {(in_m[0] + in_g[0]) * rec, (in_m[1] + in_g[1]) * rec};
            if (sby[0] + sby[1] + rby[0] + rby[1]) TRY(neighbor_sendrecv(c, sb, sby, rb, rby));
        }
        // the same layout as k_records_pull: migrants from the left | right, then ghosts from the left | right
        uint32_t first = n_old;
        for (int sd = 0; sd < 2; ++sd)
            if (in_m[sd]) {
                k_unpack_records<<<blocks_for(in_m[sd], 256), 256, 0, c->stream>>>(ra, first, in_m[sd], sl.rbuf[sd], in_m[sd] + in_g[sd], 0u, sl.pflag);
                CHECK_LAUNCH();
                first += in_m[sd];
            }
        for (int sd = 0; sd < 2; ++sd)
            if (in_g[sd]) {
                k_unpack_records<<<blocks_for(in_g[sd], 256), 256, 0, c->stream>>>(ra, first, in_g[sd], sl.rbuf[sd], in_m[sd] + in_g[sd], in_m[sd], sl.pflag);
                CHECK_LAUNCH();
                first += in_g[sd];
            }
        const uint32_t all[10] = {mine[0], mine[1], mine[2], mine[3], in_m[0], in_m[1], in_g[0], in_g[1], 0u, 0u};  // SlabCounts: out_m, out_g, in_m, in_g
        memcpy(sl.h_cnt + 8, all, sizeof(all));
        CU(cudaMemcpyAsync(dcnt, sl.h_cnt + 8, sizeof(all), cudaMemcpyHostToDevice, c->stream));
        // (3) this rank's out-migrants that are now part of a neighbour's first W columns stay here as ghosts; the counts reach the host
        // (the peer transport does this at the end of k_records_exchange)
        k_slab_retain<<<1, 32, 0, c->stream>>>(ra, c->grid, sl.sel[0], sl.sel[1], dcnt, n_old, c->cap_n, sl.max_halo, sl.col_lo, sl.col_hi, W, sl.pflag, sl.d_pcounts, hseq);
        CHECK_LAUNCH();
    }
    TRY(wait_published(c, &sl.h_pcounts->seq, hseq, c->stream));
    const SlabCounts cnt = const_cast<const PeerCounts*>(sl.h_pcounts)->cnt;
    // a particle that leaves through an end of the domain has no rank to go to
    if ((cnt.out_m[0] && !hl) || (cnt.out_m[1] && !hr))
        return fail(c, YASPH_ERR_STATE, "rank %d: %u / %u particles left the domain's first / last slab [%u, %u)", sl.rank, cnt.out_m[0], cnt.out_m[1], sl.col_lo, sl.col_hi);
    for (int sd = 0; sd < 2; ++sd)
        if (cnt.out_m[sd] + cnt.out_g[sd] > sl.max_halo || cnt.in_m[sd] + cnt.in_g[sd] > sl.max_halo)
            return fail(c, YASPH_ERR_CAPACITY, "rank %d: %u + %u particles out / %u + %u in on side %d exceed max_halo=%u", sl.rank, cnt.out_m[sd], cnt.out_g[sd],
                        cnt.in_m[sd], cnt.in_g[sd], sd, sl.max_halo);
    const uint32_t n_mig_in = cnt.in_m[0] + cnt.in_m[1], n_ghost_in = cnt.in_g[0] + cnt.in_g[1] + cnt.retained[0] + cnt.retained[1];
    const uint64_t n2 = (uint64_t)n_old + n_mig_in + n_ghost_in;
    if (n2 > c->cap_n) return fail(c, YASPH_ERR_CAPACITY, "rank %d: %llu local particles with arrivals and ghosts > max_particles=%u", sl.rank, (unsigned long long)n2, c->cap_n);
    // (4) sort keys of the arrivals: migrants are classified (they must lie inside the slab: particles move less than a cell per step and
    // a slab is many cells wide); ghosts keep their keys, they lie outside the slab by construction
    if (n_mig_in + n_ghost_in) {
        launch_chain(c, k_keygen, keygen_grid(n_mig_in + n_ghost_in, c->num_sms), KG_THREADS, 0, c->stream, c->pos, n_old, (uint32_t)n2, c->grid, c->keys[0], c->idx[0],
                     c->radix_scratch, sp, n_old + n_mig_in);
        CHECK_LAUNCH();
    }
    if (n_mig_in) {
        launch_chain(c, k_check_arrivals, blocks_for(n_mig_in, 256), 256, 0, c->stream, sl.pflag, n_old, n_old + n_mig_in, c->ctl);
        CHECK_LAUNCH();
    }
    sl.mig_out[0] = cnt.out_m[0];
    sl.mig_out[1] = cnt.out_m[1];
    sl.mig_in = n_mig_in;
    sl.n_ghost[0] = cnt.in_g[0] + cnt.retained[0];
    sl.n_ghost[1] = cnt.in_g[1] + cnt.retained[1];
    sl.halo_exchanges++;
    *n_sort = (uint32_t)n2;
    *n_keep = (uint32_t)n2 - ghosts_old - sl.mig_out[0] - sl.mig_out[1];
    return YASPH_OK;
}

// yasph_step_host hands results to the caller's (pinned) arrays as soon as they are final instead of after the step:
//   positions   once the gather has produced the sorted order -- enqueued right AFTER the neighbourhood update's read-back,
//               because a read-back issued while a large download is in flight waits for that download to drain (measured:
//               the zero-copy publish, like a small DMA copy, completes only after the bulk copy ahead of it on the link);
//   densities   after the density sweep, behind the positions on the copy stream;
//   velocities  on the main stream right after the last Jacobi kernel launched, BEFORE the solve's read-back, so the copy
//               runs while the host waits for the control block; if the solve then turns out to need more iterations the
//               copy is stale and yasph_step_host repeats it.
// A download is marked where its array becomes final (event on the main stream) but SUBMITTED only after the step's remaining
// kernels have been enqueued (submit_downloads): kernel launches issued while a bulk download occupies the link were
// measured to start ~20 us late each.
static int32_t early_download(yasph_ctx* c, float** host_dst, const void* dev, size_t bytes, int which) {
    if (!*host_dst) return YASPH_OK;
    CU(cudaEventRecord(c->ev_early[which], c->stream));
    c->pending[which].host = *host_dst;
    c->pending[which].dev = dev;
    c->pending[which].bytes = bytes;
    *host_dst = nullptr;  // done for this call
    return YASPH_OK;
}
static int32_t submit_downloads(yasph_ctx* c) {
    for (int which = 0; which < 2; ++which) {
        auto& p = c->pending[which];
        if (!p.host) continue;
        CU(cudaStreamWaitEvent(c->copy_stream, c->ev_early[which], 0));
        CU(cudaMemcpyAsync(p.host, p.dev, p.bytes, cudaMemcpyDeviceToHost, c->copy_stream));
        if (c->cfg.flags & YASPH_FLAG_PROFILE_PASSES) CU(cudaEventRecord(c->ev_host[2 + which], c->copy_stream));
        p.host = nullptr;
    }
    return YASPH_OK;
}
static int32_t ensure_own_index(yasph_ctx* c, bool no_wait);
// Slab mode: the caller gets the OWNED particles only -- they are compacted through `scratch` first (the order of the last download).
template <typename T>
static int32_t early_download_owned(yasph_ctx* c, float** host_dst, const T* dev, T* scratch, int which) {
    if (!*host_dst) return YASPH_OK;
    auto& sl = c->slab;
    if (!sl.active) return early_download(c, host_dst, dev, (size_t)c->n * sizeof(T), which);
    if (sl.n_ghost[0] + sl.n_ghost[1] == 0) return early_download(c, host_dst, dev, (size_t)sl.n_own * sizeof(T), which);
    TRY(ensure_own_index(c, true));
    if (sl.n_own) {
        k_own_compact<T><<<blocks_for(sl.n_own, 256), 256, 0, c->stream>>>(dev, sl.own_idx, sl.n_own, scratch);
        CHECK_LAUNCH();
    }
    return early_download(c, host_dst, scratch, (size_t)sl.n_own * sizeof(T), which);
}
static int32_t early_velocities(yasph_ctx* c, const float2* final_vel) {
    if (!c->early_vel_out) return YASPH_OK;
    if (c->cfg.flags & YASPH_FLAG_PROFILE_PASSES) CU(cudaEventRecord(c->ev_host[4], c->stream));
    CU(cudaMemcpyAsync(c->early_vel_out, final_vel, (size_t)c->n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    if (c->cfg.flags & YASPH_FLAG_PROFILE_PASSES) CU(cudaEventRecord(c->ev_host[5], c->stream));
    c->early_vel_stale = false;
    return YASPH_OK;
}

static int32_t neighborhood_update(yasph_ctx* c, bool keys_ready, GatherPlan& gp, bool positions_final = false, bool sorted_ready = false) {
    uint32_t n = c->n;
    c->lists_valid = false;
    c->slab.own_idx_valid = false;
    if (c->cfg.flags & YASPH_FLAG_TRACK_IDS) {  // ids travel with the particles
        if (gp.n1 >= 3) return fail(c, YASPH_ERR_STATE, "neighborhood_update: too many permuted arrays");
        gp.a1[gp.n1] = reinterpret_cast<float**>(&c->ids);
        gp.alt1[gp.n1] = reinterpret_cast<float**>(&c->ids_alt);
        gp.n1++;
    }
    pass_begin(c, YASPH_PASS_SORT);
    if (!keys_ready && n) {
        TRY(radix_prepare(c, n));
        k_keygen<<<keygen_grid(n, c->num_sms), KG_THREADS, 0, c->stream>>>(c->pos, 0u, n, c->grid, c->keys[0], c->idx[0], c->radix_scratch, slab_params(c), n);
        CHECK_LAUNCH();
    }
    uint32_t n_sort = n;
    if (c->slab.active) {
        pass_end(c);
        pass_begin(c, YASPH_PASS_MIGRATE);
        // the ghost records carry exactly the permuted arrays: those are fresh on all W ghost columns afterwards, everything else is stale
        for (int f = 0; f < SLAB_NUM_FIELDS; ++f) c->slab.valid[f] = 0;
        auto mark = [&](const void* arr) {
            const SlabField f = arr == c->pos ? SF_POS : arr == c->vel ? SF_VEL : arr == c->vstar ? SF_VSTAR : arr == c->kappa ? SF_KAPPA : arr == c->stiff ? SF_STIFF : SLAB_NUM_FIELDS;
            if (f != SLAB_NUM_FIELDS) c->slab.valid[f] = (int)c->slab.ghost_cols;
        };
        for (int q = 0; q < gp.n2; ++q) mark(*gp.a2[q]);
        for (int q = 0; q < gp.n1; ++q) mark(*gp.a1[q]);
        TRY(slab_exchange_particles(c, gp, n, &n_sort, &n));
        pass_end(c);
        pass_begin(c, YASPH_PASS_SORT);
        c->n = n;
    }
    if (!sorted_ready) TRY(radix_sort(c, n_sort));
    pass_end(c);
    pass_begin(c, YASPH_PASS_GATHER);
    bool gather_aside = false;
    if (n) {
        GatherArgs ga;
        memset(&ga, 0, sizeof(ga));
        ga.n2 = gp.n2;
        ga.n1 = gp.n1;
        for (int q = 0; q < gp.n2; ++q) {
            ga.in2[q] = *gp.a2[q];
            ga.out2[q] = *gp.alt2[q];
        }
        for (int q = 0; q < gp.n1; ++q) {
            ga.in1[q] = *gp.a1[q];
            ga.out1[q] = *gp.alt1[q];
        }
        // The gather (permuted arrays) and the cell / tile tables (sorted keys only) do not depend on each other: the gather runs on a
        // side stream next to them and is joined before the list build.  Not under YASPH_FLAG_PROFILE_PASSES (per-pass times).
        gather_aside = !(c->cfg.flags & YASPH_FLAG_PROFILE_PASSES) || g_no_pass_events;
        cudaStream_t gs = gather_aside ? c->aux_stream : c->stream;
        if (gather_aside) {
            CU(cudaEventRecord(c->ev_fork, c->stream));
            CU(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
        }
        if (c->slab.active) {
            CU(cudaMemsetAsync(&c->ctl->slab_ghost, 0, sizeof(unsigned long long), gs));
            ga.keys = c->keys[0];
            ga.ghost = c->slab.pflag;
            ga.ghost_count = &c->ctl->slab_ghost;
            ga.col_lo = c->slab.col_lo;
            ga.col_hi = c->slab.col_hi;
        }
        k_gather<<<blocks_for(n, 256), 256, 0, gs>>>(c->idx[0], n, ga);
        CHECK_LAUNCH();
        TRY(radix_clean_next(c, n, gs));  // the scratch of the next step's sort, off the main stream's kernel chain
        if (gather_aside) CU(cudaEventRecord(c->ev_join, c->aux_stream));
        for (int q = 0; q < gp.n2; ++q) std::swap(*gp.a2[q], *gp.alt2[q]);
        for (int q = 0; q < gp.n1; ++q) std::swap(*gp.a1[q], *gp.alt1[q]);
    }
    pass_end(c);
    pass_begin(c, YASPH_PASS_CELLS_TILES);
    TRY(build_cells(c, n, false));  // also zeroes the list statistics (FinishCells)
    if (n) {
        TileTableArgs ta{c->tile_key, c->tile_pstart, c->tile_cstart, c->cell_key, c->cell_start, c->stile_key, c->stile_cstart,
                         c->scell_key, c->scell_start, c->tile_runs, c->cslot_d, c->cslot_s};
        launch_chain(c, k_tile_tables, c->num_sms * (64 / TT_WARPS), TT_WARPS * 32, 0, c->stream, ta, c->ctl);
        CHECK_LAUNCH();
    }
    c->slab.halo_lists_valid = false;  // per-pass halo lists of the new structure are built when a pass needs an exchange (ensure_halo_lists)
    if (gather_aside) CU(cudaStreamWaitEvent(c->stream, c->ev_join, 0));  // the permuted arrays are in place from here on
    pass_end(c);
    // The one host round trip of the neighbourhood update: tile count (grid of every tile kernel until the next update) and
    // the largest tile (their shared-memory size).  The list build does not wait for it when the previous structure's sizes are
    // known: it is launched right away with those sizes plus a margin (its result does not depend on the capacities, only its
    // shared-memory layout does), the control block is published from the second stream while it runs, and the sizes are
    // checked afterwards -- a structure that outgrew the margin gets its lists built again.
    // yasph_step_host: the sorted positions are final here; mark the point (the copy itself is submitted later, see early_download).
    // With downloads overlapping the rest of the step the early list build below does not pay: measured 1.79 ms per call with it,
    // 1.70 ms without (the list kernel runs 30 us longer under the download), so such calls read the sizes back first.
    const bool downloads_armed = positions_final && !c->slab.active && c->early_pos_out != nullptr;
    if (positions_final && !c->slab.active) TRY(early_download(c, &c->early_pos_out, c->pos, (size_t)n * sizeof(float2), 0));
    bool lists_launched = false;
    uint32_t spec_dyn = 0, spec_stat = 0;
    if (c->lists_valid_once && c->num_tiles && n && c->list_margin_pct > -100 && !downloads_armed && !c->unstaged_tiles) {  // margin <= -100: no early launch (A/B)
        // margin: 12.5 % + 16 slots; YASPH_DEBUG_LIST_MARGIN_PCT (read at yasph_create) overrides the percentage so that a test
        // can force the rebuild path with an undersized guess
        const int pct = c->list_margin_pct;
        spec_dyn = (uint32_t)std::max<long long>(16, ((long long)c->cap_dyn * (100 + pct) / 100 + 31) & ~15ll);
        spec_stat = (uint32_t)std::max<long long>(16, ((long long)c->cap_stat * (100 + pct) / 100 + 31) & ~15ll);
        const size_t bytes = list_smem_bytes(spec_dyn, spec_stat);
        if (bytes <= c->smem_optin) {
            CU(cudaEventRecord(c->ev_tables, c->stream));
            pass_begin(c, YASPH_PASS_LISTS);
            ListArgs la{tile_tables(c), c->pos, c->bpos, c->keys[0], c->grid, c->ctl, c->lists, c->counts, c->tile_nk, spec_dyn, spec_stat, c->apron_idx};
            launch_chain(c, k_build_lists<false>, persistent_grid(c, k_build_lists<false>, bytes, NB_THREADS), NB_THREADS, bytes, c->stream, la);
            CHECK_LAUNCH();
            pass_end(c);
            lists_launched = true;
            CU(cudaStreamWaitEvent(c->ctl_stream, c->ev_tables, 0));
        }
    }
    TRY(read_control(c, lists_launched ? c->ctl_stream : nullptr));
    TRY(check_capacity_flags(c));
    if (c->slab.active) {
        auto& sl = c->slab;
        const Control& h = *c->h_ctl;
        const uint32_t g0 = (uint32_t)(h.slab_ghost & 0xFFFFFFFFull), g1 = (uint32_t)(h.slab_ghost >> 32);
        if (h.err_slab) return fail(c, YASPH_ERR_STATE, "rank %d: %u migrants arrived outside the slab [%u, %u) (a particle crossed more than one slab in a step)", sl.rank, h.err_slab, sl.col_lo, sl.col_hi);
        if (g0 != sl.n_ghost[0] || g1 != sl.n_ghost[1])
            return fail(c, YASPH_ERR_STATE, "rank %d: %u | %u ghosts in the sorted structure, %u | %u were exchanged", sl.rank, g0, g1, sl.n_ghost[0], sl.n_ghost[1]);
        sl.n_own = n - g0 - g1;
        // yasph_step_host_slab: the sorted positions are final -- the owned ones leave on the copy stream while the step goes on
        if (positions_final) TRY(early_download_owned(c, &c->early_pos_out, c->pos, c->pos_alt, 0));
    }
    c->num_tiles = n ? c->h_ctl->num_tiles : 0u;
    // Staging capacities of the tile kernels: the largest tile, unless that exceeds the configured limits or the shared memory of
    // an SM -- then the capacities are clipped and the tiles beyond them are processed unstaged, from global memory (slow and
    // correct: the reference accepts any particle density up to its 64-neighbour cap, neighborhood_search.rs:353-381).
    c->cap_dyn = (c->h_ctl->max_dyn_total + 15u) & ~15u;
    c->cap_pc = (c->h_ctl->max_pcount + 6u + 15u) & ~15u;  // + the alignment surplus of the bulk copies
    c->cap_stat = (c->h_ctl->max_stat_total + 15u) & ~15u;
    if (c->lim_dyn && c->cap_dyn > c->lim_dyn) c->cap_dyn = std::max(16u, c->lim_dyn & ~15u);
    if (c->lim_stat && c->cap_stat > c->lim_stat) c->cap_stat = std::max(16u, c->lim_stat & ~15u);
    if (c->lim_dyn && c->cap_pc > c->cap_dyn) c->cap_pc = c->cap_dyn;  // a tile's own particles are part of its dynamic candidates
    while (worst_smem_bytes(c->cap_dyn, c->cap_stat, c->cap_pc) > c->smem_optin) {
        // shrink the largest consumer by a quarter (bytes per slot: dynamic 24, static 8, own particle ~20 + staged list words)
        const size_t bd = (size_t)c->cap_dyn * 24, bs = (size_t)c->cap_stat * 8, bp = (size_t)c->cap_pc * 20;
        uint32_t* victim = bd >= bs && bd >= bp ? &c->cap_dyn : (bs >= bp ? &c->cap_stat : &c->cap_pc);
        if (*victim <= 16u) return fail(c, YASPH_ERR_CAPACITY, "the tile kernels do not fit the device's shared memory (%zu bytes) at any capacity", c->smem_optin);
        *victim = std::max(16u, (*victim - *victim / 4u) & ~15u);
    }
    if (lists_launched && (c->h_ctl->max_dyn_total > spec_dyn || c->h_ctl->max_stat_total > spec_stat)) {
        // the structure outgrew the margin: the early launch skipped the tiles that did not fit -- build the lists again
        CU(cudaMemsetAsync(&c->ctl->total_neighbors, 0, sizeof(unsigned long long) + 2 * sizeof(unsigned int), c->stream));
        lists_launched = false;
        c->list_rebuilds++;
    }
    pass_begin(c, YASPH_PASS_LISTS);
    if (c->num_tiles && !lists_launched) {
        const size_t bytes = list_smem_bytes(c->cap_dyn, c->cap_stat);
        ListArgs la{tile_tables(c), c->pos, c->bpos, c->keys[0], c->grid, c->ctl, c->lists, c->counts, c->tile_nk, c->cap_dyn, c->cap_stat, c->apron_idx};
#ifdef YASPH_LIST_TIMING
        la.dbg = c->sweep_dbg;
#endif
        launch_chain(c, k_build_lists<false>, persistent_grid(c, k_build_lists<false>, bytes, NB_THREADS), NB_THREADS, bytes, c->stream, la);
        CHECK_LAUNCH();
        spec_dyn = c->cap_dyn;
        spec_stat = c->cap_stat;
    }
    if (c->num_tiles && (c->h_ctl->max_dyn_total > spec_dyn || c->h_ctl->max_stat_total > spec_stat)) {
        // tiles too large to stage (clipped capacities): the same kernel on exactly those tiles, candidates from global memory
        const size_t bytes = list_smem_bytes(spec_dyn, spec_stat);
        ListArgs la{tile_tables(c), c->pos, c->bpos, c->keys[0], c->grid, c->ctl, c->lists, c->counts, c->tile_nk, spec_dyn, spec_stat, c->apron_idx, 1u};
        k_build_lists<true><<<persistent_grid(c, k_build_lists<true>, bytes, NB_THREADS), NB_THREADS, bytes, c->stream>>>(la);
        CHECK_LAUNCH();
    }
    c->unstaged_tiles = c->h_ctl->max_dyn_total > c->cap_dyn || c->h_ctl->max_stat_total > c->cap_stat || ((c->h_ctl->max_pcount + 6u + 15u) & ~15u) > c->cap_pc;
    pass_end(c);
    c->lists_valid = true;
    c->lists_valid_once = true;
    return YASPH_OK;
}

static void fill_report(const yasph_ctx* c, yasph_step_report* r) {
    if (!r) return;
    const Control& h = *c->h_ctl;
    memset(r, 0, sizeof(*r));
    r->dt_prev_ns = h.step_prev_ns;
    r->dt_ns = h.step_ns;
    r->dt = h.dt;
    r->max_velocity = h.max_velocity;
    r->iters_density = h.iters[0];
    r->iters_divergence = h.iters[1];
    r->avg_density_error = h.avg[0];
    r->avg_divergence = h.avg[1];
    r->warm_density = h.warm[0];
    r->warm_divergence = h.warm[1];
    r->neighbors_capped = h.capped;
    r->neighbors_dropped = h.dropped;
    r->list_rebuilds = (uint32_t)c->list_rebuilds;  // early list builds repeated since creation (the guess of the tile capacities was too small)
    r->not_converged = h.not_converged;
    r->num_cells = h.num_cells;
    r->num_tiles = h.num_tiles;
    r->total_neighbors = h.total_neighbors;
}

// a new particle set of n particles is about to be copied into pos / vel (all owned, no ghosts, no lists)
static int32_t reset_particle_set(yasph_ctx* c, uint32_t n) {
    const uint32_t n_prev = c->slab.active ? c->slab.n_own : c->n;
    if (n != n_prev) c->dfsph_ready = false;  // dfsph.rs:419: alpha_values.len() != positions.len() -> re-initialise
    c->n = n;
    c->lists_valid = false;
    c->have_particles = true;
    if (c->slab.active) {
        auto& sl = c->slab;
        sl.n_own = n;
        sl.n_ghost[0] = sl.n_ghost[1] = sl.n_send[0] = sl.n_send[1] = 0;
        sl.own_idx_valid = false;
        for (int f = 0; f < SLAB_NUM_FIELDS; ++f) sl.valid[f] = 0;
        c->dfsph_ready = false;  // the ghosts are gone: the structure must be rebuilt
        if (n) CU(cudaMemsetAsync(sl.pflag, 0, n, c->stream));
    }
    if (c->cfg.flags & YASPH_FLAG_TRACK_IDS) {
        if (!c->ids) {
            CU(dmalloc(&c->ids, (size_t)c->cap_n));
            CU(dmalloc(&c->ids_alt, (size_t)c->cap_n));
        }
        if (n) {
            k_iota<<<blocks_for(n, 256), 256, 0, c->stream>>>(c->ids, n, c->slab.id_base);
            CHECK_LAUNCH();
        }
    }
    return YASPH_OK;
}

// slab mode: sorted positions of the owned particles (built on demand, valid until the next neighbourhood update)
// no_wait: in mid-step (yasph_step_host_slab hands results over while the step still computes) the count is not read back here;
// it travels with the step's last control block and is checked there
static int32_t ensure_own_index(yasph_ctx* c, bool no_wait) {
    auto& sl = c->slab;
    if (!sl.active || sl.own_idx_valid) return YASPH_OK;
    if (!sl.own_idx) CU(dmalloc(&sl.own_idx, (size_t)c->cap_n));
    // (the deferred count has its own word of the control block: slab_own is borrowed by the halo lists in mid-step)
    TRY(select_pair(c, OwnSelIn{sl.pflag}, c->n, sl.own_idx, sl.own_idx, nullptr, c->cap_n, no_wait ? &c->ctl->slab_migrants : &c->ctl->slab_own));
    if (no_wait) {
        sl.own_idx_valid = true;
        sl.own_count_unchecked = true;
        return YASPH_OK;
    }
    unsigned long long cnt = 0;
    CU(cudaMemcpyAsync(&cnt, &c->ctl->slab_own, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if ((uint32_t)(cnt & 0xFFFFFFFFull) != sl.n_own)
        return fail(c, YASPH_ERR_STATE, "rank %d: %llu owned particles flagged, %u expected", sl.rank, cnt & 0xFFFFFFFFull, sl.n_own);
    sl.own_idx_valid = true;
    return YASPH_OK;
}
// device -> host copy of a per-particle array: owned particles only in slab mode (compacted through `scratch`)
template <typename T>
static int32_t download_array(yasph_ctx* c, const T* src, T* scratch, void* host_out) {
    if (!c->slab.active || (c->slab.n_ghost[0] + c->slab.n_ghost[1] == 0)) {
        const size_t n = c->slab.active ? c->slab.n_own : c->n;
        if (n) CU(cudaMemcpyAsync(host_out, src, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
        return YASPH_OK;
    }
    TRY(ensure_own_index(c, false));
    const uint32_t n = c->slab.n_own;
    if (!n) return YASPH_OK;
    k_own_compact<T><<<blocks_for(n, 256), 256, 0, c->stream>>>(src, c->slab.own_idx, n, scratch);
    CHECK_LAUNCH();
    CU(cudaMemcpyAsync(host_out, scratch, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    return YASPH_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// particle state
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int32_t yasph_set_boundary(yasph_ctx* c, const float* xy, uint32_t m) {
    if (!c || (m && !xy)) return YASPH_ERR_INVALID_ARGUMENT;
    if (m > c->cap_m) return fail(c, YASPH_ERR_CAPACITY, "yasph_set_boundary: %u boundary particles > max_boundary=%u", m, c->cap_m);
    CU(cudaSetDevice(c->device));
    c->m = m;
    c->lists_valid = false;
    if (m) {
        CU(cudaMemcpyAsync(c->bpos, xy, (size_t)m * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
        // update_static (neighborhood_search.rs:488-491): sort the boundary particles in place, build the static cells
        TRY(radix_prepare(c, m));
        k_keygen<<<keygen_grid(m, c->num_sms), KG_THREADS, 0, c->stream>>>(c->bpos, 0u, m, c->grid, c->keys[0], c->idx[0], c->radix_scratch, SlabParams{0u, 0u, nullptr, 0}, 0u);
        CHECK_LAUNCH();
        TRY(radix_sort(c, m));
        GatherArgs ga;
        memset(&ga, 0, sizeof(ga));
        ga.n2 = 1;
        ga.in2[0] = c->bpos;
        ga.out2[0] = c->bpos_alt;
        k_gather<<<blocks_for(m, 256), 256, 0, c->stream>>>(c->idx[0], m, ga);
        CHECK_LAUNCH();
        std::swap(c->bpos, c->bpos_alt);
    }
    TRY(build_cells(c, m, true));
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}

extern "C" int32_t yasph_upload_particles(yasph_ctx* c, const float* pos_xy, const float* vel_xy, uint32_t n) {
    if (!c || (n && !pos_xy)) return YASPH_ERR_INVALID_ARGUMENT;
    if (n > c->cap_n) return fail(c, YASPH_ERR_CAPACITY, "yasph_upload_particles: %u particles > max_particles=%u", n, c->cap_n);
    CU(cudaSetDevice(c->device));
    TRY(reset_particle_set(c, n));
    if (n) {
        CU(cudaMemcpyAsync(c->pos, pos_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
        if (vel_xy)
            CU(cudaMemcpyAsync(c->vel, vel_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
        else
            CU(cudaMemsetAsync(c->vel, 0, (size_t)n * sizeof(float2), c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}

extern "C" int32_t yasph_download_particles(yasph_ctx* c, float* pos_xy, float* vel_xy, float* densities) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    if (!c->have_particles) return fail(c, YASPH_ERR_STATE, "yasph_download_particles: no particles uploaded");
    CU(cudaSetDevice(c->device));
    if (pos_xy) TRY(download_array(c, c->pos, c->pos_alt, pos_xy));
    if (vel_xy) TRY(download_array(c, c->vel, c->vel_alt, vel_xy));
    if (densities) TRY(download_array(c, c->dens, c->f_alt0, densities));
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}

extern "C" int32_t yasph_download_field(yasph_ctx* c, int32_t field, void* out, uint64_t out_bytes) {
    if (!c || !out) return YASPH_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(c->device));
    const bool local = (field & YASPH_FIELD_LOCAL_BIT) != 0;  // slab mode: owned + ghosts, uncompacted
    field &= ~YASPH_FIELD_LOCAL_BIT;
    const size_t n = (c->slab.active && !local) ? c->slab.n_own : c->n, m = c->m;
    const float2* src2 = nullptr;
    const float* src1 = nullptr;
    switch (field) {
        case YASPH_FIELD_GHOST: {
            if (!local) return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_download_field: YASPH_FIELD_GHOST needs YASPH_FIELD_LOCAL_BIT");
            if (out_bytes < c->n) return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_download_field: buffer too small");
            if (c->slab.active) {
                if (c->n) CU(cudaMemcpyAsync(out, c->slab.pflag, c->n, cudaMemcpyDeviceToHost, c->stream));
                CU(cudaStreamSynchronize(c->stream));
            } else {
                memset(out, 0, c->n);
            }
            return YASPH_OK;
        }
        case YASPH_FIELD_POSITION: src2 = c->pos; break;
        case YASPH_FIELD_VELOCITY: src2 = c->vel; break;
        case YASPH_FIELD_ACCELERATION: src2 = c->accel; break;
        case YASPH_FIELD_DENSITY: src1 = c->dens; break;
        case YASPH_FIELD_ALPHA: src1 = c->alpha; break;
        case YASPH_FIELD_KAPPA: src1 = c->kappa; break;
        case YASPH_FIELD_STIFFNESS: src1 = c->stiff; break;
        case YASPH_FIELD_CELL_KEY: src1 = reinterpret_cast<const float*>(c->keys[0]); break;
        case YASPH_FIELD_SORT_PERMUTATION: src1 = reinterpret_cast<const float*>(c->idx[0]); break;
        case YASPH_FIELD_ID:
            if (!c->ids) return fail(c, YASPH_ERR_STATE, "yasph_download_field: ids are not tracked (YASPH_FLAG_TRACK_IDS)");
            src1 = reinterpret_cast<const float*>(c->ids);
            break;
        case YASPH_FIELD_BOUNDARY: {
            if (out_bytes < m * sizeof(float2)) return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_download_field: buffer too small");
            if (m) CU(cudaMemcpyAsync(out, c->bpos, m * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            return YASPH_OK;
        }
        default: return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_download_field: unknown field %d", field);
    }
    const size_t bytes = n * (src2 ? sizeof(float2) : sizeof(float));
    if (out_bytes < bytes) return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_download_field: buffer of %llu bytes < %zu", (unsigned long long)out_bytes, bytes);
    if (local) {
        if (bytes) CU(cudaMemcpyAsync(out, src2 ? (const void*)src2 : (const void*)src1, bytes, cudaMemcpyDeviceToHost, c->stream));
    } else if (src2) {
        TRY(download_array(c, src2, c->pos_alt, out));
    } else {
        TRY(download_array(c, src1, c->f_alt0, out));
    }
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// TimeManager mirror
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int32_t yasph_time_get_step_ns(const yasph_ctx* cc, uint64_t* step_ns) {
    yasph_ctx* c = const_cast<yasph_ctx*>(cc);
    if (!c || !step_ns) return YASPH_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(c->device));
    TRY(read_control(c));
    *step_ns = c->h_ctl->step_ns;
    return YASPH_OK;
}
extern "C" int32_t yasph_time_set_step_ns(yasph_ctx* c, uint64_t step_ns) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(c->device));
    unsigned long long v = step_ns;
    CU(cudaMemcpyAsync(&c->ctl->step_ns, &v, sizeof(v), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}
extern "C" int32_t yasph_time_set_total_simulated_ns(yasph_ctx* c, uint64_t total_ns) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(c->device));
    struct {
        unsigned long long total;
        unsigned int is_current, pad;
    } v = {total_ns, 1u, 0u};
    static_assert(offsetof(Control, total_is_current) == offsetof(Control, total_simulated_ns) + 8, "the pair is written with one copy");
    CU(cudaMemcpyAsync(&c->ctl->total_simulated_ns, &v, 12, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}
extern "C" int32_t yasph_time_get_total_simulated_ns(const yasph_ctx* cc, uint64_t* total_ns) {
    yasph_ctx* c = const_cast<yasph_ctx*>(cc);
    if (!c || !total_ns) return YASPH_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(c->device));
    TRY(read_control(c));
    *total_ns = c->h_ctl->total_simulated_ns;
    return YASPH_OK;
}
extern "C" int32_t yasph_time_restart(yasph_ctx* c) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(&c->ctl->total_simulated_ns, 0, 12, c->stream));  // timemanager.rs:131-133: total_simulated_time = 0
    return yasph_time_set_step_ns(c, c->cfg.adaptive_timestep ? c->cfg.timestep_min_ns : c->cfg.timestep_fixed_ns);
}

// ---------------------------------------------------------------------------------------------------------------------
// neighbourhood-only surface
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int32_t yasph_neighborhood_update(yasph_ctx* c, yasph_step_report* report) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    if (!c->have_particles) return fail(c, YASPH_ERR_STATE, "yasph_neighborhood_update: no particles uploaded");
    CU(cudaSetDevice(c->device));
    // update_neighborhood_datastructure(vec![], vec![]): positions and velocities are re-sorted (fluidparticleworld.rs:242-243)
    GatherPlan gp;
    gp.n2 = 2;
    gp.a2[0] = &c->pos;
    gp.alt2[0] = &c->pos_alt;
    gp.a2[1] = &c->vel;
    gp.alt2[1] = &c->vel_alt;
    TRY(neighborhood_update(c, false, gp));
    TRY(read_control(c));
    pass_resolve(c);
    TRY(check_capacity_flags(c));
    fill_report(c, report);
    return YASPH_OK;
}

extern "C" int32_t yasph_neighbors_download(yasph_ctx* c, uint16_t* count_dynamic, uint16_t* count_total, uint32_t* lists64) {
    if (!c || !count_dynamic || !count_total) return YASPH_ERR_INVALID_ARGUMENT;
    if (!c->lists_valid) return fail(c, YASPH_ERR_STATE, "yasph_neighbors_download: neighbour lists are not built");
    CU(cudaSetDevice(c->device));
    const size_t n = c->n;
    if (!n) return YASPH_OK;
    if (!c->exp_cd) {
        CU(dmalloc(&c->exp_cd, (size_t)c->cap_n));
        CU(dmalloc(&c->exp_ct, (size_t)c->cap_n));
    }
    if (lists64 && !c->exp_lists) CU(dmalloc(&c->exp_lists, (size_t)c->cap_n * YASPH_MAXN));
    k_export_lists<<<c->num_tiles ? c->num_tiles : 1, 256, 0, c->stream>>>(tile_tables(c), c->ctl, c->lists, c->counts, c->exp_cd, c->exp_ct, lists64 ? c->exp_lists : nullptr);
    CHECK_LAUNCH();
    CU(cudaMemcpyAsync(count_dynamic, c->exp_cd, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(count_total, c->exp_ct, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
    if (lists64) CU(cudaMemcpyAsync(lists64, c->exp_lists, n * YASPH_MAXN * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}

template <int KERNEL, bool WITH_PRESSURE = false>
static int32_t launch_density(yasph_ctx* c) {
    OpDensityAlpha<KERNEL, false, WITH_PRESSURE> op;
    op.dens = c->dens;
    op.alpha = nullptr;
    op.rho_p = c->vstar;  // WCSPH: (rho, Tait pressure) per particle, in the buffer DFSPH uses for the predicted velocities
    op.stiffness = c->cfg.wcsph_stiffness;
    return launch_sweep(c, op);
}
extern "C" int32_t yasph_update_densities(yasph_ctx* c, int32_t kernel) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    if (!c->lists_valid) return fail(c, YASPH_ERR_STATE, "yasph_update_densities: neighbour lists are not built");
    CU(cudaSetDevice(c->device));
    if (c->n) {
        switch (kernel) {
            case YASPH_KERNEL_WENDLAND_C2: TRY(launch_density<0>(c)); break;
            case YASPH_KERNEL_POLY6: TRY(launch_density<1>(c)); break;
            case YASPH_KERNEL_SPIKY: TRY(launch_density<2>(c)); break;
            case YASPH_KERNEL_CUBIC: TRY(launch_density<3>(c)); break;
            default: return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_update_densities: unknown kernel %d", kernel);
        }
    }
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}
extern "C" int32_t yasph_compute_alpha(yasph_ctx* c) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    if (!c->lists_valid) return fail(c, YASPH_ERR_STATE, "yasph_compute_alpha: neighbour lists are not built");
    CU(cudaSetDevice(c->device));
    if (c->n) {
        OpAlphaOnly op;
        op.dens = nullptr;
        op.alpha = c->alpha;
        op.rho_p = nullptr;
        op.stiffness = 0.f;
        TRY(launch_sweep(c, op));
    }
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// solvers
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int32_t yasph_clear_cached(yasph_ctx* c) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(c->device));
    // DFSPH: dfsph.rs:406-412 (alpha / warm-start arrays dropped, iteration counts 0); WCSPH: wscsph.rs:122-124
    c->dfsph_ready = false;
    c->dfsph_n = 0;
    unsigned int zero2[2] = {0u, 0u};
    CU(cudaMemcpyAsync(&c->ctl->iters[0], zero2, sizeof(zero2), cudaMemcpyHostToDevice, c->stream));
    c->h_ctl->iters[0] = c->h_ctl->iters[1] = 0u;
    CU(cudaMemsetAsync(c->accel, 0, (size_t)c->cap_n * sizeof(float2), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}

static ViscParams visc_params(const yasph_ctx* c) {
    ViscParams v;
    v.kind = c->cfg.viscosity == YASPH_VISCOSITY_PHYSICAL ? 1 : 0;
    v.coeff = c->cfg.viscosity_param * c->mass;  // `epsilon * massj` / `fluid_viscosity * massj`: first product, left to right
    return v;
}

// Runs one Jacobi solve (density: SOLVER 0 / divergence: SOLVER 1): optional warm start, then A/B iterations launched in
// chunks of `speculative_iterations`; kernels past the converged iteration exit immediately on the device-side stop_iter.
// first_a_done: pass A of iteration 0 (with its reduction and decision) already ran fused into the density+alpha sweep.
// advect (dfsph.rs:502-509) fused with the key generation, then the radix sort of the re-sort (dfsph.rs:512)
static int32_t enqueue_advect_sort(yasph_ctx* c, bool only_if_converged) {
    const uint32_t n = c->n;
    pass_begin(c, YASPH_PASS_ADVECT_KEYGEN);
    TRY(radix_prepare(c, n));
    launch_chain(c, k_advect_keygen, keygen_grid(n, c->num_sms), KG_THREADS, 0, c->stream, c->pos, c->vstar, n, c->ctl, c->grid, c->keys[0], c->idx[0], c->radix_scratch,
                 slab_params(c), only_if_converged ? 1u : 0u);
    CHECK_LAUNCH();
    pass_end(c);
    if (only_if_converged) {  // the sort too (neighborhood_update skips it then)
        pass_begin(c, YASPH_PASS_SORT);
        TRY(radix_sort(c, n));
        pass_end(c);
    }
    return YASPH_OK;
}

// The head of a DFSPH step: TimeManager bookkeeping at step entry, then the non-pressure forces + CFL maximum (dfsph.rs:433-477).
// guarded: enqueued by the previous step ahead of its last read-back (yasph_step_n); `vel` is the velocity array of the step.
// seq_out != null: the control block of the step that ends here is snapshot by k_begin_step and published from the side stream.
static int32_t dfsph_head(yasph_ctx* c, bool guarded, const float2* vel, unsigned int* seq_out = nullptr) {
    if (++c->step_token == 0u) c->step_token = 1u;
    launch_chain(c, k_begin_step, 1, 32, 0, c->stream, c->ctl, c->step_token, guarded ? 1u : 0u, seq_out ? c->ctl_snap : nullptr);
    CHECK_LAUNCH();
    if (seq_out) {
        CU(cudaEventRecord(c->ev_tables, c->stream));
        CU(cudaStreamWaitEvent(c->ctl_stream, c->ev_tables, 0));
        TRY(publish_control(c, c->ctl_stream, seq_out, c->ctl_snap));
    }
    // slab mode: v_j and rho_j of the ghosts, if stale.  A guarded head is enqueued only when they are fresh (jacobi_solve), and ahead
    // of the swap that makes v* the velocity (dfsph.rs:524): the field to look at is still called v* then.
    const SlabField vel_field = guarded ? SF_VSTAR : SF_VEL;
    if (!guarded) {
        TRY(slab_refresh(c, SF_VEL, c->vel));
        TRY(slab_refresh(c, SF_DENS, c->dens));
    }
    c->slab.valid[SF_ACCEL] = slab_out_valid(c, {SF_POS, vel_field, SF_DENS}, {});
    pass_begin(c, YASPH_PASS_VISCOSITY);
    OpViscosity v;
    v.vel = vel;
    v.dens = c->dens;
    v.accel = c->accel;
    const float2 g = make_float2(c->cfg.gravity[0], c->cfg.gravity[1]);
    v.base_accel = (g * c->mass) / c->mass;  // dfsph.rs:442-444
    v.vp = visc_params(c);
    v.dt = 0.f;
    v.need_token = guarded ? c->step_token : 0u;
    TRY(launch_sweep(c, v));
    pass_end(c);
    return YASPH_OK;
}

template <int SOLVER>
static int32_t jacobi_solve(yasph_ctx* c, bool first_a_done = false) {
    const float rho0 = c->cfg.fluid_density;
    float* warm_arr = SOLVER == 0 ? c->kappa : c->stiff;
    const SlabField warm_field = SOLVER == 0 ? SF_KAPPA : SF_STIFF;
    pass_begin(c, SOLVER == 0 ? YASPH_PASS_DENSITY_WARM : YASPH_PASS_DIVERGENCE_WARM);
    // The warm start runs iff the previous solve took more than one iteration (dfsph.rs:199,354).  The host mirror of the
    // control block still holds that count (it is refreshed at every read-back and this solve has not started), so the
    // launch is skipped altogether when it would be a no-op; the kernel checks the device-side flag as well.
    const uint32_t prev_iters = c->h_ctl->iters[SOLVER];
    if (prev_iters > 1u) {
        OpJacobiB<SOLVER, true> w;
        w.vstar = c->vstar;
        w.kfac = nullptr;
        w.warm = warm_arr;
        w.clamp_min = -0.5f * rho0 * rho0;
        w.iter_index = 0;
        pass_end(c);
        TRY(slab_refresh(c, warm_field, warm_arr));  // slab mode: the ghosts' warm-start values, if stale
        pass_begin(c, SOLVER == 0 ? YASPH_PASS_DENSITY_WARM : YASPH_PASS_DIVERGENCE_WARM);
        TRY(launch_sweep(c, w));
        c->slab.valid[SF_VSTAR] = slab_out_valid(c, {warm_field}, {SF_VSTAR});
    }
    pass_end(c);
    SolverParams sp;
    sp.max_error = SOLVER == 0 ? c->cfg.dfsph_max_avg_density_error : c->cfg.dfsph_max_divergence_error;
    sp.max_iters = SOLVER == 0 ? c->cfg.dfsph_max_density_iters : c->cfg.dfsph_max_divergence_iters;
    uint32_t it = 0;
    const bool slab = c->slab.active && c->slab.world > 1;
    // Iterations launched before the first read-back: as many as the previous solve needed (the count changes slowly from step to
    // step; at most 8); later chunks launch `speculative_iterations` at a time.  Sweeps past the converged iteration exit on the
    // device-side stop_iter.  Slab mode: the all-reduce (and any halo exchange) of such an iteration still runs on every rank -- all
    // ranks take the same decision from the same all-reduced residual, so the exchanges stay matched; they just move stale values.
    const uint32_t spec = c->cfg.speculative_iterations;
    uint32_t chunk = prev_iters < 1u ? 1u : std::min(prev_iters, std::max(spec, 8u));
    const int solve_pass = SOLVER == 0 ? YASPH_PASS_DENSITY_SOLVE : YASPH_PASS_DIVERGENCE_SOLVE;
    bool radix_ready = false;  // fused advect: the sort's scratch is prepared for the B launches of the coming chunk
    while (true) {
        if (SOLVER == 0 && c->advect_fused_now && !radix_ready) {
            TRY(radix_prepare(c, c->n));
            radix_ready = true;
        }
        pass_begin(c, solve_pass);
        for (uint32_t q = 0; q < chunk; ++q, ++it) {
            if (!(first_a_done && it == 0)) {
                if (slab) {
                    pass_end(c);
                    TRY(slab_refresh(c, SF_VSTAR, c->vstar));  // v* of the ghosts, if stale in the first ghost column
                    pass_begin(c, solve_pass);
                }
                c->slab.valid[SF_KFAC] = slab_out_valid(c, {SF_VSTAR}, {SF_DENS, SF_ALPHA});
                OpJacobiA<SOLVER> a;
                a.vstar = c->vstar;
                a.dens = c->dens;
                a.alpha = c->alpha;
                a.kfac = c->err_buf;
                a.sp = sp;
                a.iter_index = it;
                TRY(launch_sweep(c, a));
            }
            if (c->slab.active) {
                // global residual: sum over the ranks, then the loop decision every rank takes identically
                if (c->slab.peer && c->slab.world > 1) {  // residual all-reduce and the loop decision in one launch
                    JacobiDecideAfter<SOLVER> after{sp, it, (float)c->slab.n_global, c->cfg.fluid_density};
                    launch_chain(c, k_allreduce_peer<JacobiDecideAfter<SOLVER>>, 1, 32, 0, c->stream, (void*)&c->ctl->resid_sum, 1, c->slab.d_boxes, c->slab.rank, c->slab.world,
                                 ++c->slab.ar_seq, c->ctl, after);
                    CHECK_LAUNCH();
                    c->slab.allreduces++;
                } else {
                    TRY(allreduce_scalar(c, &c->ctl->resid_sum, ncclDouble, ncclSum));
                    k_jacobi_decide<SOLVER><<<1, 32, 0, c->stream>>>(c->ctl, sp, it, (float)c->slab.n_global, c->cfg.fluid_density);
                    CHECK_LAUNCH();
                }
                if (slab) {
                    pass_end(c);
                    TRY(slab_refresh(c, SF_KFAC, c->err_buf));  // k_j of the ghosts for pass B, if stale
                    pass_begin(c, solve_pass);
                }
            }
            c->slab.valid[SF_VSTAR] = slab_out_valid(c, {SF_KFAC}, {SF_VSTAR});
            c->slab.valid[warm_field] = it == 0 ? c->slab.valid[SF_KFAC] : slab_own_valid(c, {warm_field, SF_KFAC});
            if (SOLVER == 0 && c->advect_fused_now) {
                OpJacobiBAdvect b;  // the solve's last B also advects and generates the sort keys (sweeps.cuh)
                b.vstar = c->vstar;
                b.kfac = c->err_buf;
                b.warm = warm_arr;
                b.clamp_min = 0.f;
                b.iter_index = it;
                b.pos_out = c->pos_adv;
                b.keys = c->keys[0];
                b.idx = c->idx[0];
                b.sort_scratch = c->radix_scratch;
                b.grid = c->grid;
                b.last = 0u;
                TRY(launch_sweep(c, b));
            } else {
                OpJacobiB<SOLVER, false> b;
                b.vstar = c->vstar;
                b.kfac = c->err_buf;
                b.warm = warm_arr;
                b.clamp_min = 0.f;
                b.iter_index = it;
                TRY(launch_sweep(c, b));
            }
        }
        pass_end(c);  // the pass times are device time of the launches; the read-back below is host latency
        // v* becomes the velocity (dfsph.rs:524) unless the loop goes on: worth a download when this chunk reaches the previous
        // solve's iteration count (a stale copy costs its transfer time ahead of the next chunk's kernels)
        if (SOLVER == 1) TRY(submit_downloads(c));
        if (SOLVER == 1 && it >= prev_iters) TRY(early_velocities(c, c->vstar));
        // Density solve, device-resident stepping: when this chunk reaches the previous solve's iteration count, what follows
        // the solve -- advect + key generation + sort -- is enqueued now, guarded on the device by the solver's own verdict, and
        // the control block is published from the side stream while it runs: the GPU does not idle through the read-back.  If
        // the solve needs more iterations the guarded advect did nothing and the (then meaningless) sort is simply repeated.
        bool spec_now = false;
        if (SOLVER == 0 && c->spec_advect && it >= prev_iters) {
            CU(cudaEventRecord(c->ev_tables, c->stream));
            CU(cudaStreamWaitEvent(c->ctl_stream, c->ev_tables, 0));
            if (c->advect_fused_now) {  // keys and histograms come out of the last B: only the sort is left to enqueue
                pass_begin(c, YASPH_PASS_SORT);
                TRY(radix_sort(c, c->n));
                pass_end(c);
                radix_ready = false;  // a sort that turns out to be premature leaves the scratch dirty
            } else {
                TRY(enqueue_advect_sort(c, true));
            }
            spec_now = true;
            c->spec_advect = false;  // once per solve
        }
        // (slab mode: only if the head needs no halo exchange -- the host cannot guard a rendezvous)
        const bool head_fresh = !slab || (c->slab.valid[SF_VSTAR] >= 1 && c->slab.valid[SF_DENS] >= 1);
        if (SOLVER == 1 && c->spec_head && it >= prev_iters && head_fresh) {
            // yasph_step_n: snapshot of the control block (this step's report) in stream order, then the head of the next step,
            // guarded on the device by this solve's verdict; the host waits for the snapshot while the head runs
            unsigned int seq = 0;
            TRY(dfsph_head(c, true, c->vstar, &seq));  // v* becomes the velocity (dfsph.rs:524)
            TRY(await_control(c, c->ctl_stream, seq));
            if (c->h_ctl->stop_iter[SOLVER] != 0xFFFFFFFFu) {
                c->head_enqueued = true;
                break;
            }
        } else {
            TRY(read_control(c, spec_now ? c->ctl_stream : nullptr));
            if (c->h_ctl->stop_iter[SOLVER] != 0xFFFFFFFFu) {
                if (spec_now) c->spec_advect_done = true;
                break;
            }
        }
        if (SOLVER == 1) c->early_vel_stale = true;
        if (it > sp.max_iters + spec + 1) return fail(c, YASPH_ERR_STATE, "jacobi_solve: device loop control did not terminate");
        chunk = spec;
    }
    return YASPH_OK;
}

// dfsph.rs:419-428: first call (or particle count changed): zero the warm-start arrays, sort, densities, alpha
static int32_t dfsph_initialize(yasph_ctx* c) {
    // Vec::resize(n, 0.0) (dfsph.rs:420-423): existing warm-start values stay where they are, only a new tail is zero-filled
    // (clear_cached_data empties the arrays: everything is zero-filled then).  Slab mode: the local set was rebuilt, start clean.
    const uint32_t keep = c->slab.active ? 0u : std::min(c->dfsph_n, c->n);
    if (keep < c->cap_n) {
        CU(cudaMemsetAsync(c->kappa + keep, 0, (size_t)(c->cap_n - keep) * sizeof(float), c->stream));
        CU(cudaMemsetAsync(c->stiff + keep, 0, (size_t)(c->cap_n - keep) * sizeof(float), c->stream));
    }
    c->dfsph_n = c->n;
    GatherPlan gp;
    gp.n2 = 2;
    gp.a2[0] = &c->pos;
    gp.alt2[0] = &c->pos_alt;
    gp.a2[1] = &c->vel;
    gp.alt2[1] = &c->vel_alt;
    TRY(neighborhood_update(c, false, gp));
    if (c->slab.active) c->slab.valid[SF_KAPPA] = c->slab.valid[SF_STIFF] = (int)c->slab.ghost_cols;  // all zero, on every rank
    pass_begin(c, YASPH_PASS_DENSITY_ALPHA);
    OpDensityAlpha<0, true> da;
    da.dens = c->dens;
    da.alpha = c->alpha;
    da.rho_p = nullptr;
    da.stiffness = 0.f;
    TRY(launch_sweep(c, da));
    pass_end(c);
    c->slab.valid[SF_DENS] = c->slab.valid[SF_ALPHA] = slab_out_valid(c, {SF_POS}, {});
    c->dfsph_ready = true;
    return YASPH_OK;
}

static int32_t dfsph_step(yasph_ctx* c) {
    if (!c->dfsph_ready) TRY(dfsph_initialize(c));
    const uint32_t n = c->n;  // local particles (slab mode: owned + ghosts); the neighbourhood update below changes it
    // step entry + non-pressure forces + CFL maximum (dfsph.rs:433-477) -- unless the previous step of a yasph_step_n call enqueued them
    if (c->head_enqueued)
        c->head_enqueued = false;
    else
        TRY(dfsph_head(c, false, c->vel));
    TRY(allreduce_scalar(c, &c->ctl->max_v2_bits, ncclFloat, ncclMax));  // CFL maximum over all ranks (values are >= 0)
    // update timestep + velocity prediction (dfsph.rs:478-491)
    pass_begin(c, YASPH_PASS_PREDICT);
    launch_chain(c, k_timestep_apply<0>, std::max(1u, blocks_for((n + 1) / 2, 256)), 256, 0, c->stream, c->ctl, c->tp, c->radius * 2.0f, c->vel, c->accel, c->vstar, n);
    CHECK_LAUNCH();
    pass_end(c);
    c->slab.valid[SF_VSTAR] = slab_own_valid(c, {SF_VEL, SF_ACCEL});  // the ghosts predict with their own (recomputed) accelerations
    c->spec_advect = !c->slab.active;  // one GPU (also under yasph_step_host: no download is in flight before the positions are final)
    c->spec_advect_done = false;
    // one GPU, every tile staged: the solve's last Jacobi B advects and generates the sort keys itself (OpJacobiBAdvect)
    c->advect_fused_now = c->fuse_advect && !c->slab.active && !c->unstaged_tiles && c->num_tiles != 0;
    TRY(jacobi_solve<0>(c));  // dfsph.rs:496
    c->spec_advect = false;
    // advect (dfsph.rs:502-509) fused with the key generation of the re-sort (dfsph.rs:512) -- unless it already ran ahead of the read-back
    const bool sorted_ready = c->spec_advect_done;
    if (c->advect_fused_now) {
        std::swap(c->pos, c->pos_adv);  // the advected positions (the old ones are not needed any more)
        c->advect_fused_now = false;
    } else if (!sorted_ready) {
        TRY(enqueue_advect_sort(c, false));
    }
    c->slab.valid[SF_POS] = slab_own_valid(c, {SF_POS, SF_VSTAR});
    {
        // the reference also permutes the old velocities, which are discarded at the final swap (quirk Q7): skipped.
        GatherPlan gp;
        gp.n2 = 2;
        gp.a2[0] = &c->pos;
        gp.alt2[0] = &c->pos_alt;
        gp.a2[1] = &c->vstar;
        gp.alt2[1] = &c->vstar_alt;
        if ((c->cfg.flags & YASPH_FLAG_PERMUTE_WARMSTART) || c->slab.active) {
            gp.n1 = 2;
            gp.a1[0] = &c->kappa;
            gp.alt1[0] = &c->f_alt0;
            gp.a1[1] = &c->stiff;
            gp.alt1[1] = &c->f_alt1;
        }
        TRY(neighborhood_update(c, true, gp, true, sorted_ready));  // positions are final (dfsph.rs:502-512)
    }
    pass_begin(c, YASPH_PASS_DENSITY_ALPHA);
    // The divergence warm start runs iff the previous divergence solve took more than one iteration (dfsph.rs:354); the
    // host mirror still holds that count.  Without a warm start, iteration 0's density-change pass reads the same v* and
    // positions as the density / alpha passes and is fused into their sweep.
    const bool fuse_div_a0 = c->h_ctl->iters[1] <= 1u;
    if (fuse_div_a0) {
        OpDensityAlphaDiv f;  // dfsph.rs:516-518 + iteration 0 of dfsph.rs:372
        f.vstar = c->vstar;
        f.dens = c->dens;
        f.alpha = c->alpha;
        f.kfac = c->err_buf;
        f.sp.max_error = c->cfg.dfsph_max_divergence_error;
        f.sp.max_iters = c->cfg.dfsph_max_divergence_iters;
        TRY(launch_sweep(c, f));
        c->slab.valid[SF_KFAC] = slab_out_valid(c, {SF_POS, SF_VSTAR}, {});
    } else {
        OpDensityAlpha<0, true> da;  // dfsph.rs:516-518
        da.dens = c->dens;
        da.alpha = c->alpha;
        da.rho_p = nullptr;
        da.stiffness = 0.f;
        TRY(launch_sweep(c, da));
    }
    pass_end(c);
    TRY(early_download_owned(c, &c->early_dens_out, c->dens, c->f_alt0, 1));  // densities are final (dfsph.rs:516)
    c->slab.valid[SF_DENS] = c->slab.valid[SF_ALPHA] = slab_out_valid(c, {SF_POS}, {});
    TRY(jacobi_solve<1>(c, fuse_div_a0));  // dfsph.rs:521
    std::swap(c->vel, c->vstar);    // dfsph.rs:524
    std::swap(c->slab.valid[SF_VEL], c->slab.valid[SF_VSTAR]);
    return YASPH_OK;
}

// The head of a WCSPH step: step entry, then leap frog 1 (wscsph.rs:141-150) fused with key generation.  Nothing in it depends on a
// verdict of the previous step, so yasph_step_n enqueues it ahead of that step's read-back without a guard.
static int32_t wcsph_head(yasph_ctx* c, unsigned int* seq_out = nullptr) {
    const uint32_t n = c->n;
    if (++c->step_token == 0u) c->step_token = 1u;
    launch_chain(c, k_begin_step, 1, 32, 0, c->stream, c->ctl, c->step_token, 0u, seq_out ? c->ctl_snap : nullptr);
    CHECK_LAUNCH();
    if (seq_out) {  // the report of the step that ends here leaves from the side stream while this step's first kernels run
        CU(cudaEventRecord(c->ev_tables, c->stream));
        CU(cudaStreamWaitEvent(c->ctl_stream, c->ev_tables, 0));
        TRY(publish_control(c, c->ctl_stream, seq_out, c->ctl_snap));
    }
    c->slab.valid[SF_VEL] = slab_own_valid(c, {SF_VEL, SF_ACCEL});
    c->slab.valid[SF_POS] = slab_own_valid(c, {SF_POS, SF_VEL});
    pass_begin(c, YASPH_PASS_ADVECT_KEYGEN);
    TRY(radix_prepare(c, n));
    launch_chain(c, k_kickdrift_keygen, keygen_grid(n, c->num_sms), KG_THREADS, 0, c->stream, c->pos, c->vel, c->accel, n, c->ctl, c->grid, c->keys[0], c->idx[0],
                 c->radix_scratch, slab_params(c));
    CHECK_LAUNCH();
    pass_end(c);
    return YASPH_OK;
}

static int32_t wcsph_step(yasph_ctx* c) {
    uint32_t n = c->n;
    if (c->head_enqueued)
        c->head_enqueued = false;
    else
        TRY(wcsph_head(c));
    GatherPlan gp;
    gp.n2 = 2;
    gp.a2[0] = &c->pos;
    gp.alt2[0] = &c->pos_alt;
    gp.a2[1] = &c->vel;
    gp.alt2[1] = &c->vel_alt;
    TRY(neighborhood_update(c, true, gp, true));  // wscsph.rs:153; positions are final (wscsph.rs:141-153)
    n = c->n;
    pass_begin(c, YASPH_PASS_DENSITY_ALPHA);
    TRY((launch_density<1, true>(c)));  // Poly6, wscsph.rs:154; + Tait pressure per particle (wscsph.rs:91-92)
    pass_end(c);
    TRY(early_download_owned(c, &c->early_dens_out, c->dens, c->f_alt0, 1));  // densities are final (wscsph.rs:154)
    c->slab.valid[SF_DENS] = c->slab.valid[SF_VSTAR] = slab_out_valid(c, {SF_POS}, {});  // (rho, p) lives in the v* buffer
    TRY(slab_refresh(c, SF_VSTAR, c->vstar));  // (rho, p) and v of the ghosts, if stale in the first ghost column
    TRY(slab_refresh(c, SF_VEL, c->vel));
    c->slab.valid[SF_ACCEL] = slab_out_valid(c, {SF_POS, SF_VEL, SF_VSTAR}, {});
    pass_begin(c, YASPH_PASS_WCSPH_ACCEL);
    {
        OpWcsphAccel a;  // wscsph.rs:155
        a.vel = c->vel;
        a.rho_p = c->vstar;
        a.accel = c->accel;
        a.gravity = make_float2(c->cfg.gravity[0], c->cfg.gravity[1]);
        a.vp = visc_params(c);
        a.boundary_force_factor = c->cfg.wcsph_boundary_force_factor;
        a.dt = 0.f;
        TRY(launch_sweep(c, a));
    }
    pass_end(c);
    TRY(allreduce_scalar(c, &c->ctl->max_v2_bits, ncclFloat, ncclMax));
    // update timestep + leap frog 2 (wscsph.rs:160-177)
    pass_begin(c, YASPH_PASS_WCSPH_KICK);
    launch_chain(c, k_timestep_apply<1>, std::max(1u, blocks_for((n + 1) / 2, 256)), 256, 0, c->stream, c->ctl, c->tp, c->radius * 2.0f, c->vel, c->accel, c->vel, n);
    CHECK_LAUNCH();
    pass_end(c);
    c->slab.valid[SF_VEL] = slab_own_valid(c, {SF_VEL, SF_ACCEL});
    return YASPH_OK;
}

extern "C" int32_t yasph_solver_state_get(yasph_ctx* c, yasph_solver_state* out) {
    if (!c || !out) return YASPH_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(c->device));
    TRY(read_control(c));
    memset(out, 0, sizeof(*out));
    out->step_ns = c->h_ctl->step_ns;
    out->iters_density = c->h_ctl->iters[0];
    out->iters_divergence = c->h_ctl->iters[1];
    out->initialized = (c->cfg.solver == YASPH_SOLVER_WCSPH || c->dfsph_ready) ? 1u : 0u;
    // between steps the device holds the total up to and including the last step (k_begin_step adds the next one); a total the host
    // supplied for the coming step (total_is_current) already contains that step
    out->total_simulated_ns = c->h_ctl->total_simulated_ns - (c->h_ctl->total_is_current ? c->h_ctl->step_ns : 0ull);
    return YASPH_OK;
}
extern "C" int32_t yasph_solver_state_set(yasph_ctx* c, const yasph_solver_state* in) {
    if (!c || !in) return YASPH_ERR_INVALID_ARGUMENT;
    if (c->slab.active) return fail(c, YASPH_ERR_STATE, "yasph_solver_state_set: not available in slab mode");
    if (!c->have_particles || c->n == 0) return fail(c, YASPH_ERR_STATE, "yasph_solver_state_set: upload the particles first");
    CU(cudaSetDevice(c->device));
    if (in->initialized && c->cfg.solver == YASPH_SOLVER_DFSPH && !c->dfsph_ready) {
        TRY(dfsph_initialize(c));
        TRY(read_control(c));
        TRY(check_capacity_flags(c));
    }
    const unsigned int it[2] = {in->iters_density, in->iters_divergence};
    CU(cudaMemcpyAsync(&c->ctl->iters[0], it, sizeof(it), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(&c->ctl->step_ns, &in->step_ns, sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    struct {
        unsigned long long total;
        unsigned int is_current, pad;
    } tv = {in->total_simulated_ns, 0u, 0u};  // k_begin_step adds the next step itself
    CU(cudaMemcpyAsync(&c->ctl->total_simulated_ns, &tv, 12, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->h_ctl->iters[0] = it[0];
    c->h_ctl->iters[1] = it[1];
    c->h_ctl->step_ns = in->step_ns;
    return YASPH_OK;
}
extern "C" int32_t yasph_upload_field(yasph_ctx* c, int32_t field, const void* data, uint64_t bytes) {
    if (!c || !data) return YASPH_ERR_INVALID_ARGUMENT;
    if (c->slab.active) return fail(c, YASPH_ERR_STATE, "yasph_upload_field: not available in slab mode");
    CU(cudaSetDevice(c->device));
    void* dst = nullptr;
    size_t elem = sizeof(float);
    switch (field) {
        case YASPH_FIELD_KAPPA: dst = c->kappa; break;
        case YASPH_FIELD_STIFFNESS: dst = c->stiff; break;
        case YASPH_FIELD_ACCELERATION:
            dst = c->accel;
            elem = sizeof(float2);
            break;
        default: return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_upload_field: field %d cannot be uploaded", field);
    }
    if (bytes != (uint64_t)c->n * elem)
        return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_upload_field: %llu bytes, expected %zu", (unsigned long long)bytes, (size_t)c->n * elem);
    if (bytes) CU(cudaMemcpyAsync(dst, data, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}

#if defined(YASPH_SWEEP_TIMING) || defined(YASPH_LIST_TIMING)
// profiling builds only: cycle counters of the sweep pipeline / the list build since the last call (then reset)
extern "C" int32_t yasph_debug_sweep_counters(yasph_ctx* c, unsigned long long* out8) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(out8, c->sweep_dbg, 128, cudaMemcpyDeviceToHost));
    CU(cudaMemset(c->sweep_dbg, 0, 128));
    CU(cudaMemset(c->sweep_dbg + 8, 0xFF, 8));
    return YASPH_OK;
}
#endif

#ifdef YASPH_RADIX_TIMING
extern "C" int32_t yasph_debug_radix_counters(yasph_ctx* c, unsigned long long* out8) {
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpyFromSymbol(out8, g_radix_dbg, 64));
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CU(cudaMemcpyToSymbol(g_radix_dbg, z, 64));
    return YASPH_OK;
}
#endif

extern "C" int32_t yasph_step(yasph_ctx* c, yasph_step_report* report) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    if (!c->have_particles || (c->n == 0 && !c->slab.active)) return fail(c, YASPH_ERR_STATE, "yasph_step: no particles uploaded");
    CU(cudaSetDevice(c->device));
    if (c->cfg.solver == YASPH_SOLVER_WCSPH)
        TRY(wcsph_step(c));
    else
        TRY(dfsph_step(c));
    // the divergence solve ends with a read-back (the snapshot of the step's control block): the DFSPH step's h_ctl is current
    if (c->cfg.solver == YASPH_SOLVER_WCSPH) {
        TRY(submit_downloads(c));
        TRY(early_velocities(c, c->vel));
        if (c->spec_head) {  // yasph_step_n: the next step's head runs while the host waits for the snapshot
            unsigned int seq = 0;
            TRY(wcsph_head(c, &seq));
            c->head_enqueued = true;
            TRY(await_control(c, c->ctl_stream, seq));
        } else {
            TRY(read_control(c));
        }
    }
    pass_resolve(c);
    TRY(check_capacity_flags(c));
    fill_report(c, report);
    if (c->h_ctl->nonfinite) return fail(c, YASPH_ERR_NONFINITE, "non-finite Jacobi residual (solver mask %u)", c->h_ctl->nonfinite);
    return YASPH_OK;
}

// `steps` calls of yasph_step (the application's frame loop, main.rs:339-360: several simulation steps per frame); reports, if not
// null, receives one report per step.  Same results as the single calls; with device-resident particles the head of step s + 1 is
// enqueued ahead of the read-back that ends step s (guarded on the device by step s's own verdict where it depends on one; on slabs
// only when that head needs no halo exchange), so the GPU does not idle between the steps.
extern "C" int32_t yasph_step_n(yasph_ctx* c, uint32_t steps, yasph_step_report* reports) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    const bool can_spec = !(c->cfg.flags & YASPH_FLAG_PROFILE_PASSES) && c->early_pos_out == nullptr && c->early_vel_out == nullptr && c->early_dens_out == nullptr;
    int32_t rc = YASPH_OK;
    for (uint32_t s = 0; s < steps && rc == YASPH_OK; ++s) {
        c->spec_head = can_spec && s + 1 < steps && (c->cfg.solver == YASPH_SOLVER_WCSPH || c->dfsph_ready);
        rc = yasph_step(c, reports ? reports + s : nullptr);
    }
    c->spec_head = false;
    if (rc != YASPH_OK && c->head_enqueued) {  // a head that is already in the stream belongs to a step that will not run
        cudaStreamSynchronize(c->stream);
        c->head_enqueued = false;
    }
    return rc;
}

static bool is_pinned_host(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

extern "C" int32_t yasph_step_host(yasph_ctx* c, float* pos_xy, float* vel_xy, float* densities, uint32_t n, yasph_step_report* report) {
    return yasph_step_host_ex(c, pos_xy, vel_xy, densities, n, 0u, report);
}
extern "C" int32_t yasph_step_host_ex(yasph_ctx* c, float* pos_xy, float* vel_xy, float* densities, uint32_t n, uint32_t options, yasph_step_report* report) {
    if (!c || !pos_xy || !vel_xy) return YASPH_ERR_INVALID_ARGUMENT;
    if (n == 0 || n > c->cap_n) return fail(c, YASPH_ERR_CAPACITY, "yasph_step_host: n=%u out of range (max_particles=%u)", n, c->cap_n);
    if (c->slab.active) return fail(c, YASPH_ERR_STATE, "yasph_step_host: the particle count of a slab changes with migration, use yasph_step_host_slab");
    CU(cudaSetDevice(c->device));
    const bool unchanged = (options & YASPH_HOST_INPUT_UNCHANGED) != 0;
    if (unchanged && (!c->have_particles || n != c->n))
        return fail(c, YASPH_ERR_STATE, "yasph_step_host_ex: YASPH_HOST_INPUT_UNCHANGED, but the device holds %u particles and the call passes %u", c->have_particles ? c->n : 0u, n);
    if (n != c->n) TRY(reset_particle_set(c, n));
    const bool prof = (c->cfg.flags & YASPH_FLAG_PROFILE_PASSES) != 0;
    if (prof) CU(cudaEventRecord(c->ev_host[0], c->stream));
    if (!unchanged) {
        CU(cudaMemcpyAsync(c->pos, pos_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->vel, vel_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
    }
    if (prof) CU(cudaEventRecord(c->ev_host[1], c->stream));
    // Pinned (or registered) host arrays take their results as soon as they are final, on the copy stream, overlapped with the
    // remaining passes; pageable arrays would block the launching thread in mid-step, so they are copied at the end as before.
    const bool pin_pos = is_pinned_host(pos_xy), pin_vel = is_pinned_host(vel_xy), pin_dens = densities && is_pinned_host(densities);
    c->early_pos_out = pin_pos ? pos_xy : nullptr;
    c->early_dens_out = pin_dens ? densities : nullptr;
    c->early_vel_out = pin_vel ? vel_xy : nullptr;
    c->early_vel_stale = true;  // until a download has been enqueued
    int32_t rc = yasph_step(c, report);
    // still armed: the step did not pass the hand-over point
    float* late_pos = (!pin_pos || c->early_pos_out) ? pos_xy : nullptr;
    float* late_dens = densities && (!pin_dens || c->early_dens_out) ? densities : nullptr;
    const bool late_vel = !pin_vel || c->early_vel_stale;
    c->early_pos_out = c->early_dens_out = c->early_vel_out = nullptr;
    if (rc != YASPH_OK) c->pending[0].host = c->pending[1].host = nullptr;
    if (rc != YASPH_OK) {
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->copy_stream);  // nothing of this call stays in flight towards the caller's arrays
        return rc;
    }
    if (late_vel) {
        if (prof) CU(cudaEventRecord(c->ev_host[4], c->stream));
        CU(cudaMemcpyAsync(vel_xy, c->vel, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
        if (prof) CU(cudaEventRecord(c->ev_host[5], c->stream));
    }
    if (late_pos) {
        CU(cudaMemcpyAsync(late_pos, c->pos, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
        if (prof) CU(cudaEventRecord(c->ev_host[2], c->stream));
    }
    if (late_dens) {
        CU(cudaMemcpyAsync(late_dens, c->dens, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        if (prof) CU(cudaEventRecord(c->ev_host[3], c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    if (prof) {
        for (int i = 0; i < 6; ++i) c->host_us[i] = 0.f;
        for (int i = 1; i < 6; ++i) {
            if (i == 3 && !densities) continue;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, c->ev_host[0], c->ev_host[i]) == cudaSuccess) c->host_us[i] = ms * 1000.f;
        }
        cudaGetLastError();
    }
    return YASPH_OK;
}

extern "C" int32_t yasph_host_step_times(yasph_ctx* c, float* out_us) {
    if (!c || !out_us) return YASPH_ERR_INVALID_ARGUMENT;
    memcpy(out_us, c->host_us, sizeof(c->host_us));
    return YASPH_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// multi-GPU surface
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int32_t yasph_comm_unique_id(void* out_id, uint64_t bytes) {
    yasph_ctx* c = nullptr;
    if (!out_id || bytes < YASPH_COMM_ID_BYTES) return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_comm_unique_id: need a %u-byte buffer", YASPH_COMM_ID_BYTES);
    static_assert(sizeof(ncclUniqueId) == YASPH_COMM_ID_BYTES, "ncclUniqueId size");
    if (!nccl_load()) return fail(c, YASPH_ERR_COMM, "%s", g_nccl.error.c_str());
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(out_id, &id, sizeof(id));
    return YASPH_OK;
}

extern "C" int32_t yasph_comm_init(yasph_ctx* c, int32_t rank, int32_t world, const void* id_bytes, uint64_t bytes) {
    if (!c || !id_bytes || bytes < YASPH_COMM_ID_BYTES || world < 1 || rank < 0 || rank >= world)
        return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_comm_init: bad arguments (rank %d of %d)", rank, world);
    if (c->slab.comm) return fail(c, YASPH_ERR_STATE, "yasph_comm_init: communicator already initialised");
    if (!nccl_load()) return fail(c, YASPH_ERR_COMM, "%s", g_nccl.error.c_str());
    CU(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t comm = nullptr;
    NC(g_nccl.CommInitRank(&comm, world, id, rank));
    c->slab.comm = comm;
    c->slab.rank = rank;
    c->slab.world = world;
    return YASPH_OK;
}

extern "C" int32_t yasph_loopback_create(int32_t world, void** fabric) {
    if (!fabric || world < 1) return YASPH_ERR_INVALID_ARGUMENT;
    Fabric* f = new Fabric();
    f->world = world;
    f->post.resize((size_t)world * 2);
    f->ack.resize((size_t)world * 2);
    f->ar_val.resize(world);
    *fabric = f;
    return YASPH_OK;
}
extern "C" int32_t yasph_loopback_destroy(void* fabric) {
    if (!fabric) return YASPH_ERR_INVALID_ARGUMENT;
    delete static_cast<Fabric*>(fabric);
    return YASPH_OK;
}
extern "C" int32_t yasph_comm_init_loopback(yasph_ctx* c, void* fabric, int32_t rank) {
    if (!c || !fabric) return YASPH_ERR_INVALID_ARGUMENT;
    Fabric* f = static_cast<Fabric*>(fabric);
    if (rank < 0 || rank >= f->world) return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_comm_init_loopback: rank %d of %d", rank, f->world);
    if (c->slab.comm || c->slab.fabric) return fail(c, YASPH_ERR_STATE, "yasph_comm_init_loopback: communicator already initialised");
    CU(cudaSetDevice(c->device));
    for (int sd = 0; sd < 2; ++sd) {
        CU(cudaEventCreateWithFlags(&c->slab.ev_ready[sd], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->slab.ev_done[sd], cudaEventDisableTiming));
    }
    c->slab.fabric = f;
    c->slab.rank = rank;
    c->slab.world = f->world;
    return YASPH_OK;
}

// Peer-memory transport: allocate the mailbox, hand its IPC handle to every rank (all-gather over the NCCL communicator that
// exists anyway), map the peers' mailboxes.  Collective; falls back to NCCL send/recv + all-reduce on every rank if any rank
// cannot map a peer (no NVLink / peer access) or YASPH_FLAG_NO_PEER_TRANSPORT is set.
static int32_t peer_setup(yasph_ctx* c) {
    auto& sl = c->slab;
    if (sl.peer_tried || !sl.comm || sl.world < 2) return YASPH_OK;
    sl.peer_tried = true;
    if (sl.world > PEER_MAX_RANKS) return YASPH_OK;
    ncclComm_t comm = (ncclComm_t)sl.comm;
    int ok = (c->cfg.flags & YASPH_FLAG_NO_PEER_TRANSPORT) ? 0 : 1;
    const size_t bytes = peer_box_bytes(sl.max_halo);
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok && cudaMalloc(&sl.box, bytes) != cudaSuccess) ok = 0;
    if (ok && cudaMemset(sl.box, 0, bytes) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, sl.box) != cudaSuccess) ok = 0;
    cudaGetLastError();
    // all-gather (ok flag, handle) records
    struct Rec {
        int ok;
        int pad;
        cudaIpcMemHandle_t h;
    };
    Rec rec;
    rec.ok = ok;
    rec.pad = 0;
    rec.h = mine;
    Rec* d_all = nullptr;
    CU(cudaMalloc((void**)&d_all, sizeof(Rec) * (sl.world + 1)));
    CU(cudaMemcpyAsync(d_all + sl.world, &rec, sizeof(Rec), cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllGather(d_all + sl.world, d_all, sizeof(Rec), ncclChar, comm, c->stream));
    std::vector<Rec> all(sl.world);
    CU(cudaMemcpyAsync(all.data(), d_all, sizeof(Rec) * sl.world, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < sl.world; ++r) ok = ok && all[r].ok;
    if (ok) {
        for (int r = 0; r < sl.world && ok; ++r) {
            if (r == sl.rank) {
                sl.peer_box[r] = sl.box;
            } else if (cudaIpcOpenMemHandle(&sl.peer_box[r], all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                sl.peer_box[r] = nullptr;
                ok = 0;
                cudaGetLastError();
            }
        }
    }
    // every rank must take the same path: all-reduce the verdict (the record buffer doubles as scratch)
    int* d_ok = reinterpret_cast<int*>(d_all);
    CU(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, comm, c->stream));
    CU(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaFree(d_all));
    if (ok) {
        CU(cudaMalloc((void**)&sl.d_boxes, sizeof(void*) * PEER_MAX_RANKS));
        CU(cudaMemcpy(sl.d_boxes, sl.peer_box, sizeof(void*) * PEER_MAX_RANKS, cudaMemcpyHostToDevice));
        CU(cudaMalloc((void**)&sl.d_ticket, sizeof(unsigned int)));
        CU(cudaMemset(sl.d_ticket, 0, sizeof(unsigned int)));
        sl.peer = true;
    } else {
        for (int r = 0; r < sl.world; ++r)
            if (r != sl.rank && sl.peer_box[r]) cudaIpcCloseMemHandle(sl.peer_box[r]);
        memset(sl.peer_box, 0, sizeof(sl.peer_box));
        if (sl.box) cudaFree(sl.box);
        sl.box = nullptr;
        cudaGetLastError();
    }
    return YASPH_OK;
}
static void peer_teardown(yasph_ctx* c) {
    auto& sl = c->slab;
    for (int r = 0; r < PEER_MAX_RANKS; ++r)
        if (r != sl.rank && sl.peer_box[r]) cudaIpcCloseMemHandle(sl.peer_box[r]);
    memset(sl.peer_box, 0, sizeof(sl.peer_box));
    if (sl.box) cudaFree(sl.box);
    if (sl.d_boxes) cudaFree(sl.d_boxes);
    if (sl.d_ticket) cudaFree(sl.d_ticket);
    sl.box = nullptr;
    sl.d_boxes = nullptr;
    sl.d_ticket = nullptr;
    sl.peer = false;
}

extern "C" int32_t yasph_slab_set(yasph_ctx* c, uint32_t col_lo, uint32_t col_hi, uint64_t n_global, uint32_t id_base) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    auto& sl = c->slab;
    if (sl.world > 1 && !sl.comm && !sl.fabric) return fail(c, YASPH_ERR_STATE, "yasph_slab_set: call yasph_comm_init first");
    if (sl.rank == 0) col_lo = 0;                  // the end slabs extend to the ends of the grid
    if (sl.rank == sl.world - 1) col_hi = 65536u;
    if (col_lo >= col_hi || col_hi > 65536u) return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_slab_set: empty or invalid column range [%u, %u)", col_lo, col_hi);
    CU(cudaSetDevice(c->device));
    sl.col_lo = col_lo;
    sl.col_hi = col_hi;
    sl.n_global = n_global;
    sl.id_base = id_base;
    sl.max_halo = c->cfg.max_halo;
    sl.ghost_cols = c->cfg.ghost_columns ? c->cfg.ghost_columns : 1u;
    if (sl.world > 1 && sl.rank > 0 && sl.rank + 1 < sl.world && col_hi - col_lo < sl.ghost_cols + 1u)
        return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_slab_set: slab [%u, %u) between two others must be wider than the ghost layer (ghost_columns = %u)", col_lo, col_hi,
                    sl.ghost_cols);
    if (!sl.pflag) {
        CU(dmalloc(&sl.pflag, (size_t)c->cap_n));
        CU(cudaMemsetAsync(sl.pflag, 0, c->cap_n, c->stream));
        for (int sd = 0; sd < 2; ++sd) {
            CU(dmalloc(&sl.sel[sd], (size_t)sl.max_halo));
            CU(dmalloc(&sl.sel_g[sd], (size_t)sl.max_halo));
            CU(dmalloc(&sl.send_idx[sd], (size_t)sl.max_halo));
            CU(dmalloc(&sl.ghost_idx[sd], (size_t)sl.max_halo));
            CU(dmalloc(&sl.sbuf[sd], (size_t)sl.max_halo * RECORD_MAX_BYTES));
            CU(dmalloc(&sl.rbuf[sd], (size_t)sl.max_halo * RECORD_MAX_BYTES));
        }
        CU(dmalloc(&sl.d_cnt, 8));
        CU(cudaMallocHost((void**)&sl.h_cnt, 32 * sizeof(uint32_t)));
        CU(cudaHostAlloc((void**)&sl.h_pcounts, sizeof(PeerCounts), cudaHostAllocMapped));
        memset(sl.h_pcounts, 0, sizeof(PeerCounts));
        CU(cudaHostGetDevicePointer((void**)&sl.d_pcounts, sl.h_pcounts, 0));
    }
    TRY(peer_setup(c));
    sl.active = true;
    sl.n_own = 0;
    sl.n_ghost[0] = sl.n_ghost[1] = sl.n_send[0] = sl.n_send[1] = 0;
    c->n = 0;
    c->have_particles = false;
    c->lists_valid = false;
    c->dfsph_ready = false;
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}

extern "C" int32_t yasph_slab_get(yasph_ctx* c, yasph_slab_info* out) {
    if (!c || !out) return YASPH_ERR_INVALID_ARGUMENT;
    const auto& sl = c->slab;
    memset(out, 0, sizeof(*out));
    out->rank = sl.rank;
    out->world = sl.world;
    out->col_lo = sl.col_lo;
    out->col_hi = sl.col_hi;
    out->n_own = sl.active ? sl.n_own : c->n;
    out->n_local = c->n;
    out->n_ghost_left = sl.n_ghost[0];
    out->n_ghost_right = sl.n_ghost[1];
    out->migrated_out_left = sl.mig_out[0];
    out->migrated_out_right = sl.mig_out[1];
    out->migrated_in = sl.mig_in;
    out->peer_transport = sl.peer ? 1u : 0u;
    out->n_global = sl.active ? sl.n_global : c->n;
    out->halo_exchanges = sl.halo_exchanges;
    out->allreduces = sl.allreduces;
    return YASPH_OK;
}

extern "C" int32_t yasph_cell_column(const yasph_config* cfg, float x, uint32_t* column) {
    if (!cfg || !column || !(cfg->smoothing_length > 0.f)) return YASPH_ERR_INVALID_ARGUMENT;
    const float inv = 1.0f / cfg->smoothing_length;  // neighborhood_search.rs:475
    *column = f32_as_u16((x - cfg->grid_min[0]) * inv);
    return YASPH_OK;
}

extern "C" int32_t yasph_step_host_slab(yasph_ctx* c, float* pos_xy, float* vel_xy, float* densities, uint32_t n_in, uint32_t capacity, uint32_t* n_out,
                                        yasph_step_report* report) {
    return yasph_step_host_slab_ex(c, pos_xy, vel_xy, densities, n_in, capacity, 0u, n_out, report);
}
extern "C" int32_t yasph_step_host_slab_ex(yasph_ctx* c, float* pos_xy, float* vel_xy, float* densities, uint32_t n_in, uint32_t capacity, uint32_t options,
                                           uint32_t* n_out, yasph_step_report* report) {
    if (!c || !pos_xy || !vel_xy || !n_out) return YASPH_ERR_INVALID_ARGUMENT;
    if (!c->slab.active) return fail(c, YASPH_ERR_STATE, "yasph_step_host_slab: yasph_slab_set has not been called");
    if (n_in > c->cap_n || n_in > capacity) return fail(c, YASPH_ERR_CAPACITY, "yasph_step_host_slab: n_in=%u out of range", n_in);
    CU(cudaSetDevice(c->device));
    auto& sl = c->slab;
    if (options & YASPH_HOST_INPUT_UNCHANGED) {
        // the caller has not written to the arrays since the previous call handed them back: the device still holds this rank's
        // particles (between its ghosts, in the sorted order) -- nothing to upload, the arrays are outputs only
        if (!c->have_particles || !c->lists_valid || n_in != sl.n_own)
            return fail(c, YASPH_ERR_STATE, "yasph_step_host_slab_ex: YASPH_HOST_INPUT_UNCHANGED, but the device holds %u owned particles and the call passes %u",
                        c->have_particles ? sl.n_own : 0u, n_in);
    } else if (c->have_particles && c->lists_valid && n_in == sl.n_own) {
        // the arrays are the owned particles in the order of the last download: put them back between the ghosts
        TRY(ensure_own_index(c, false));
        if (n_in) {
            CU(cudaMemcpyAsync(c->pos_alt, pos_xy, (size_t)n_in * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(c->vel_alt, vel_xy, (size_t)n_in * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
            k_own_scatter<float2><<<blocks_for(n_in, 256), 256, 0, c->stream>>>(c->pos, sl.own_idx, n_in, c->pos_alt);
            CHECK_LAUNCH();
            k_own_scatter<float2><<<blocks_for(n_in, 256), 256, 0, c->stream>>>(c->vel, sl.own_idx, n_in, c->vel_alt);
            CHECK_LAUNCH();
            sl.valid[SF_VEL] = 0;  // the caller may have changed the velocities it owns: the neighbours' ghosts of them are refreshed
        }
    } else {
        TRY(reset_particle_set(c, n_in));
        if (n_in) {
            CU(cudaMemcpyAsync(c->pos, pos_xy, (size_t)n_in * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(c->vel, vel_xy, (size_t)n_in * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
        }
    }
    // Pinned host arrays with room for every particle this rank can own take the positions and densities as soon as they are final,
    // on the copy stream, while the step goes on (as yasph_step_host does); the velocities and everything else follow at the end.
    const bool roomy = capacity >= c->cap_n;
    c->early_pos_out = roomy && is_pinned_host(pos_xy) ? pos_xy : nullptr;
    c->early_dens_out = roomy && densities && is_pinned_host(densities) ? densities : nullptr;
    sl.own_count_unchecked = false;
    // every rank steps, also one that currently owns no particle (the collectives are matched)
    int32_t rc = c->cfg.solver == YASPH_SOLVER_WCSPH ? wcsph_step(c) : dfsph_step(c);
    if (rc == YASPH_OK) rc = submit_downloads(c);
    if (rc == YASPH_OK) rc = read_control(c);
    const bool late_pos = c->early_pos_out != nullptr || !(roomy && is_pinned_host(pos_xy));
    const bool late_dens = densities && (c->early_dens_out != nullptr || !(roomy && is_pinned_host(densities)));
    c->early_pos_out = c->early_dens_out = nullptr;
    if (rc != YASPH_OK) {
        c->pending[0].host = c->pending[1].host = nullptr;
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->copy_stream);  // nothing of this call stays in flight towards the caller's arrays
        return rc;
    }
    pass_resolve(c);
    rc = check_capacity_flags(c);
    if (rc == YASPH_OK && sl.own_count_unchecked && (uint32_t)(c->h_ctl->slab_migrants & 0xFFFFFFFFull) != sl.n_own)
        rc = fail(c, YASPH_ERR_STATE, "rank %d: %llu owned particles flagged, %u expected", sl.rank, c->h_ctl->slab_migrants & 0xFFFFFFFFull, sl.n_own);
    sl.own_count_unchecked = false;
    fill_report(c, report);
    if (rc == YASPH_OK && c->h_ctl->nonfinite) rc = fail(c, YASPH_ERR_NONFINITE, "non-finite Jacobi residual (solver mask %u)", c->h_ctl->nonfinite);
    *n_out = sl.n_own;
    if (rc == YASPH_OK && sl.n_own > capacity) rc = fail(c, YASPH_ERR_CAPACITY, "yasph_step_host_slab: %u owned particles after the step > capacity %u", sl.n_own, capacity);
    if (rc == YASPH_OK) rc = download_array(c, c->vel, c->vel_alt, vel_xy);
    if (rc == YASPH_OK && late_pos) rc = download_array(c, c->pos, c->pos_alt, pos_xy);
    if (rc == YASPH_OK && late_dens) rc = download_array(c, c->dens, c->f_alt0, densities);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    return rc;
}

#include "scene.inl"
