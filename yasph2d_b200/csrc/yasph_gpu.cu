// yasph_gpu.cu -- context, step orchestration and the C ABI (include/yasph_gpu.h) of libyasph_gpu.so.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false (see __graft_entry__.build()).
// There is no CPU fallback anywhere in this library: without a CUDA device yasph_create fails with
// YASPH_ERR_NO_DEVICE.
#include "../../include/yasph_gpu.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "scan.cuh"
#include "sweeps.cuh"

using namespace yasph;

static thread_local std::string g_create_error;

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
    uint32_t num_tiles = 0;                  // host copy of Control::num_tiles for the current structure == grid of the tile kernels
    size_t smem_optin = 0;
    float mass = 0, radius = 0;
    GridParams grid;
    KernelConsts kc;
    TimeParams tp;
    // particle state (ping-pong pairs are swapped by the gather)
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
    uchar2* counts = nullptr;
    // scratch
    uint32_t* radix_scratch = nullptr;
    unsigned long long *scan_chunks = nullptr, *scan_total = nullptr;
    double* partials = nullptr;
    Control* ctl = nullptr;
    Control* h_ctl = nullptr;  // pinned mirror
    // export scratch (allocated on demand)
    uint16_t *exp_cd = nullptr, *exp_ct = nullptr;
    uint32_t* exp_lists = nullptr;
    // pinned staging for yasph_step_host
    float *h_stage = nullptr;
    size_t h_stage_bytes = 0;
    // state flags
    bool have_particles = false, lists_valid = false, dfsph_ready = false;
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
static int32_t fail(yasph_ctx* c, int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c)
        c->err = buf;
    else
        g_create_error = buf;
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
    return cudaMalloc((void**)p, (count ? count : 1) * sizeof(T));
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
static void pass_begin(yasph_ctx* c, int pass) {
    if (!(c->cfg.flags & YASPH_FLAG_PROFILE_PASSES)) return;
    c->cur_pass = pass;
    c->cur_start = get_event(c);
    cudaEventRecord(c->cur_start, c->stream);
}
static void pass_end(yasph_ctx* c) {
    if (!(c->cfg.flags & YASPH_FLAG_PROFILE_PASSES) || c->cur_pass < 0) return;
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
static size_t worst_smem_bytes(uint32_t cap_dyn, uint32_t cap_stat) {
    const size_t a = list_smem_bytes(cap_dyn, cap_stat), b = sweep_smem_bytes<OpWcsphAccel>(cap_dyn, cap_stat);
    return a > b ? a : b;
}

static void free_all(yasph_ctx* c) {
    void* ptrs[] = {c->pos, c->pos_alt, c->vel, c->vel_alt, c->vstar, c->vstar_alt, c->accel, c->dens, c->alpha, c->kappa, c->stiff, c->err_buf,
                    c->f_alt0, c->f_alt1, c->keys[0], c->keys[1], c->idx[0], c->idx[1], c->cell_key, c->cell_start, c->tile_key, c->tile_pstart,
                    c->tile_cstart, c->bpos, c->bpos_alt, c->scell_key, c->scell_start, c->stile_key, c->stile_cstart, c->tile_runs, c->cslot_d,
                    c->cslot_s, c->lists, c->counts,
                    c->radix_scratch, c->scan_chunks, c->scan_total, c->partials, c->ctl, c->exp_cd, c->exp_ct, c->exp_lists};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (c->h_ctl) cudaFreeHost(c->h_ctl);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    for (auto e : c->event_pool) cudaEventDestroy(e);
    for (auto& e : c->events) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
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

    c->cap_n = cfg->max_particles;
    c->cap_m = cfg->max_boundary;
    c->max_tiles = cfg->max_tiles ? cfg->max_tiles : cfg->max_particles / 32 + 4096;
    c->smem_optin = (size_t)prop.sharedMemPerBlockOptin - 1024;  // dynamic part; 1 KB is kept for the kernels' static shared memory
    c->lim_dyn = cfg->tile_dynamic_capacity;
    c->lim_stat = cfg->tile_static_capacity;
    if (c->lim_dyn > 65535u || c->lim_stat > 65535u) CREATE_FAIL(YASPH_ERR_INVALID_ARGUMENT, "tile capacities must be <= 65535 slots (16-bit slot indices)");
    if ((c->lim_dyn || c->lim_stat) && worst_smem_bytes(c->lim_dyn, c->lim_stat) > c->smem_optin)
        CREATE_FAIL(YASPH_ERR_CAPACITY, "tile capacities %u/%u need %zu bytes of shared memory per CTA, the device allows %zu", c->lim_dyn, c->lim_stat,
                    worst_smem_bytes(c->lim_dyn, c->lim_stat), c->smem_optin);
    if (c->cfg.speculative_iterations == 0) c->cfg.speculative_iterations = 2;
    c->cfg.max_tiles = c->max_tiles;

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
    c->tp.cfl_factor = cfg->cfl_factor;

    const size_t N = c->cap_n, M = c->cap_m, NM = N > M ? N : M;
    CUC(dmalloc(&c->pos, N));
    CUC(dmalloc(&c->pos_alt, N));
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
    CUC(dmalloc(&c->lists, N * (YASPH_MAXN / 4)));
    CUC(dmalloc(&c->counts, N));
    CUC(dmalloc(&c->radix_scratch, radix_scratch_words((uint32_t)NM)));
    CUC(dmalloc(&c->scan_chunks, (size_t)scan_num_chunks((uint32_t)NM) + 1));
    CUC(dmalloc(&c->scan_total, 1));
    CUC(dmalloc(&c->ctl, 1));
    CUC(cudaMallocHost((void**)&c->h_ctl, sizeof(Control)));
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
    CUC((prepare_sweep<OpViscosity>(c)));
    CUC((prepare_sweep<OpJacobiA<0>>(c)));
    CUC((prepare_sweep<OpJacobiA<1>>(c)));
    CUC((prepare_sweep<OpJacobiB<0, false>>(c)));
    CUC((prepare_sweep<OpJacobiB<0, true>>(c)));
    CUC((prepare_sweep<OpJacobiB<1, false>>(c)));
    CUC((prepare_sweep<OpJacobiB<1, true>>(c)));
    CUC((prepare_sweep<OpWcsphAccel>(c)));
    CUC(allow_max_smem(c, k_build_lists));
    CUC(dmalloc(&c->partials, (size_t)c->max_tiles + 1));

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
    if (n) *n = c->n;
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

// Stable LSD radix sort of (keys[0], idx[0]) over n elements; result back in buffer 0 (4 passes).  radix_prepare must be
// enqueued BEFORE the kernel that generates the keys, because that kernel accumulates the digit histograms.
static int32_t radix_prepare(yasph_ctx* c, uint32_t n) {
    if (n) CU(cudaMemsetAsync(c->radix_scratch, 0, radix_scratch_words(n) * sizeof(uint32_t), c->stream));
    return YASPH_OK;
}
static int32_t radix_sort(yasph_ctx* c, uint32_t n) {
    if (n == 0) return YASPH_OK;
    const uint32_t ntiles = radix_num_tiles(n);
    int src = 0;
    for (int pass = 0; pass < RS_PASSES; ++pass) {
        k_radix_pass<<<ntiles, RS_THREADS, sizeof(RadixPassSmem), c->stream>>>(c->keys[src], c->idx[src], c->keys[src ^ 1], c->idx[src ^ 1], n, pass,
                                                                             c->radix_scratch, ntiles);
        CHECK_LAUNCH();
        src ^= 1;
    }
    return YASPH_OK;
}

// cells (+ tiles for the dynamic grid) from the sorted keys in keys[0]
static int32_t build_cells(yasph_ctx* c, uint32_t n, bool is_static) {
    uint32_t* ck = is_static ? c->scell_key : c->cell_key;
    uint32_t* cs = is_static ? c->scell_start : c->cell_start;
    if (n) {
        const uint32_t nch = scan_num_chunks(n);
        HeadFlagsIn in{c->keys[0]};
        HeadCompactOut out{c->keys[0], ck, cs, is_static ? c->stile_key : c->tile_key, is_static ? nullptr : c->tile_pstart,
                           is_static ? c->stile_cstart : c->tile_cstart, is_static ? c->cap_m : c->max_tiles};
        k_scan_reduce<unsigned long long, HeadFlagsIn><<<nch, SCAN_THREADS, 0, c->stream>>>(in, n, c->scan_chunks);
        CHECK_LAUNCH();
        k_scan_chunks<unsigned long long><<<1, SCAN_THREADS, 0, c->stream>>>(c->scan_chunks, nch, c->scan_total);
        CHECK_LAUNCH();
        k_scan_apply<unsigned long long, HeadFlagsIn, HeadCompactOut><<<nch, SCAN_THREADS, 0, c->stream>>>(in, n, c->scan_chunks, out);
        CHECK_LAUNCH();
    }
    k_finish_cells<<<1, 32, 0, c->stream>>>(c->scan_total, n, ck, cs, c->tile_pstart, is_static ? c->stile_cstart : c->tile_cstart,
                                            is_static ? c->cap_m : c->max_tiles, c->ctl, is_static ? 1 : 0);
    CHECK_LAUNCH();
    return YASPH_OK;
}

static TileTables tile_tables(const yasph_ctx* c) { return TileTables{c->tile_runs, c->cslot_d, c->cslot_s}; }

static SweepCommon sweep_common(const yasph_ctx* c) {
    SweepCommon s;
    s.tt = tile_tables(c);
    s.lists = c->lists;
    s.counts = c->counts;
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
    return s;
}
// persistent grid of a tile kernel: every SM filled to the kernel's occupancy at this shared-memory size, at most one CTA per tile
template <class K>
static uint32_t persistent_grid(const yasph_ctx* c, K kernel, size_t smem_bytes) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TILE_THREADS, smem_bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
    const uint32_t g = (uint32_t)(c->num_sms * per_sm);
    return g < c->num_tiles ? g : (c->num_tiles ? c->num_tiles : 1u);
}
template <class Op>
static int32_t launch_sweep(yasph_ctx* c, Op op) {
    if (c->num_tiles == 0) return YASPH_OK;
    const size_t bytes = sweep_smem_bytes<Op>(c->cap_dyn, c->cap_stat);
    k_sweep<Op><<<persistent_grid(c, k_sweep<Op>, bytes), SW_THREADS, bytes, c->stream>>>(sweep_common(c), op);
    CHECK_LAUNCH();
    return YASPH_OK;
}

// copies the control block to the host (synchronises the stream)
static int32_t read_control(yasph_ctx* c) {
    CU(cudaMemcpyAsync(c->h_ctl, c->ctl, sizeof(Control), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}
static int32_t check_capacity_flags(yasph_ctx* c) {
    if (c->h_ctl->err_tile_count)
        return fail(c, YASPH_ERR_CAPACITY, "%u non-empty tiles exceed max_tiles=%u (particles too sparse for the configured capacity)", c->h_ctl->err_tile_count, c->max_tiles);
    if (c->h_ctl->err_tile_capacity)
        return fail(c, YASPH_ERR_CAPACITY, "a tile stages %u candidates; slots are 16-bit (<= 65535)", c->h_ctl->err_tile_capacity);
    return YASPH_OK;
}

// CompactMortonCellGrid::update for the dynamic particles + NeighborLists::update.
// `keys_ready`: keys[0]/idx[0] were already produced by a fused advect / kick kernel.
// gather2: float2 arrays permuted with the particles (pointer to the ctx member pair), gather1 likewise for float arrays.
struct GatherPlan {
    float2** a2[3];
    float2** alt2[3];
    int n2 = 0;
    float** a1[2];
    float** alt1[2];
    int n1 = 0;
};
static int32_t neighborhood_update(yasph_ctx* c, bool keys_ready, GatherPlan& gp) {
    const uint32_t n = c->n;
    c->lists_valid = false;
    pass_begin(c, YASPH_PASS_SORT);
    if (!keys_ready && n) {
        TRY(radix_prepare(c, n));
        k_keygen<<<blocks_for(n, RS_THREADS), RS_THREADS, 0, c->stream>>>(c->pos, n, c->grid, c->keys[0], c->idx[0], c->radix_scratch);
        CHECK_LAUNCH();
    }
    TRY(radix_sort(c, n));
    pass_end(c);
    pass_begin(c, YASPH_PASS_GATHER);
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
        k_gather<<<blocks_for(n, 256), 256, 0, c->stream>>>(c->idx[0], n, ga);
        CHECK_LAUNCH();
        for (int q = 0; q < gp.n2; ++q) std::swap(*gp.a2[q], *gp.alt2[q]);
        for (int q = 0; q < gp.n1; ++q) std::swap(*gp.a1[q], *gp.alt1[q]);
    }
    pass_end(c);
    pass_begin(c, YASPH_PASS_CELLS_TILES);
    TRY(build_cells(c, n, false));
    CU(cudaMemsetAsync(&c->ctl->total_neighbors, 0, sizeof(unsigned long long) + 2 * sizeof(unsigned int), c->stream));
    if (n) {
        TileTableArgs ta{c->tile_key, c->tile_pstart, c->tile_cstart, c->cell_key, c->cell_start, c->stile_key, c->stile_cstart,
                         c->scell_key, c->scell_start, c->tile_runs, c->cslot_d, c->cslot_s};
        k_tile_tables<<<c->num_sms * 16, TT_WARPS * 32, 0, c->stream>>>(ta, c->ctl);
        CHECK_LAUNCH();
    }
    pass_end(c);
    // The one host round trip of the neighbourhood update: tile count (grid of every tile kernel until the next update) and
    // the largest tile (their shared-memory size).
    TRY(read_control(c));
    TRY(check_capacity_flags(c));
    c->num_tiles = n ? c->h_ctl->num_tiles : 0u;
    c->cap_dyn = (c->h_ctl->max_dyn_total + 15u) & ~15u;
    c->cap_stat = (c->h_ctl->max_stat_total + 15u) & ~15u;
    if ((c->lim_dyn && c->h_ctl->max_dyn_total > c->lim_dyn) || (c->lim_stat && c->h_ctl->max_stat_total > c->lim_stat))
        return fail(c, YASPH_ERR_CAPACITY, "a tile stages %u dynamic / %u static candidates, above tile_dynamic_capacity=%u / tile_static_capacity=%u",
                    c->h_ctl->max_dyn_total, c->h_ctl->max_stat_total, c->lim_dyn, c->lim_stat);
    if (worst_smem_bytes(c->cap_dyn, c->cap_stat) > c->smem_optin)
        return fail(c, YASPH_ERR_CAPACITY, "a tile stages %u dynamic / %u static candidates: %zu bytes of shared memory per CTA, the device allows %zu",
                    c->h_ctl->max_dyn_total, c->h_ctl->max_stat_total, worst_smem_bytes(c->cap_dyn, c->cap_stat), c->smem_optin);
    pass_begin(c, YASPH_PASS_LISTS);
    if (c->num_tiles) {
        const size_t bytes = list_smem_bytes(c->cap_dyn, c->cap_stat);
        ListArgs la{tile_tables(c), c->pos, c->bpos, c->keys[0], c->grid, c->ctl, c->lists, c->counts, c->cap_dyn, c->cap_stat};
        k_build_lists<<<persistent_grid(c, k_build_lists, bytes), NB_THREADS, bytes, c->stream>>>(la);
        CHECK_LAUNCH();
    }
    pass_end(c);
    c->lists_valid = true;
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
    r->not_converged = h.not_converged;
    r->num_cells = h.num_cells;
    r->num_tiles = h.num_tiles;
    r->total_neighbors = h.total_neighbors;
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
        k_keygen<<<blocks_for(m, RS_THREADS), RS_THREADS, 0, c->stream>>>(c->bpos, m, c->grid, c->keys[0], c->idx[0], c->radix_scratch);
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
    if (n != c->n) c->dfsph_ready = false;  // dfsph.rs:419: alpha_values.len() != positions.len() -> re-initialise
    c->n = n;
    c->lists_valid = false;
    c->have_particles = true;
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
    const size_t n = c->n;
    if (n) {
        if (pos_xy) CU(cudaMemcpyAsync(pos_xy, c->pos, n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
        if (vel_xy) CU(cudaMemcpyAsync(vel_xy, c->vel, n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
        if (densities) CU(cudaMemcpyAsync(densities, c->dens, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}

extern "C" int32_t yasph_download_field(yasph_ctx* c, int32_t field, void* out, uint64_t out_bytes) {
    if (!c || !out) return YASPH_ERR_INVALID_ARGUMENT;
    CU(cudaSetDevice(c->device));
    const void* src = nullptr;
    size_t bytes = 0;
    const size_t n = c->n, m = c->m;
    switch (field) {
        case YASPH_FIELD_POSITION: src = c->pos; bytes = n * sizeof(float2); break;
        case YASPH_FIELD_VELOCITY: src = c->vel; bytes = n * sizeof(float2); break;
        case YASPH_FIELD_DENSITY: src = c->dens; bytes = n * sizeof(float); break;
        case YASPH_FIELD_ALPHA: src = c->alpha; bytes = n * sizeof(float); break;
        case YASPH_FIELD_KAPPA: src = c->kappa; bytes = n * sizeof(float); break;
        case YASPH_FIELD_STIFFNESS: src = c->stiff; bytes = n * sizeof(float); break;
        case YASPH_FIELD_ACCELERATION: src = c->accel; bytes = n * sizeof(float2); break;
        case YASPH_FIELD_CELL_KEY: src = c->keys[0]; bytes = n * sizeof(uint32_t); break;
        case YASPH_FIELD_SORT_PERMUTATION: src = c->idx[0]; bytes = n * sizeof(uint32_t); break;
        case YASPH_FIELD_BOUNDARY: src = c->bpos; bytes = m * sizeof(float2); break;
        default: return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_download_field: unknown field %d", field);
    }
    if (out_bytes < bytes) return fail(c, YASPH_ERR_INVALID_ARGUMENT, "yasph_download_field: buffer of %llu bytes < %zu", (unsigned long long)out_bytes, bytes);
    if (bytes) CU(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, c->stream));
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
extern "C" int32_t yasph_time_restart(yasph_ctx* c) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
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
template <int SOLVER>
static int32_t jacobi_solve(yasph_ctx* c) {
    const float rho0 = c->cfg.fluid_density;
    float* warm_arr = SOLVER == 0 ? c->kappa : c->stiff;
    pass_begin(c, SOLVER == 0 ? YASPH_PASS_DENSITY_WARM : YASPH_PASS_DIVERGENCE_WARM);
    // The warm start runs iff the previous solve took more than one iteration (dfsph.rs:199,354).  The host mirror of the
    // control block still holds that count (it is refreshed at every read-back and this solve has not started), so the
    // launch is skipped altogether when it would be a no-op; the kernel checks the device-side flag as well.
    if (c->h_ctl->iters[SOLVER] > 1u) {
        OpJacobiB<SOLVER, true> w;
        w.vstar = c->vstar;
        w.kfac = nullptr;
        w.warm = warm_arr;
        w.clamp_min = -0.5f * rho0 * rho0;
        w.iter_index = 0;
        TRY(launch_sweep(c, w));
    }
    pass_end(c);
    pass_begin(c, SOLVER == 0 ? YASPH_PASS_DENSITY_SOLVE : YASPH_PASS_DIVERGENCE_SOLVE);
    SolverParams sp;
    sp.max_error = SOLVER == 0 ? c->cfg.dfsph_max_avg_density_error : c->cfg.dfsph_max_divergence_error;
    sp.max_iters = SOLVER == 0 ? c->cfg.dfsph_max_density_iters : c->cfg.dfsph_max_divergence_iters;
    uint32_t it = 0;
    const uint32_t chunk = c->cfg.speculative_iterations;
    while (true) {
        for (uint32_t q = 0; q < chunk; ++q, ++it) {
            OpJacobiA<SOLVER> a;
            a.vstar = c->vstar;
            a.dens = c->dens;
            a.alpha = c->alpha;
            a.kfac = c->err_buf;
            a.sp = sp;
            a.iter_index = it;
            TRY(launch_sweep(c, a));
            OpJacobiB<SOLVER, false> b;
            b.vstar = c->vstar;
            b.kfac = c->err_buf;
            b.warm = warm_arr;
            b.clamp_min = 0.f;
            b.iter_index = it;
            TRY(launch_sweep(c, b));
        }
        TRY(read_control(c));
        if (c->h_ctl->stop_iter[SOLVER] != 0xFFFFFFFFu) break;
        if (it > sp.max_iters + chunk + 1) return fail(c, YASPH_ERR_STATE, "jacobi_solve: device loop control did not terminate");
    }
    pass_end(c);
    return YASPH_OK;
}

static int32_t dfsph_step(yasph_ctx* c) {
    const uint32_t n = c->n;
    if (!c->dfsph_ready) {
        // dfsph.rs:419-428: first call (or particle count changed): zero the warm-start arrays, sort, densities, alpha
        CU(cudaMemsetAsync(c->kappa, 0, (size_t)c->cap_n * sizeof(float), c->stream));
        CU(cudaMemsetAsync(c->stiff, 0, (size_t)c->cap_n * sizeof(float), c->stream));
        GatherPlan gp;
        gp.n2 = 2;
        gp.a2[0] = &c->pos;
        gp.alt2[0] = &c->pos_alt;
        gp.a2[1] = &c->vel;
        gp.alt2[1] = &c->vel_alt;
        TRY(neighborhood_update(c, false, gp));
        pass_begin(c, YASPH_PASS_DENSITY_ALPHA);
        OpDensityAlpha<0, true> da;
        da.dens = c->dens;
        da.alpha = c->alpha;
        da.rho_p = nullptr;
        da.stiffness = 0.f;
        TRY(launch_sweep(c, da));
        pass_end(c);
        c->dfsph_ready = true;
    }
    k_begin_step<<<1, 32, 0, c->stream>>>(c->ctl);
    CHECK_LAUNCH();
    // non-pressure forces + CFL maximum (dfsph.rs:436-477)
    pass_begin(c, YASPH_PASS_VISCOSITY);
    {
        OpViscosity v;
        v.vel = c->vel;
        v.dens = c->dens;
        v.accel = c->accel;
        const float2 g = make_float2(c->cfg.gravity[0], c->cfg.gravity[1]);
        v.base_accel = (g * c->mass) / c->mass;  // dfsph.rs:442-444
        v.vp = visc_params(c);
        v.dt = 0.f;
        TRY(launch_sweep(c, v));
    }
    pass_end(c);
    // update timestep + velocity prediction (dfsph.rs:478-491)
    pass_begin(c, YASPH_PASS_PREDICT);
    k_timestep_apply<0><<<blocks_for(n, 256), 256, 0, c->stream>>>(c->ctl, c->tp, c->radius * 2.0f, c->vel, c->accel, c->vstar, n);
    CHECK_LAUNCH();
    pass_end(c);
    TRY(jacobi_solve<0>(c));  // dfsph.rs:496
    // advect (dfsph.rs:502-509) fused with the key generation of the re-sort (dfsph.rs:512)
    pass_begin(c, YASPH_PASS_ADVECT_KEYGEN);
    TRY(radix_prepare(c, n));
    k_advect_keygen<<<blocks_for(n, RS_THREADS), RS_THREADS, 0, c->stream>>>(c->pos, c->vstar, n, c->ctl, c->grid, c->keys[0], c->idx[0], c->radix_scratch);
    CHECK_LAUNCH();
    pass_end(c);
    {
        // the reference also permutes the old velocities, which are discarded at the final swap (quirk Q7): skipped.
        GatherPlan gp;
        gp.n2 = 2;
        gp.a2[0] = &c->pos;
        gp.alt2[0] = &c->pos_alt;
        gp.a2[1] = &c->vstar;
        gp.alt2[1] = &c->vstar_alt;
        if (c->cfg.flags & YASPH_FLAG_PERMUTE_WARMSTART) {
            gp.n1 = 2;
            gp.a1[0] = &c->kappa;
            gp.alt1[0] = &c->f_alt0;
            gp.a1[1] = &c->stiff;
            gp.alt1[1] = &c->f_alt1;
        }
        TRY(neighborhood_update(c, true, gp));
    }
    pass_begin(c, YASPH_PASS_DENSITY_ALPHA);
    {
        OpDensityAlpha<0, true> da;  // dfsph.rs:516-518
        da.dens = c->dens;
        da.alpha = c->alpha;
        da.rho_p = nullptr;
        da.stiffness = 0.f;
        TRY(launch_sweep(c, da));
    }
    k_begin_divergence<<<1, 32, 0, c->stream>>>(c->ctl);
    CHECK_LAUNCH();
    pass_end(c);
    TRY(jacobi_solve<1>(c));        // dfsph.rs:521
    std::swap(c->vel, c->vstar);    // dfsph.rs:524
    return YASPH_OK;
}

static int32_t wcsph_step(yasph_ctx* c) {
    const uint32_t n = c->n;
    k_begin_step<<<1, 32, 0, c->stream>>>(c->ctl);
    CHECK_LAUNCH();
    // leap frog 1 (wscsph.rs:141-150) fused with key generation
    pass_begin(c, YASPH_PASS_ADVECT_KEYGEN);
    TRY(radix_prepare(c, n));
    k_kickdrift_keygen<<<blocks_for(n, RS_THREADS), RS_THREADS, 0, c->stream>>>(c->pos, c->vel, c->accel, n, c->ctl, c->grid, c->keys[0], c->idx[0], c->radix_scratch);
    CHECK_LAUNCH();
    pass_end(c);
    GatherPlan gp;
    gp.n2 = 2;
    gp.a2[0] = &c->pos;
    gp.alt2[0] = &c->pos_alt;
    gp.a2[1] = &c->vel;
    gp.alt2[1] = &c->vel_alt;
    TRY(neighborhood_update(c, true, gp));  // wscsph.rs:153
    pass_begin(c, YASPH_PASS_DENSITY_ALPHA);
    TRY((launch_density<1, true>(c)));  // Poly6, wscsph.rs:154; + Tait pressure per particle (wscsph.rs:91-92)
    pass_end(c);
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
    // update timestep + leap frog 2 (wscsph.rs:160-177)
    pass_begin(c, YASPH_PASS_WCSPH_KICK);
    k_timestep_apply<1><<<blocks_for(n, 256), 256, 0, c->stream>>>(c->ctl, c->tp, c->radius * 2.0f, c->vel, c->accel, c->vel, n);
    CHECK_LAUNCH();
    pass_end(c);
    return YASPH_OK;
}

extern "C" int32_t yasph_step(yasph_ctx* c, yasph_step_report* report) {
    if (!c) return YASPH_ERR_INVALID_ARGUMENT;
    if (!c->have_particles || c->n == 0) return fail(c, YASPH_ERR_STATE, "yasph_step: no particles uploaded");
    CU(cudaSetDevice(c->device));
    if (c->cfg.solver == YASPH_SOLVER_WCSPH)
        TRY(wcsph_step(c));
    else
        TRY(dfsph_step(c));
    TRY(read_control(c));
    pass_resolve(c);
    TRY(check_capacity_flags(c));
    fill_report(c, report);
    if (c->h_ctl->nonfinite) return fail(c, YASPH_ERR_NONFINITE, "non-finite Jacobi residual (solver mask %u)", c->h_ctl->nonfinite);
    return YASPH_OK;
}

extern "C" int32_t yasph_step_host(yasph_ctx* c, float* pos_xy, float* vel_xy, float* densities, uint32_t n, yasph_step_report* report) {
    if (!c || !pos_xy || !vel_xy) return YASPH_ERR_INVALID_ARGUMENT;
    if (n == 0 || n > c->cap_n) return fail(c, YASPH_ERR_CAPACITY, "yasph_step_host: n=%u out of range (max_particles=%u)", n, c->cap_n);
    CU(cudaSetDevice(c->device));
    if (n != c->n) c->dfsph_ready = false;
    c->n = n;
    c->have_particles = true;
    CU(cudaMemcpyAsync(c->pos, pos_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->vel, vel_xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
    int32_t rc = yasph_step(c, report);
    if (rc != YASPH_OK) return rc;
    CU(cudaMemcpyAsync(pos_xy, c->pos, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(vel_xy, c->vel, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, c->stream));
    if (densities) CU(cudaMemcpyAsync(densities, c->dens, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return YASPH_OK;
}

#include "scene.inl"
