// C ABI of the batched environment step (include/rocket_landing_b200.h).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "rl_kernels.cuh"
#include "rl_sim.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define RL_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return fail(RL_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                               \
    } while (0)


}  // namespace

struct RlEnv {
    rl::DevParams dev;
    RlParams params;
    rl::StatePlanes planes;
    void* state_buf = nullptr;
    bool owns_state = false;
    double* stats = nullptr;          // device, rl::kStatSlots x RL_STATS_WORDS
    int64_t n = 0;
    int64_t gid0 = 0;
    uint64_t seed = 0;
    unsigned long long* epoch = nullptr;   // device: {launch epoch}
    unsigned int* ticket = nullptr;        // device
    int device = 0;
    int sms = 0;
    bool folded = false;              // constants equal the reference defaults: use the constexpr kernels
    int step_grid[2][2] = {{0, 0}, {0, 0}};   // [auto_reset][single_step]: SMs x resident CTAs (persistent grid)
    int tiles_per_cta = -1;           // consecutive tiles per CTA of the step kernel; 0 = persistent grid, -1 = by launch size
    int64_t launches = 0;
    // host-buffer path (rl_step_host): device staging, allocated on first use
    void* h_stage = nullptr;
    size_t h_stage_bytes = 0;
    unsigned long long* h_epoch_snapshot = nullptr;    // inside h_stage
    void* h_pinned = nullptr;         // small batches: pinned, device-mapped host staging (same layout as h_stage)
    int64_t small_host = 16384;       // rl_step_host: at most this many rockets take the zero-copy path
    cudaStream_t h_streams[2] = {nullptr, nullptr};
    cudaEvent_t h_events[2] = {nullptr, nullptr};
    // Ordering between the two API families on one handle: the device API only enqueues on the
    // caller's stream, the host API runs on the handle's private streams.  Every device-API call
    // notes its stream here; a host-API call first makes its private stream wait for everything
    // enqueued on that stream so far (order_after_device_api), so "rl_step on my stream, then
    // rl_step_host / rl_get_state_host" needs no synchronisation by the caller.  The reverse order
    // is safe by construction: host-API calls return only when their work is complete.
    cudaStream_t last_stream = nullptr;
    bool last_stream_valid = false;
    cudaEvent_t order_event = nullptr;
};

namespace {

// grid of the helper kernels (reset / state copy): one CTA per 256 rockets, capped at 8 per SM
int grid_for(const RlEnv* e, int64_t n) {
    const int64_t ctas = (n + rl::kCtaThreads - 1) / rl::kCtaThreads;
    const int64_t cap = (int64_t)e->sms * 8;
    return (int)(ctas < cap ? (ctas > 0 ? ctas : 1) : cap);
}

using StepKernel = void (*)(const rl::DevParams, const rl::StepArgs);

// [folded constants][auto_reset][single_step][chunked grid]
StepKernel step_kernel_for(bool folded, bool auto_reset, bool single, bool chunked) {
    using rl::DefaultParams;
    using rl::DevParams;
    static const StepKernel table[2][2][2][2] = {
        {{{rl::step_kernel<false, false, false, DevParams>, rl::step_kernel<false, false, true, DevParams>},
          {rl::step_kernel<false, true, false, DevParams>, rl::step_kernel<false, true, true, DevParams>}},
         {{rl::step_kernel<true, false, false, DevParams>, rl::step_kernel<true, false, true, DevParams>},
          {rl::step_kernel<true, true, false, DevParams>, rl::step_kernel<true, true, true, DevParams>}}},
        {{{rl::step_kernel<false, false, false, DefaultParams>, rl::step_kernel<false, false, true, DefaultParams>},
          {rl::step_kernel<false, true, false, DefaultParams>, rl::step_kernel<false, true, true, DefaultParams>}},
         {{rl::step_kernel<true, false, false, DefaultParams>, rl::step_kernel<true, false, true, DefaultParams>},
          {rl::step_kernel<true, true, false, DefaultParams>, rl::step_kernel<true, true, true, DefaultParams>}}}};
    return table[folded ? 1 : 0][auto_reset ? 1 : 0][single ? 1 : 0][chunked ? 1 : 0];
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Enqueue the step of tiles [tile_begin, tile_end) of the handle's rockets (a tile = 256 rockets;
// all buffer pointers are those of rocket 0).  `epoch_src` is the epoch the launch reads; the launch
// advances the handle's epoch when it ends iff `bump`.  rl_step passes the whole range, the handle's
// epoch and bump = true; the pipelined host path steps chunks against a snapshot of the epoch and
// lets only the last chunk advance it, so that it draws exactly the resets rl_step would.
int launch_step(RlEnv* e, const float* actions, float* obs, float* reward, uint8_t* terminated,
                uint8_t* truncated, float* terminal_obs, float* episode_return, int32_t* episode_length,
                int8_t* landing, float* raw_accel, int substeps, int auto_reset, cudaStream_t stream, int64_t tile_begin = 0,
                int64_t tile_end = -1, const unsigned long long* epoch_src = nullptr, bool bump = true) {
    rl::StepArgs a;
    const int64_t all_tiles = (e->n + rl::kCtaThreads - 1) / rl::kCtaThreads;
    if (tile_end < 0) tile_end = all_tiles;
    a.tile_begin = (uint32_t)tile_begin;
    a.tile_end = (uint32_t)tile_end;
    a.epoch_bump = bump ? e->epoch : nullptr;
    a.st = e->planes;
    a.actions = reinterpret_cast<const float2*>(actions);
    a.obs = reinterpret_cast<float4*>(obs);
    a.reward = reward;
    a.terminated = terminated;
    a.truncated = truncated;
    a.terminal_obs = reinterpret_cast<float4*>(terminal_obs);
    a.episode_return = episode_return;
    a.episode_length = episode_length;
    a.landing = landing;
    a.raw_accel = reinterpret_cast<float2*>(raw_accel);
    a.optional = (landing ? 1u : 0u) | (raw_accel ? 2u : 0u);
    a.stats = e->stats;
    a.n = (uint32_t)e->n;
    a.gid0 = e->gid0;
    a.seed = e->seed;
    a.epoch = epoch_src ? epoch_src : e->epoch;
    a.ticket = e->ticket;
    a.substeps = substeps;
    const bool single = substeps == 1;
    const int64_t ctas = tile_end - tile_begin;                              // one tile of 256 rockets each
    if (ctas <= 0) return RL_OK;
    const int cap = e->step_grid[auto_reset ? 1 : 0][single ? 1 : 0];
    // Grid shape (profiles/README.md, "grid order"): large launches run `chunk` consecutive tiles per
    // CTA, CTAs dispatched in address order by the hardware, which keeps the set of DRAM pages being
    // streamed compact (+20 % bandwidth over a persistent strided grid at 64 Mi rockets); that needs
    // at least ~6 waves of CTAs to balance, so smaller launches keep the persistent grid.
    int chunk = e->tiles_per_cta;
    if (chunk < 0) {
        chunk = 0;
        for (int c : {8, 4})
            if (ctas >= (int64_t)6 * cap * c) { chunk = c; break; }
    }
    a.chunk = (uint32_t)chunk;
    const int grid = a.chunk ? (int)((ctas + a.chunk - 1) / a.chunk) : (int)(ctas < cap ? ctas : cap);
    RL_CUDA(rl::rl_launch(step_kernel_for(e->folded, auto_reset != 0, single, chunk != 0), dim3(grid), dim3(rl::kCtaThreads), 0, stream,
                          e->dev, a));
    e->launches += 1;
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

bool misaligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) != 0; }

int ensure_streams(RlEnv* e) {
    if (e->h_streams[0]) return RL_OK;
    for (int k = 0; k < 2; ++k) {
        RL_CUDA(cudaStreamCreateWithFlags(&e->h_streams[k], cudaStreamNonBlocking));
        RL_CUDA(cudaEventCreateWithFlags(&e->h_events[k], cudaEventDisableTiming));
    }
    RL_CUDA(cudaEventCreateWithFlags(&e->order_event, cudaEventDisableTiming));
    return RL_OK;
}

// Make `hs` (a private host-path stream) wait for everything the device API has enqueued on the
// caller's stream so far.  A stream that is being captured or has been destroyed cannot be waited
// for: the error is cleared and the call proceeds unordered (the caller synchronises in that case).
int order_after_device_api(RlEnv* e, cudaStream_t hs) {
    if (!e->last_stream_valid) return RL_OK;
    if (cudaEventRecord(e->order_event, e->last_stream) != cudaSuccess) {
        cudaGetLastError();
        e->last_stream_valid = false;
        return RL_OK;
    }
    RL_CUDA(cudaStreamWaitEvent(hs, e->order_event, 0));
    return RL_OK;
}

inline void note_stream(RlEnv* e, void* stream) {
    e->last_stream = (cudaStream_t)stream;
    e->last_stream_valid = true;
}

}  // namespace

extern "C" {

int rl_abi_version(void) { return RL_ABI_VERSION; }

const char* rl_last_error(void) { return g_err; }

void rl_params_default(RlParams* p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->time_step = 0.1;
    p->gravity = -9.81;
    p->air_density = 1.225;
    p->thrust_power = 1.0e7;
    p->cold_gas_thrust_power = 7.0e4;
    p->fuel_consumption_rate = 1700.0;
    p->radius = 1.85;
    p->cold_gas_moment_arm = 1.85;
    p->reference_area = 10.8;
    p->drag_coefficient = 0.8;
    p->angular_damping = 0.05;
    const double lo[10] = {-1500, 1800, -25, -200, -5, -5, -15, -10, 34000, 370000};
    const double hi[10] = {1500, 2200, 25, -150, 5, 5, 15, 10, 38000, 410000};
    memcpy(p->reset_low, lo, sizeof(lo));
    memcpy(p->reset_high, hi, sizeof(hi));
    const double perfect[3] = {30, 30, 5}, good[3] = {40, 40, 8}, ok[3] = {100, 100, 10};
    memcpy(p->land_perfect, perfect, sizeof(perfect));
    memcpy(p->land_good, good, sizeof(good));
    memcpy(p->land_ok, ok, sizeof(ok));
    p->max_horizontal_position = 50000.0;
    p->max_altitude = 50000.0;
    p->tip_over_angle = 90.0;
    p->landing_perfect = 4000.0;
    p->landing_good = 3000.0;
    p->landing_ok = 2000.0;
    p->crash_ground = -800.0;
    p->out_of_bounds = -300.0;
    p->tipped_over = -300.0;
    p->throttle_descent_reward_scale = 0.3;
    p->cold_gas_reward_scale = 0.7;
    p->free_fall_penalty_scale = 0.2;
    p->angle_aware_throttle_scale = 0.3;
    p->correct_direction_bonus = 1.5;
    p->gamma = 0.99;
    // lander.py:73-101: [x, y, vx, vy, ax, ay, angle, omega]; y low is the literal 0.0
    const int src[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    for (int k = 0; k < 8; ++k) {
        p->obs_low[k] = lo[src[k]];
        p->obs_high[k] = hi[src[k]];
    }
    p->obs_low[1] = 0.0;
    p->max_episode_steps = 1000;
}

int64_t rl_state_bytes(int64_t n_envs) { return n_envs > 0 ? 40 * rl::padded_envs(n_envs) : 0; }

int rl_create(const RlParams* params, int64_t n_envs, int64_t env_id_offset, uint64_t seed, int device,
              void* state_buf, RlEnv** out) {
    if (!params || !out) return fail(RL_E_INVALID, "rl_create: null argument");
    if (n_envs <= 0) return fail(RL_E_INVALID, "rl_create: n_envs must be positive (got %lld)", (long long)n_envs);
    if (n_envs > 2147483392LL)    // 32-bit rocket index inside a shard (2^31 - 256); shard larger batches
        return fail(RL_E_INVALID, "rl_create: at most 2147483392 rockets per handle (got %lld)", (long long)n_envs);
    if (env_id_offset < 0) return fail(RL_E_INVALID, "rl_create: env_id_offset must be >= 0");
    if (state_buf && misaligned(state_buf, 256)) return fail(RL_E_INVALID, "rl_create: state_buf must be 256-byte aligned");
    RlEnv* e = new (std::nothrow) RlEnv();
    if (!e) return fail(RL_E_NOMEM, "rl_create: out of host memory");
    int rc = rl::derive_params(*params, e->dev, g_err, sizeof(g_err));
    if (rc != RL_OK) { delete e; return rc; }
    e->params = *params;
    e->n = n_envs;
    e->gid0 = env_id_offset;
    e->seed = seed;
    e->device = device;
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || device < 0 || device >= count) {
        delete e;
        return fail(RL_E_CUDA, "rl_create: CUDA device %d not available (%s)", device,
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "bad ordinal");
    }
    DeviceGuard guard(device);
    if (!guard.ok) { delete e; return fail(RL_E_CUDA, "rl_create: cudaSetDevice(%d) failed", device); }
    const size_t bytes = (size_t)rl_state_bytes(n_envs);
    if (state_buf) {
        e->state_buf = state_buf;
    } else {
        if (cudaMalloc(&e->state_buf, bytes) != cudaSuccess) {
            delete e;
            return fail(RL_E_NOMEM, "rl_create: cudaMalloc(%zu) failed", bytes);
        }
        e->owns_state = true;
    }
    e->planes = rl::carve_state(e->state_buf, n_envs);
    if (cudaMalloc(&e->stats, rl::kStatSlots * RL_STATS_WORDS * sizeof(double)) != cudaSuccess ||
        cudaMemset(e->stats, 0, rl::kStatSlots * RL_STATS_WORDS * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&e->epoch, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(e->epoch, 0, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&e->ticket, sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(e->ticket, 0, sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(e->state_buf, 0, bytes) != cudaSuccess) {
        rl_destroy(e);
        return fail(RL_E_CUDA, "rl_create: device initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    // Constants equal to the reference's config.yaml run the instantiation with them folded in as
    // immediates; anything else (or RL_B200_GENERIC_KERNEL=1, used by the tests) runs the
    // instantiation that reads them from the kernel-parameter constant bank.
    const char* force_generic = getenv("RL_B200_GENERIC_KERNEL");
    e->folded = rl::matches_default(e->dev) && !(force_generic && force_generic[0] == '1');
    cudaDeviceGetAttribute(&e->sms, cudaDevAttrMultiProcessorCount, device);
    if (const char* t = getenv("RL_B200_SMALL_HOST")) {           // experiment knob: zero-copy host path up to this size
        const long long v = strtoll(t, nullptr, 10);
        if (v >= 0) e->small_host = v;
    }
    if (const char* t = getenv("RL_B200_TILES_PER_CTA")) {      // experiment knob: force the grid shape
        const long v = strtol(t, nullptr, 10);
        if (v >= -1 && v <= 65536) e->tiles_per_cta = (int)v;
    }
    for (int ar = 0; ar < 2; ++ar)
        for (int single = 0; single < 2; ++single) {
            int per_sm = 0;
            // each CTA stages its warps' tiles in ~33 KiB of static shared memory: ask for the largest
            // carve-out so that registers, not shared memory, bound the resident CTAs
            for (int chunked = 1; chunked >= 0; --chunked) {     // (both grid shapes have the same footprint)
                StepKernel fn = step_kernel_for(e->folded, ar != 0, single != 0, chunked != 0);
                cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, rl::kCtaThreads, 0);
            }
            if (e->sms <= 0 || per_sm <= 0) {
                rl_destroy(e);
                return fail(RL_E_CUDA, "rl_create: occupancy query failed (no sm_100a image for this device?): %s",
                            cudaGetErrorString(cudaGetLastError()));
            }
            e->step_grid[ar][single] = e->sms * per_sm;
        }
    // the zero-fills above ran on the legacy default stream, which a caller's non-blocking stream
    // (and the handle's private host-path streams) does not wait for: finish them here, once
    if (cudaDeviceSynchronize() != cudaSuccess) {
        rl_destroy(e);
        return fail(RL_E_CUDA, "rl_create: device initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    *out = e;
    return RL_OK;
}

int rl_destroy(RlEnv* e) {
    if (!e) return RL_OK;
    DeviceGuard guard(e->device);
    if (e->owns_state && e->state_buf) cudaFree(e->state_buf);
    if (e->stats) cudaFree(e->stats);
    if (e->epoch) cudaFree(e->epoch);
    if (e->ticket) cudaFree(e->ticket);
    if (e->h_stage) cudaFree(e->h_stage);
    if (e->h_pinned) cudaFreeHost(e->h_pinned);
    if (e->order_event) cudaEventDestroy(e->order_event);
    for (int k = 0; k < 2; ++k) {
        if (e->h_streams[k]) cudaStreamDestroy(e->h_streams[k]);
        if (e->h_events[k]) cudaEventDestroy(e->h_events[k]);
    }
    delete e;
    return RL_OK;
}

int64_t rl_num_envs(const RlEnv* e) { return e ? e->n : 0; }
double rl_fuel_unit(const RlEnv* e) { return e ? e->dev.fuel_unit_d : 0.0; }
double rl_fuel_unit_for(const RlParams* p) { return p ? rl::fuel_unit_for(p->reset_high[9]) : 0.0; }
double rl_pos_unit(const RlEnv* e) { return e ? (double)e->dev.pos_unit : 0.0; }
double rl_pos_unit_for(const RlParams* p) { return p ? rl::pos_unit_for(*p) : 0.0; }
int rl_uses_folded_constants(const RlEnv* e) { return (e && e->folded) ? 1 : 0; }
uint64_t rl_get_epoch(RlEnv* e) {
    if (!e) return 0;
    DeviceGuard guard(e->device);
    unsigned long long v = 0;
    // ordered after everything this handle has enqueued (device API: the caller's last stream; host
    // API: already complete); no device-wide synchronisation
    if (ensure_streams(e) != RL_OK || order_after_device_api(e, e->h_streams[0]) != RL_OK ||
        cudaMemcpyAsync(&v, e->epoch, sizeof(v), cudaMemcpyDeviceToHost, e->h_streams[0]) != cudaSuccess ||
        cudaStreamSynchronize(e->h_streams[0]) != cudaSuccess) {
        fail(RL_E_CUDA, "rl_get_epoch: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    return (uint64_t)v;
}
int rl_set_epoch(RlEnv* e, uint64_t epoch) {
    if (!e) return fail(RL_E_INVALID, "rl_set_epoch: null handle");
    DeviceGuard guard(e->device);
    const unsigned long long v = epoch;
    int rc = ensure_streams(e);
    if (rc != RL_OK) return rc;
    rc = order_after_device_api(e, e->h_streams[0]);
    if (rc != RL_OK) return rc;
    RL_CUDA(cudaMemcpyAsync(e->epoch, &v, sizeof(v), cudaMemcpyHostToDevice, e->h_streams[0]));
    RL_CUDA(cudaStreamSynchronize(e->h_streams[0]));       // later device-API calls are enqueued after this returns
    return RL_OK;
}
int64_t rl_launch_count(const RlEnv* e) { return e ? e->launches : 0; }

int rl_reset(RlEnv* e, const uint8_t* mask, float* obs, void* stream) {
    if (!e) return fail(RL_E_INVALID, "rl_reset: null handle");
    if (obs && misaligned(obs, 16)) return fail(RL_E_INVALID, "rl_reset: obs must be 16-byte aligned");
    DeviceGuard guard(e->device);
    note_stream(e, stream);
    rl::reset_kernel<<<grid_for(e, e->n), rl::kCtaThreads, 0, (cudaStream_t)stream>>>(
        e->dev, e->planes, mask, reinterpret_cast<float4*>(obs), (uint32_t)e->n, e->gid0, e->seed, e->epoch, e->ticket);
    e->launches += 1;
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

int rl_step(RlEnv* e, const float* actions, float* obs, float* reward, uint8_t* terminated, uint8_t* truncated,
            float* terminal_obs, float* episode_return, int32_t* episode_length, int8_t* landing, float* raw_accel,
            int substeps, int auto_reset, void* stream) {
    if (!e) return fail(RL_E_INVALID, "rl_step: null handle");
    if (!actions || !obs || !reward || !terminated || !truncated)
        return fail(RL_E_INVALID, "rl_step: actions, obs, reward, terminated and truncated are required");
    if (substeps < 1 || substeps > 4096) return fail(RL_E_INVALID, "rl_step: substeps must be in [1, 4096] (got %d)", substeps);
    if (misaligned(actions, 8) || misaligned(obs, 16) || misaligned(reward, 4) ||
        (terminal_obs && misaligned(terminal_obs, 16)) || (raw_accel && misaligned(raw_accel, 8)))
        return fail(RL_E_INVALID, "rl_step: actions / raw_accel need 8-byte, obs / terminal_obs 16-byte alignment");
    DeviceGuard guard(e->device);
    note_stream(e, stream);
    return launch_step(e, actions, obs, reward, terminated, truncated, terminal_obs, episode_return,
                       episode_length, landing, raw_accel, substeps, auto_reset, (cudaStream_t)stream);
}

int rl_set_state(RlEnv* e, const uint32_t* state_soa, const uint8_t* mask, void* stream) {
    if (!e || !state_soa) return fail(RL_E_INVALID, "rl_set_state: null argument");
    DeviceGuard guard(e->device);
    note_stream(e, stream);
    rl::set_state_kernel<<<grid_for(e, e->n), rl::kCtaThreads, 0, (cudaStream_t)stream>>>(e->planes, state_soa, mask, (uint32_t)e->n);
    e->launches += 1;
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

int rl_get_state(RlEnv* e, uint32_t* state_soa, void* stream) {
    if (!e || !state_soa) return fail(RL_E_INVALID, "rl_get_state: null argument");
    DeviceGuard guard(e->device);
    note_stream(e, stream);
    rl::get_state_kernel<<<grid_for(e, e->n), rl::kCtaThreads, 0, (cudaStream_t)stream>>>(e->planes, state_soa, (uint32_t)e->n);
    e->launches += 1;
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

int rl_stats(RlEnv* e, double* out16, int reset_after, void* stream) {
    if (!e || !out16) return fail(RL_E_INVALID, "rl_stats: null argument");
    DeviceGuard guard(e->device);
    note_stream(e, stream);
    rl::stats_copy_kernel<<<1, RL_STATS_WORDS * 32, 0, (cudaStream_t)stream>>>(e->stats, out16, reset_after);
    e->launches += 1;
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

// ---- simulation mode (controller.py / controls.py / protocol.py) ------------------------------
int rl_sim_step(RlEnv* e, const float* actions, float* telemetry, float* reward, uint8_t* done, void* stream) {
    if (!e || !actions || !telemetry) return fail(RL_E_INVALID, "rl_sim_step: actions and telemetry are required");
    if (misaligned(actions, 8) || misaligned(telemetry, 16)) return fail(RL_E_INVALID, "rl_sim_step: actions need 8-byte, telemetry 16-byte alignment");
    DeviceGuard guard(e->device);
    note_stream(e, stream);
    if (e->folded)
        rl::sim_step_kernel<rl::DefaultParams><<<grid_for(e, e->n), rl::kCtaThreads, 0, (cudaStream_t)stream>>>(
            e->dev, e->planes, reinterpret_cast<const float2*>(actions), reinterpret_cast<float4*>(telemetry), reward, done, (uint32_t)e->n);
    else
        rl::sim_step_kernel<rl::DevParams><<<grid_for(e, e->n), rl::kCtaThreads, 0, (cudaStream_t)stream>>>(
            e->dev, e->planes, reinterpret_cast<const float2*>(actions), reinterpret_cast<float4*>(telemetry), reward, done, (uint32_t)e->n);
    e->launches += 1;
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

// ---- host-buffer path ------------------------------------------------------------------
// Device staging for one full batch: actions | obs | reward | terminated | truncated |
// terminal_obs | episode_return | episode_length | landing | raw_accel, each 256-byte aligned.
// rl_step_host cuts the batch into chunks and runs them as a two-stream software pipeline:
//   chunk c:  H2D(actions[c]) -> step(tiles of c) -> D2H(obs[c], reward[c], flags[c], ...)
// so that the (dominant) device-to-host copy of chunk c overlaps the host-to-device copy and the
// kernel of chunk c + 1 (PCIe is full duplex).  Every chunk reads the same SNAPSHOT of the launch
// epoch and only the last one advances it: the call draws exactly the resets one rl_step would.
namespace {
struct Stage {
    float *actions, *obs, *reward, *terminal_obs, *episode_return, *raw_accel;
    uint8_t *terminated, *truncated;
    int32_t* episode_length;
    int8_t* landing;
};
constexpr int kStageArrays = 10;
size_t align256(size_t v) { return (v + 255) / 256 * 256; }
size_t stage_sizes(const RlEnv* e, size_t sz[kStageArrays]) {
    const size_t n = (size_t)e->n;
    const size_t v[kStageArrays] = {align256(8 * n), align256(32 * n), align256(4 * n), align256(n), align256(n),
                                    align256(32 * n), align256(4 * n), align256(4 * n), align256(n), align256(8 * n)};
    size_t total = 256;                                   // + the epoch snapshot
    for (int k = 0; k < kStageArrays; ++k) { sz[k] = v[k]; total += v[k]; }
    return total;
}
void carve_stage(char* p, const size_t sz[kStageArrays], Stage& s) {
    p += 256;
    s.actions = (float*)p; p += sz[0];
    s.obs = (float*)p; p += sz[1];
    s.reward = (float*)p; p += sz[2];
    s.terminated = (uint8_t*)p; p += sz[3];
    s.truncated = (uint8_t*)p; p += sz[4];
    s.terminal_obs = (float*)p; p += sz[5];
    s.episode_return = (float*)p; p += sz[6];
    s.episode_length = (int32_t*)p; p += sz[7];
    s.landing = (int8_t*)p; p += sz[8];
    s.raw_accel = (float*)p;
}
int ensure_stage(RlEnv* e, Stage& s) {
    size_t sz[kStageArrays];
    const size_t total = stage_sizes(e, sz);
    if (!e->h_stage) {
        if (cudaMalloc(&e->h_stage, total) != cudaSuccess) return fail(RL_E_NOMEM, "host path: cudaMalloc(%zu) failed", total);
        e->h_stage_bytes = total;
    }
    int rc = ensure_streams(e);
    if (rc != RL_OK) return rc;
    e->h_epoch_snapshot = reinterpret_cast<unsigned long long*>(e->h_stage);
    carve_stage(static_cast<char*>(e->h_stage), sz, s);
    return RL_OK;
}
// Small batches: pinned host memory the device reads and writes directly (zero-copy over PCIe).
// `host` are the CPU's pointers, `dev` the device's aliases of the same bytes.
int ensure_pinned(RlEnv* e, Stage& host, Stage& dev) {
    size_t sz[kStageArrays];
    const size_t total = stage_sizes(e, sz);
    if (!e->h_pinned) {
        if (cudaHostAlloc(&e->h_pinned, total, cudaHostAllocMapped) != cudaSuccess)
            return fail(RL_E_NOMEM, "host path: cudaHostAlloc(%zu) failed", total);
        memset(e->h_pinned, 0, total);
    }
    int rc = ensure_streams(e);
    if (rc != RL_OK) return rc;
    void* d = nullptr;
    RL_CUDA(cudaHostGetDevicePointer(&d, e->h_pinned, 0));
    carve_stage(static_cast<char*>(e->h_pinned), sz, host);
    carve_stage(static_cast<char*>(d), sz, dev);
    return RL_OK;
}
}  // namespace

int rl_step_host(RlEnv* e, const float* actions, float* obs, float* reward, uint8_t* terminated, uint8_t* truncated,
                 float* terminal_obs, float* episode_return, int32_t* episode_length, int8_t* landing, float* raw_accel,
                 int substeps, int auto_reset) {
    if (!e) return fail(RL_E_INVALID, "rl_step_host: null handle");
    if (!actions || !obs || !reward || !terminated || !truncated)
        return fail(RL_E_INVALID, "rl_step_host: actions, obs, reward, terminated and truncated are required");
    if (substeps < 1 || substeps > 4096) return fail(RL_E_INVALID, "rl_step_host: substeps must be in [1, 4096]");
    DeviceGuard guard(e->device);
    const int64_t n = e->n;
    if (n <= e->small_host) {
        // Small batch (the reference trains with rl.n_envs = 8, config.yaml:95): latency, not bandwidth.
        // The kernel reads the actions from and writes its outputs to pinned host memory directly, so
        // the call is one launch and one stream synchronisation instead of six copies around it.
        Stage h, d;
        int rc = ensure_pinned(e, h, d);
        if (rc != RL_OK) return rc;
        rc = order_after_device_api(e, e->h_streams[0]);
        if (rc != RL_OK) return rc;
        const size_t m = (size_t)n;
        memcpy(h.actions, actions, 8 * m);
        rc = launch_step(e, d.actions, d.obs, d.reward, d.terminated, d.truncated, terminal_obs ? d.terminal_obs : nullptr,
                         episode_return ? d.episode_return : nullptr, episode_length ? d.episode_length : nullptr,
                         landing ? d.landing : nullptr, raw_accel ? d.raw_accel : nullptr, substeps, auto_reset,
                         e->h_streams[0]);
        if (rc != RL_OK) return rc;
        RL_CUDA(cudaStreamSynchronize(e->h_streams[0]));
        memcpy(obs, h.obs, 32 * m);
        memcpy(reward, h.reward, 4 * m);
        memcpy(terminated, h.terminated, m);
        memcpy(truncated, h.truncated, m);
        if (terminal_obs) memcpy(terminal_obs, h.terminal_obs, 32 * m);      // (only rows of done rockets are meaningful)
        if (episode_return) memcpy(episode_return, h.episode_return, 4 * m);
        if (episode_length) memcpy(episode_length, h.episode_length, 4 * m);
        if (landing) memcpy(landing, h.landing, m);
        if (raw_accel) memcpy(raw_accel, h.raw_accel, 8 * m);
        return RL_OK;
    }
    Stage s;
    int rc = ensure_stage(e, s);
    if (rc != RL_OK) return rc;
    rc = order_after_device_api(e, e->h_streams[0]);
    if (rc != RL_OK) return rc;
    const int64_t tiles = (n + rl::kCtaThreads - 1) / rl::kCtaThreads;
    // chunks of >= 1 Mi rockets (PCIe-efficient, and a full persistent grid), at most 16 of them
    int64_t chunks = tiles / 4096;
    chunks = chunks < 1 ? 1 : (chunks > 16 ? 16 : chunks);
    const int64_t tiles_per_chunk = (tiles + chunks - 1) / chunks;
    // snapshot of the launch epoch, ordered before every chunk
    RL_CUDA(cudaMemcpyAsync(e->h_epoch_snapshot, e->epoch, sizeof(unsigned long long), cudaMemcpyDeviceToDevice,
                            e->h_streams[0]));
    RL_CUDA(cudaEventRecord(e->h_events[0], e->h_streams[0]));
    RL_CUDA(cudaStreamWaitEvent(e->h_streams[1], e->h_events[0], 0));
    for (int64_t c = 0; c < chunks; ++c) {
        const int64_t t0 = c * tiles_per_chunk, t1 = (t0 + tiles_per_chunk < tiles) ? t0 + tiles_per_chunk : tiles;
        if (t0 >= t1) break;
        const size_t lo = (size_t)t0 * rl::kCtaThreads;
        const size_t hi = (size_t)t1 * rl::kCtaThreads < (size_t)n ? (size_t)t1 * rl::kCtaThreads : (size_t)n;
        const size_t m = hi - lo;
        cudaStream_t st = e->h_streams[c & 1];
        const bool last = (t1 == tiles);
        RL_CUDA(cudaMemcpyAsync(s.actions + 2 * lo, actions + 2 * lo, 8 * m, cudaMemcpyHostToDevice, st));
        rc = launch_step(e, s.actions, s.obs, s.reward, s.terminated, s.truncated,
                         terminal_obs ? s.terminal_obs : nullptr, episode_return ? s.episode_return : nullptr,
                         episode_length ? s.episode_length : nullptr, landing ? s.landing : nullptr,
                         raw_accel ? s.raw_accel : nullptr, substeps, auto_reset, st, t0, t1, e->h_epoch_snapshot, last);
        if (rc != RL_OK) return rc;
        RL_CUDA(cudaMemcpyAsync(obs + 8 * lo, s.obs + 8 * lo, 32 * m, cudaMemcpyDeviceToHost, st));
        RL_CUDA(cudaMemcpyAsync(reward + lo, s.reward + lo, 4 * m, cudaMemcpyDeviceToHost, st));
        RL_CUDA(cudaMemcpyAsync(terminated + lo, s.terminated + lo, m, cudaMemcpyDeviceToHost, st));
        RL_CUDA(cudaMemcpyAsync(truncated + lo, s.truncated + lo, m, cudaMemcpyDeviceToHost, st));
        if (terminal_obs) RL_CUDA(cudaMemcpyAsync(terminal_obs + 8 * lo, s.terminal_obs + 8 * lo, 32 * m, cudaMemcpyDeviceToHost, st));
        if (episode_return) RL_CUDA(cudaMemcpyAsync(episode_return + lo, s.episode_return + lo, 4 * m, cudaMemcpyDeviceToHost, st));
        if (episode_length) RL_CUDA(cudaMemcpyAsync(episode_length + lo, s.episode_length + lo, 4 * m, cudaMemcpyDeviceToHost, st));
        if (landing) RL_CUDA(cudaMemcpyAsync(landing + lo, s.landing + lo, m, cudaMemcpyDeviceToHost, st));
        if (raw_accel) RL_CUDA(cudaMemcpyAsync(raw_accel + 2 * lo, s.raw_accel + 2 * lo, 8 * m, cudaMemcpyDeviceToHost, st));
    }
    RL_CUDA(cudaStreamSynchronize(e->h_streams[0]));
    RL_CUDA(cudaStreamSynchronize(e->h_streams[1]));
    return RL_OK;
}

// One tick with HOST buffers: `actions` float32 [N,2]; `message` receives the complete binary
// telemetry message of protocol.py: 1 header byte (1 = telemetry) + 64 bytes per rocket.
int rl_sim_step_host(RlEnv* e, const float* actions, uint8_t* message) {
    if (!e || !actions || !message) return fail(RL_E_INVALID, "rl_sim_step_host: null argument");
    DeviceGuard guard(e->device);
    Stage s;
    int rc = ensure_stage(e, s);
    if (rc != RL_OK) return rc;
    const size_t n = (size_t)e->n;
    cudaStream_t st = e->h_streams[0];
    rc = order_after_device_api(e, st);
    if (rc != RL_OK) return rc;
    // The 64 B / rocket of records occupy the staging span that starts at the obs area (obs 32 n +
    // reward 4 n + flags 2 n + terminal_obs 32 n >= 64 n bytes, contiguous); the gym-mode host path
    // and this one never run concurrently on a handle.
    float* records = s.obs;
    RL_CUDA(cudaMemcpyAsync(s.actions, actions, 8 * n, cudaMemcpyHostToDevice, st));
    rc = rl_sim_step(e, s.actions, records, nullptr, nullptr, st);
    if (rc != RL_OK) return rc;
    message[0] = 1;                   // BinaryProtocol.MSG_TELEMETRY
    RL_CUDA(cudaMemcpyAsync(message + 1, records, 64 * n, cudaMemcpyDeviceToHost, st));
    RL_CUDA(cudaStreamSynchronize(st));
    return RL_OK;
}

int rl_reset_host(RlEnv* e, float* obs) {
    if (!e) return fail(RL_E_INVALID, "rl_reset_host: null handle");
    DeviceGuard guard(e->device);
    if (e->n <= e->small_host) {                       // small batch: the kernel writes the pinned host buffer itself
        Stage h, d;
        int rc = ensure_pinned(e, h, d);
        if (rc != RL_OK) return rc;
        rc = order_after_device_api(e, e->h_streams[0]);
        if (rc != RL_OK) return rc;
        rc = rl_reset(e, nullptr, obs ? d.obs : nullptr, e->h_streams[0]);
        if (rc != RL_OK) return rc;
        RL_CUDA(cudaStreamSynchronize(e->h_streams[0]));
        if (obs) memcpy(obs, h.obs, 32 * (size_t)e->n);
        return RL_OK;
    }
    Stage s;
    int rc = ensure_stage(e, s);
    if (rc != RL_OK) return rc;
    rc = order_after_device_api(e, e->h_streams[0]);
    if (rc != RL_OK) return rc;
    rc = rl_reset(e, nullptr, obs ? s.obs : nullptr, e->h_streams[0]);
    if (rc != RL_OK) return rc;
    if (obs) RL_CUDA(cudaMemcpyAsync(obs, s.obs, 32 * (size_t)e->n, cudaMemcpyDeviceToHost, e->h_streams[0]));
    RL_CUDA(cudaStreamSynchronize(e->h_streams[0]));
    return RL_OK;
}

// rl_get_state into HOST memory (uint32 [10][N], RL_W_*); synchronous.  Small batches go through
// pinned, device-mapped memory (one launch + one synchronisation).
int rl_get_state_host(RlEnv* e, uint32_t* state_soa) {
    if (!e || !state_soa) return fail(RL_E_INVALID, "rl_get_state_host: null argument");
    DeviceGuard guard(e->device);
    const size_t bytes = 40 * (size_t)e->n;
    if (e->n <= e->small_host) {
        Stage h, d;
        int rc = ensure_pinned(e, h, d);
        if (rc != RL_OK) return rc;
        rc = order_after_device_api(e, e->h_streams[0]);
        if (rc != RL_OK) return rc;
        // 40 n bytes fit in the staging span that starts at the observation area (obs 32 n + reward 4 n +
        // flags 2 n + terminal observations 32 n, contiguous); host-path calls on a handle never overlap
        rc = rl_get_state(e, reinterpret_cast<uint32_t*>(d.obs), e->h_streams[0]);
        if (rc != RL_OK) return rc;
        RL_CUDA(cudaStreamSynchronize(e->h_streams[0]));
        memcpy(state_soa, h.obs, bytes);
        return RL_OK;
    }
    Stage s;
    int rc = ensure_stage(e, s);
    if (rc != RL_OK) return rc;
    rc = order_after_device_api(e, e->h_streams[0]);
    if (rc != RL_OK) return rc;
    rc = rl_get_state(e, reinterpret_cast<uint32_t*>(s.obs), e->h_streams[0]);
    if (rc != RL_OK) return rc;
    RL_CUDA(cudaMemcpyAsync(state_soa, s.obs, bytes, cudaMemcpyDeviceToHost, e->h_streams[0]));
    RL_CUDA(cudaStreamSynchronize(e->h_streams[0]));
    return RL_OK;
}

}  // extern "C"
