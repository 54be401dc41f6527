// Generalised advantage estimation over a device-resident rollout buffer (SURVEY.md section 8f
// row 2, BASELINE.json configs[4]).  The reference trains with stable-baselines3 PPO
// (scripts/train.py:119-134, gamma / gae_lambda from config.yaml:100-104); its RolloutBuffer.
// compute_returns_and_advantage is the recurrence below, run on the host over [T, n_envs] numpy
// arrays.  Here one thread owns one rocket and walks its T steps backwards; for a fixed step the
// accesses of a warp are contiguous ([T][N] layout), so the kernel streams 9 B in / 8 B out per
// (step, rocket).  Only the arithmetic of the recurrence is serial: the loads of kAhead steps are
// issued together, ahead of the chain that consumes them, and CTAs own consecutive rockets and are
// dispatched in address order (profiles/README.md "grid order").
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include <cuda_runtime.h>

#include "../../include/rocket_landing_b200.h"
#include "rl_launch.cuh"

namespace {

thread_local char g_ro_err[256] = "";
int ro_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ro_err, sizeof(g_ro_err), fmt, ap);
    va_end(ap);
    return code;
}

// SB3 buffers.py RolloutBuffer.compute_returns_and_advantage:
//   for step in reversed(range(T)):
//       next_non_terminal = 1 - (dones if step == T-1 else episode_starts[step+1])
//       next_values       = last_values if step == T-1 else values[step+1]
//       delta = rewards[step] + gamma * next_values * next_non_terminal - values[step]
//       last_gae_lam = delta + gamma * gae_lambda * next_non_terminal * last_gae_lam
//       advantages[step] = last_gae_lam
//   returns = advantages + values
constexpr int kAhead = 4;

__global__ void __launch_bounds__(256) gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                  const uint8_t* __restrict__ episode_starts,
                                                  const float* __restrict__ last_values,
                                                  const uint8_t* __restrict__ last_dones, float* __restrict__ advantages,
                                                  float* __restrict__ returns, int T, int64_t n, float gamma, float lam) {
    const int64_t i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= n) return;
    float next_value = last_values[i];
    float next_non_terminal = last_dones[i] ? 0.0f : 1.0f;
    float gae = 0.0f;
    for (int t0 = T - 1; t0 >= 0; t0 -= kAhead) {
        float v[kAhead], r[kAhead];
        uint8_t es[kAhead];
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {                 // steps t0, t0 - 1, ...: every load first
            const int t = t0 - u;
            if (t >= 0) {
                const int64_t k = (int64_t)t * n + i;
                v[u] = values[k];
                r[u] = rewards[k];
                es[u] = episode_starts[k];
            }
        }
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
            const int t = t0 - u;
            if (t >= 0) {
                const int64_t k = (int64_t)t * n + i;
                const float delta = fmaf(gamma * next_value, next_non_terminal, r[u]) - v[u];
                gae = fmaf(gamma * lam * next_non_terminal, gae, delta);
                advantages[k] = gae;
                returns[k] = gae + v[u];
                next_value = v[u];
                next_non_terminal = es[u] ? 0.0f : 1.0f;
            }
        }
    }
}

// What SB3's collect_rollouts derives on the host from `dones` and `infos` after every step, in one pass
// over the two flag arrays: the next step's episode_start, the "ended by the time limit alone" mask whose
// rewards get the gamma V(terminal observation) bootstrap, and whether there is any such rocket at all.
__global__ void __launch_bounds__(256) episode_flags_kernel(const uint8_t* __restrict__ terminated,
                                                            const uint8_t* __restrict__ truncated,
                                                            uint8_t* __restrict__ episode_start_next,
                                                            uint8_t* __restrict__ time_limit, int* __restrict__ any_time_limit,
                                                            int64_t n, int vectorised) {
    rl::rl_grid_dependency_wait();                     // (programmatic dependent launch: rl_launch.cuh)
    const int64_t t = blockIdx.x * 256ll + threadIdx.x;
    uint32_t any = 0;
    if (vectorised) {                                      // 16 rockets per thread (all pointers 16-byte aligned)
        if (16 * t + 16 <= n) {
            const uint4 a = reinterpret_cast<const uint4*>(terminated)[t], b = reinterpret_cast<const uint4*>(truncated)[t];
            const uint4 tl = make_uint4(b.x & ~a.x, b.y & ~a.y, b.z & ~a.z, b.w & ~a.w);   // flags are 0 / 1 bytes
            if (episode_start_next) reinterpret_cast<uint4*>(episode_start_next)[t] = make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w);
            reinterpret_cast<uint4*>(time_limit)[t] = tl;
            any = tl.x | tl.y | tl.z | tl.w;
        } else {
            for (int64_t i = 16 * t; i < n; ++i) {
                const uint8_t a = terminated[i], b = truncated[i], tl = (uint8_t)(b & (a ^ 1u));
                if (episode_start_next) episode_start_next[i] = a | b;
                time_limit[i] = tl;
                any |= tl;
            }
        }
    } else if (t < n) {
        const uint8_t a = terminated[t], b = truncated[t], tl = (uint8_t)(b & (a ^ 1u));
        if (episode_start_next) episode_start_next[t] = a | b;
        time_limit[t] = tl;
        any = tl;
    }
    // (every writer stores the same 1: a plain store, which also works when the word lives in pinned host memory)
    if (__any_sync(0xffffffffu, any != 0) && (threadIdx.x & 31) == 0) *reinterpret_cast<volatile int*>(any_time_limit) = 1;
}

}  // namespace

extern "C" {

const char* rl_rollout_last_error(void) { return g_ro_err; }

// All pointers are DEVICE pointers; rewards / values / advantages / returns are float32 [T][N],
// episode_starts uint8 [T][N] (1 where the step's observation is the first of an episode),
// last_values float32 [N] and last_dones uint8 [N] describe the state after the last step.
int rl_gae(const float* rewards, const float* values, const uint8_t* episode_starts, const float* last_values,
           const uint8_t* last_dones, float* advantages, float* returns, int T, int64_t n_envs, double gamma,
           double gae_lambda, void* stream) {
    if (!rewards || !values || !episode_starts || !last_values || !last_dones || !advantages || !returns)
        return ro_fail(RL_E_INVALID, "rl_gae: null argument");
    if (T < 1 || n_envs < 1) return ro_fail(RL_E_INVALID, "rl_gae: T and n_envs must be positive");
    if (n_envs > 2147483392LL) return ro_fail(RL_E_INVALID, "rl_gae: at most 2147483392 rockets per call");
    const int grid = (int)((n_envs + 255) / 256);             // one CTA per 256 consecutive rockets
    gae_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rewards, values, episode_starts, last_values, last_dones,
                                                       advantages, returns, T, n_envs, (float)gamma, (float)gae_lambda);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ro_fail(RL_E_CUDA, "rl_gae launch failed: %s", cudaGetErrorString(e));
    return RL_OK;
}

// terminated, truncated: uint8 [N] (0 / 1) of the step just taken.  Writes episode_start_next[i] = terminated |
// truncated (nullable), time_limit[i] = truncated & !terminated, and stores 1 into *any_time_limit (int32: device
// memory, or pinned host memory the caller reads after an event) if any rocket ended by the time limit alone.
// DEVICE pointers, enqueue only.
int rl_episode_flags(const uint8_t* terminated, const uint8_t* truncated, uint8_t* episode_start_next, uint8_t* time_limit,
                     int32_t* any_time_limit, int64_t n_envs, void* stream) {
    if (!terminated || !truncated || !time_limit || !any_time_limit) return ro_fail(RL_E_INVALID, "rl_episode_flags: null argument");
    if (n_envs < 1) return ro_fail(RL_E_INVALID, "rl_episode_flags: n_envs must be positive");
    const uintptr_t bits = reinterpret_cast<uintptr_t>(terminated) | reinterpret_cast<uintptr_t>(truncated) |
                           reinterpret_cast<uintptr_t>(episode_start_next) | reinterpret_cast<uintptr_t>(time_limit);
    const int vectorised = (bits & 15) == 0;
    const int64_t threads = vectorised ? (n_envs + 15) / 16 : n_envs;
    if ((threads + 255) / 256 > 2147483647LL) return ro_fail(RL_E_INVALID, "rl_episode_flags: too many rockets per call");
    cudaError_t e = rl::rl_launch(episode_flags_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
                                  terminated, truncated, episode_start_next, time_limit, any_time_limit, n_envs, vectorised);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return ro_fail(RL_E_CUDA, "rl_episode_flags launch failed: %s", cudaGetErrorString(e));
    return RL_OK;
}

}  // extern "C"
