// Generalised advantage estimation over a device-resident rollout buffer (SURVEY.md section 8f
// row 2, BASELINE.json configs[4]).  The reference trains with stable-baselines3 PPO
// (scripts/train.py:119-134, gamma / gae_lambda from config.yaml:100-104); its RolloutBuffer.
// compute_returns_and_advantage is the recurrence below, run on the host over [T, n_envs] numpy
// arrays.  Here one thread owns one rocket and walks its T steps backwards; for a fixed step the
// accesses of a warp are contiguous ([T][N] layout), so the kernel streams 9 B in / 8 B out per
// (step, rocket) at HBM speed.
#include <cstdarg>
#include <cstdio>

#include <cuda_runtime.h>

#include "../../include/rocket_landing_b200.h"

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
__global__ void __launch_bounds__(256) gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                  const uint8_t* __restrict__ episode_starts,
                                                  const float* __restrict__ last_values,
                                                  const uint8_t* __restrict__ last_dones, float* __restrict__ advantages,
                                                  float* __restrict__ returns, int T, int64_t n, float gamma, float lam) {
    for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        float next_value = last_values[i];
        float next_non_terminal = last_dones[i] ? 0.0f : 1.0f;
        float gae = 0.0f;
        for (int t = T - 1; t >= 0; --t) {
            const int64_t k = (int64_t)t * n + i;
            const float v = values[k];
            const float delta = fmaf(gamma * next_value, next_non_terminal, rewards[k]) - v;
            gae = fmaf(gamma * lam * next_non_terminal, gae, delta);
            advantages[k] = gae;
            returns[k] = gae + v;
            next_value = v;
            next_non_terminal = episode_starts[k] ? 0.0f : 1.0f;
        }
    }
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
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ctas = (n_envs + 255) / 256;
    const int grid = (int)(ctas < (int64_t)sms * 8 ? ctas : (int64_t)sms * 8);
    gae_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rewards, values, episode_starts, last_values, last_dones,
                                                       advantages, returns, T, n_envs, (float)gamma, (float)gae_lambda);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return ro_fail(RL_E_CUDA, "rl_gae launch failed: %s", cudaGetErrorString(e));
    return RL_OK;
}

}  // extern "C"
