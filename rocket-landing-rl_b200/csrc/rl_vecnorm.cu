// On-device VecNormalize: the wrapper the reference puts around its vectorised environment
// (scripts/train.py:86-88: VecNormalize(norm_obs=True, norm_reward=True, clip_obs=10, gamma)).
// Semantics follow stable-baselines3 2.3.2 (common/running_mean_std.py, common/vec_env/
// vec_normalize.py), which is not part of the reference tree; oracle/vecnormalize_oracle.py
// restates it.  One step of the wrapper is three stream-ordered launches:
//
//   moments   : ret <- ret gamma + reward (stored); per-CTA fp64 sums of the shifted observation
//               words, their squares, the returns and their squares -> 18 global fp64 atomics / CTA
//   finalize  : one warp folds the batch moments into the running mean / var / count (Chan's
//               parallel update, fp64) and publishes fp32 mean and 1/sqrt(var + eps)
//   apply     : obs <- clip((obs - mean) inv_std), reward <- clip(reward inv_std_ret), terminal
//               observations of done rockets likewise, ret[done] <- 0
//
// The normalisation of a batch depends on statistics that already include that batch
// (vec_normalize.py: obs_rms.update(obs) precedes normalize_obs(obs)), hence the grid-wide
// reduction between the first and the last launch.  All three are HBM-streaming kernels
// (moments: 40 B read + 4 B written per rocket; apply: 36-40 B read + 36-40 B written).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <new>

#include <cuda_runtime.h>

#include "../../include/rocket_landing_b200.h"

namespace {

thread_local char g_vn_err[256] = "";
int vn_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_vn_err, sizeof(g_vn_err), fmt, ap);
    va_end(ap);
    return code;
}

constexpr int kThreads = 256;
constexpr int kObs = 8;
constexpr int kAcc = 2 * kObs + 2;     // sum, sum of squares per observation word; same for the return

// device-resident state of the wrapper
struct VnState {
    double obs_mean[kObs], obs_var[kObs], obs_count;     // RunningMeanStd(shape=(8,)), epsilon count 1e-4
    double ret_mean, ret_var, ret_count;                 // RunningMeanStd(shape=())
    double acc[kAcc];                                    // batch sums, zeroed by finalize
    float shift[kObs + 1];                               // summation shift = running means at the last finalize
    float mean_f[kObs], inv_std_f[kObs], ret_inv_std_f;  // what apply reads
};

__global__ void __launch_bounds__(kThreads) vn_moments(VnState* __restrict__ st, const float4* __restrict__ obs,
                                                       const float* __restrict__ reward, float* __restrict__ returns,
                                                       uint32_t n, float gamma, int with_obs, int with_ret) {
    __shared__ double red[kThreads / 32][kAcc];
    double acc[kAcc];
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc[k] = 0.0;
    float sh[kObs + 1];
#pragma unroll
    for (int k = 0; k <= kObs; ++k) sh[k] = st->shift[k];
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        if (with_obs) {
            const float4 a = obs[2 * (size_t)i], b = obs[2 * (size_t)i + 1];
            const float v[kObs] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < kObs; ++k) {
                const double d = (double)v[k] - (double)sh[k];
                acc[k] += d;
                acc[kObs + k] += d * d;
            }
        }
        if (with_ret) {
            const float r = fmaf(returns[i], gamma, reward[i]);     // vec_normalize.py: ret = ret * gamma + reward
            returns[i] = r;
            const double d = (double)r - (double)sh[kObs];
            acc[2 * kObs] += d;
            acc[2 * kObs + 1] += d * d;
        }
    }
#pragma unroll
    for (int k = 0; k < kAcc; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kAcc; ++k) red[warp][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < kAcc) {
        double s = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) s += red[w][threadIdx.x];
        if (s != 0.0) atomicAdd(&st->acc[threadIdx.x], s);
    }
}

// running_mean_std.py: update_from_moments
__device__ void fold(double& mean, double& var, double count, double batch_mean, double batch_var, double batch_count) {
    const double delta = batch_mean - mean;
    const double tot = count + batch_count;
    const double m2 = var * count + batch_var * batch_count + delta * delta * count * batch_count / tot;
    mean = mean + delta * batch_count / tot;
    var = m2 / tot;
}

__global__ void vn_finalize(VnState* st, double batch_count, double epsilon, int update_obs, int update_ret) {
    const int k = threadIdx.x;
    if (update_obs && k < kObs) {
        const double s = st->acc[k] / batch_count, s2 = st->acc[kObs + k] / batch_count;
        fold(st->obs_mean[k], st->obs_var[k], st->obs_count, (double)st->shift[k] + s, s2 - s * s, batch_count);
    }
    if (update_ret && k == kObs) {
        const double s = st->acc[2 * kObs] / batch_count, s2 = st->acc[2 * kObs + 1] / batch_count;
        fold(st->ret_mean, st->ret_var, st->ret_count, (double)st->shift[kObs] + s, s2 - s * s, batch_count);
    }
    __syncwarp();
    if (k == 0) {
        if (update_obs) st->obs_count += batch_count;
        if (update_ret) st->ret_count += batch_count;
    }
    if (k < kAcc) st->acc[k] = 0.0;
    if (k < kObs) {
        st->mean_f[k] = (float)st->obs_mean[k];
        st->inv_std_f[k] = (float)(1.0 / sqrt(st->obs_var[k] + epsilon));
        st->shift[k] = (float)st->obs_mean[k];
    }
    if (k == kObs) {
        st->ret_inv_std_f = (float)(1.0 / sqrt(st->ret_var + epsilon));
        st->shift[kObs] = (float)st->ret_mean;
    }
}

__device__ __forceinline__ float4 norm4(float4 v, const float* m, const float* s, float c) {
    return make_float4(fminf(fmaxf((v.x - m[0]) * s[0], -c), c), fminf(fmaxf((v.y - m[1]) * s[1], -c), c),
                       fminf(fmaxf((v.z - m[2]) * s[2], -c), c), fminf(fmaxf((v.w - m[3]) * s[3], -c), c));
}

__global__ void __launch_bounds__(kThreads) vn_apply(const VnState* __restrict__ st, const float4* __restrict__ obs_in,
                                                     float4* __restrict__ obs_out, const float* __restrict__ rew_in,
                                                     float* __restrict__ rew_out, const uint8_t* __restrict__ terminated,
                                                     const uint8_t* __restrict__ truncated, float4* __restrict__ terminal_obs,
                                                     float* __restrict__ returns, uint32_t n, float clip_obs,
                                                     float clip_reward, int norm_obs, int norm_reward) {
    float m[kObs], s[kObs];
#pragma unroll
    for (int k = 0; k < kObs; ++k) {
        m[k] = st->mean_f[k];
        s[k] = st->inv_std_f[k];
    }
    const float rs = st->ret_inv_std_f;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        if (obs_in) {
            float4 a = obs_in[2 * (size_t)i], b = obs_in[2 * (size_t)i + 1];
            if (norm_obs) {
                a = norm4(a, m, s, clip_obs);
                b = norm4(b, m + 4, s + 4, clip_obs);
            }
            obs_out[2 * (size_t)i] = a;
            obs_out[2 * (size_t)i + 1] = b;
        }
        if (rew_in) {
            const float r = rew_in[i];
            rew_out[i] = norm_reward ? fminf(fmaxf(r * rs, -clip_reward), clip_reward) : r;
        }
        if (terminated) {
            const bool done = (terminated[i] | truncated[i]) != 0;
            if (done) {
                returns[i] = 0.0f;                                   // vec_normalize.py: self.returns[dones] = 0
                if (terminal_obs && norm_obs) {
                    terminal_obs[2 * (size_t)i] = norm4(terminal_obs[2 * (size_t)i], m, s, clip_obs);
                    terminal_obs[2 * (size_t)i + 1] = norm4(terminal_obs[2 * (size_t)i + 1], m + 4, s + 4, clip_obs);
                }
            }
        }
    }
}

__global__ void vn_init(VnState* st, double count0) {
    const int k = threadIdx.x;
    if (k < kObs) {
        st->obs_mean[k] = 0.0;
        st->obs_var[k] = 1.0;
        st->mean_f[k] = 0.0f;
        st->inv_std_f[k] = 1.0f;
        st->shift[k] = 0.0f;
    }
    if (k < kAcc) st->acc[k] = 0.0;
    if (k == 0) {
        st->obs_count = count0;
        st->ret_mean = 0.0;
        st->ret_var = 1.0;
        st->ret_count = count0;
        st->ret_inv_std_f = 1.0f;
        st->shift[kObs] = 0.0f;
    }
}

}  // namespace

struct RlVecNorm {
    VnState* st = nullptr;
    float* returns = nullptr;
    int64_t n = 0;
    double gamma = 0.99, epsilon = 1e-8, clip_obs = 10.0, clip_reward = 10.0;
    int device = 0, grid = 1;
    int64_t launches = 0;
};

#define VN_CUDA(expr)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess) return vn_fail(RL_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));   \
    } while (0)

namespace {
struct Guard {
    int prev = -1;
    explicit Guard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) cudaSetDevice(dev); }
    ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

extern "C" {

const char* rl_vecnorm_last_error(void) { return g_vn_err; }

int rl_vecnorm_create(int64_t n_envs, double gamma, double epsilon, double clip_obs, double clip_reward, int device,
                      RlVecNorm** out) {
    if (!out || n_envs <= 0 || n_envs > 2147483392LL) return vn_fail(RL_E_INVALID, "rl_vecnorm_create: bad n_envs");
    if (!(epsilon > 0) || !(clip_obs > 0) || !(clip_reward > 0)) return vn_fail(RL_E_INVALID, "rl_vecnorm_create: epsilon and clips must be positive");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        return vn_fail(RL_E_CUDA, "rl_vecnorm_create: CUDA device %d not available", device);
    RlVecNorm* v = new (std::nothrow) RlVecNorm();
    if (!v) return vn_fail(RL_E_NOMEM, "rl_vecnorm_create: out of host memory");
    v->n = n_envs; v->gamma = gamma; v->epsilon = epsilon; v->clip_obs = clip_obs; v->clip_reward = clip_reward;
    v->device = device;
    Guard g(device);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int64_t ctas = (n_envs + kThreads - 1) / kThreads;
    v->grid = (int)(ctas < (int64_t)sms * 8 ? ctas : (int64_t)sms * 8);
    if (cudaMalloc(&v->st, sizeof(VnState)) != cudaSuccess || cudaMalloc(&v->returns, 4 * (size_t)n_envs) != cudaSuccess ||
        cudaMemset(v->returns, 0, 4 * (size_t)n_envs) != cudaSuccess) {
        rl_vecnorm_destroy(v);
        return vn_fail(RL_E_NOMEM, "rl_vecnorm_create: device allocation failed");
    }
    vn_init<<<1, 32>>>(v->st, 1e-4);            // RunningMeanStd(epsilon=1e-4)
    VN_CUDA(cudaGetLastError());
    VN_CUDA(cudaDeviceSynchronize());
    *out = v;
    return RL_OK;
}

int rl_vecnorm_destroy(RlVecNorm* v) {
    if (!v) return RL_OK;
    Guard g(v->device);
    if (v->st) cudaFree(v->st);
    if (v->returns) cudaFree(v->returns);
    delete v;
    return RL_OK;
}

// VecNormalize.reset: returns <- 0, fold the reset observations into the statistics (if training
// and norm_obs), write the normalised observations.  obs_out may alias obs_in.
int rl_vecnorm_reset(RlVecNorm* v, const float* obs_in, float* obs_out, int training, int norm_obs, void* stream) {
    if (!v || !obs_in || !obs_out) return vn_fail(RL_E_INVALID, "rl_vecnorm_reset: null argument");
    Guard g(v->device);
    cudaStream_t st = (cudaStream_t)stream;
    VN_CUDA(cudaMemsetAsync(v->returns, 0, 4 * (size_t)v->n, st));
    if (training && norm_obs) {
        vn_moments<<<v->grid, kThreads, 0, st>>>(v->st, (const float4*)obs_in, nullptr, v->returns, (uint32_t)v->n, 0.0f, 1, 0);
        vn_finalize<<<1, 32, 0, st>>>(v->st, (double)v->n, v->epsilon, 1, 0);
        v->launches += 2;
    }
    vn_apply<<<v->grid, kThreads, 0, st>>>(v->st, (const float4*)obs_in, (float4*)obs_out, nullptr, nullptr, nullptr,
                                          nullptr, nullptr, v->returns, (uint32_t)v->n, (float)v->clip_obs,
                                          (float)v->clip_reward, norm_obs, 0);
    v->launches += 1;
    VN_CUDA(cudaGetLastError());
    return RL_OK;
}

// VecNormalize.step_wait on the outputs of rl_step.  obs_out / reward_out may alias the inputs;
// terminal_obs (nullable) is normalised in place for done rockets.
int rl_vecnorm_step(RlVecNorm* v, const float* obs_in, const float* reward_in, const uint8_t* terminated,
                    const uint8_t* truncated, float* terminal_obs, float* obs_out, float* reward_out, int training,
                    int norm_obs, int norm_reward, void* stream) {
    if (!v || !obs_in || !reward_in || !terminated || !truncated || !obs_out || !reward_out)
        return vn_fail(RL_E_INVALID, "rl_vecnorm_step: null argument");
    Guard g(v->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (training) {
        vn_moments<<<v->grid, kThreads, 0, st>>>(v->st, (const float4*)obs_in, reward_in, v->returns, (uint32_t)v->n,
                                                (float)v->gamma, norm_obs, 1);
        vn_finalize<<<1, 32, 0, st>>>(v->st, (double)v->n, v->epsilon, norm_obs, 1);
        v->launches += 2;
    }
    vn_apply<<<v->grid, kThreads, 0, st>>>(v->st, (const float4*)obs_in, (float4*)obs_out, reward_in, reward_out,
                                          terminated, truncated, (float4*)terminal_obs, v->returns, (uint32_t)v->n,
                                          (float)v->clip_obs, (float)v->clip_reward, norm_obs, norm_reward);
    v->launches += 1;
    VN_CUDA(cudaGetLastError());
    return RL_OK;
}

// normalize_obs of an arbitrary batch with the current statistics (evaluation, agent.py:181)
int rl_vecnorm_normalize_obs(RlVecNorm* v, const float* obs_in, float* obs_out, int64_t rows, void* stream) {
    if (!v || !obs_in || !obs_out || rows <= 0) return vn_fail(RL_E_INVALID, "rl_vecnorm_normalize_obs: bad argument");
    Guard g(v->device);
    const int64_t ctas = (rows + kThreads - 1) / kThreads;
    vn_apply<<<(int)(ctas < v->grid ? ctas : v->grid), kThreads, 0, (cudaStream_t)stream>>>(
        v->st, (const float4*)obs_in, (float4*)obs_out, nullptr, nullptr, nullptr, nullptr, nullptr, v->returns,
        (uint32_t)rows, (float)v->clip_obs, (float)v->clip_reward, 1, 0);
    v->launches += 1;
    VN_CUDA(cudaGetLastError());
    return RL_OK;
}

// statistics as 20 doubles on the HOST: obs mean[8], obs var[8], obs count, ret mean, ret var, ret count
int rl_vecnorm_get_stats(RlVecNorm* v, double* out20) {
    if (!v || !out20) return vn_fail(RL_E_INVALID, "rl_vecnorm_get_stats: null argument");
    Guard g(v->device);
    VN_CUDA(cudaDeviceSynchronize());
    VN_CUDA(cudaMemcpy(out20, v->st, 20 * sizeof(double), cudaMemcpyDeviceToHost));
    return RL_OK;
}

int rl_vecnorm_set_stats(RlVecNorm* v, const double* in20) {
    if (!v || !in20) return vn_fail(RL_E_INVALID, "rl_vecnorm_set_stats: null argument");
    Guard g(v->device);
    VN_CUDA(cudaDeviceSynchronize());
    VN_CUDA(cudaMemcpy(v->st, in20, 20 * sizeof(double), cudaMemcpyHostToDevice));
    vn_finalize<<<1, 32>>>(v->st, 1.0, v->epsilon, 0, 0);      // republish fp32 mean / inv_std / shift
    VN_CUDA(cudaGetLastError());
    VN_CUDA(cudaDeviceSynchronize());
    return RL_OK;
}

int64_t rl_vecnorm_launch_count(const RlVecNorm* v) { return v ? v->launches : 0; }

}  // extern "C"
