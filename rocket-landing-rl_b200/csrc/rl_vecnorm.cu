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
// (moments: 40 B read + 4 B written per rocket; apply: 36-40 B read + 36-40 B written), shaped by
// what the step kernel taught (profiles/README.md "grid order"): CTA b owns the kChunk (kMomChunk) consecutive
// rockets from b * kChunk and the hardware dispatches CTAs in address order (no persistent
// grid-stride loop), and the observation matrix is walked as a flat float4 array -- element e is
// words 4 (e & 1) .. 4 (e & 1) + 3 of rocket e / 2 -- so that every load and store instruction of a
// warp covers 512 contiguous bytes instead of half of each of 32 sectors.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <new>

#include <cuda_runtime.h>

#include "../../include/rocket_landing_b200.h"
#include "rl_launch.cuh"

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
constexpr int kChunk = 4 * kThreads;   // apply: rockets per CTA -- 8 float4 of observations + 4 rewards per thread, all loaded up front
// moments: rockets per CTA (amortises the CTA's reduction + 18 atomics).  16 tiles for large batches; 8 up to 2 Mi
// rockets, where 16 leave fewer than two CTAs per SM (1 Mi rockets: env step + VecNormalize 63.3 -> 59.4 us)
constexpr int kMomChunk = 16 * kThreads, kMomChunkSmall = 8 * kThreads;
constexpr int64_t kMomSmallBatch = 2 * 1024 * 1024;
constexpr int kAccSlots = 32;          // copies of the batch sums (CTA b adds into copy b mod 32; finalize folds them)
constexpr int kBatch = 2 * kChunk / kThreads;   // float4 observation elements a thread loads before it consumes any
constexpr int kRows = kChunk / kThreads;        // rockets per thread per batch

// device-resident state of the wrapper
struct VnState {
    double obs_mean[kObs], obs_var[kObs], obs_count;     // RunningMeanStd(shape=(8,)), epsilon count 1e-4
    double ret_mean, ret_var, ret_count;                 // RunningMeanStd(shape=())
    double acc[kAccSlots][kAcc];                         // batch sums, zeroed by finalize
    float shift[kObs + 1];                               // summation shift = running means at the last finalize
    float mean_f[kObs], inv_std_f[kObs], ret_inv_std_f;  // what apply reads
};

__global__ void __launch_bounds__(kThreads) vn_moments(VnState* __restrict__ st, const float4* __restrict__ obs,
                                                       const float* __restrict__ reward, float* __restrict__ returns,
                                                       uint32_t n, float gamma, int with_obs, int with_ret, uint32_t chunk) {
    rl::rl_grid_dependency_wait();                     // (programmatic dependent launch: rl_launch.cuh)
    __shared__ double red[kThreads / 32][kAcc];
    const uint32_t tid = threadIdx.x;
    const uint32_t r0 = blockIdx.x * chunk;                                  // this CTA's rockets [r0, r1)
    const uint32_t r1 = min(n, r0 + chunk);
    // a thread only ever sees float4 elements of ONE parity (the stride 256 is even): the low or the
    // high four words of an observation row -> 4 sums + 4 sums of squares per thread
    const int half = (int)(tid & 1u);
    double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0}, rs1 = 0.0, rs2 = 0.0;
    if (with_obs) {
        const float sh0 = st->shift[4 * half], sh1 = st->shift[4 * half + 1], sh2 = st->shift[4 * half + 2],
                    sh3 = st->shift[4 * half + 3];
        const size_t e1 = 2 * (size_t)r1;
        for (size_t base = 2 * (size_t)r0 + tid; base < e1; base += (size_t)kThreads * kBatch) {
            float4 v[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {              // kBatch loads in flight per thread
                const size_t e = base + (size_t)u * kThreads;
                v[u] = e < e1 ? obs[e] : make_float4(sh0, sh1, sh2, sh3);            // (adds exactly 0)
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const double d0 = (double)v[u].x - (double)sh0, d1 = (double)v[u].y - (double)sh1;
                const double d2 = (double)v[u].z - (double)sh2, d3 = (double)v[u].w - (double)sh3;
                s1[0] += d0; s2[0] += d0 * d0;
                s1[1] += d1; s2[1] += d1 * d1;
                s1[2] += d2; s2[2] += d2 * d2;
                s1[3] += d3; s2[3] += d3 * d3;
            }
        }
    }
    if (with_ret) {
        const float shr = st->shift[kObs];
        for (uint32_t base = r0 + tid; base < r1; base += (uint32_t)kThreads * kBatch) {
            float ret[kBatch], rew[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const uint32_t i = base + (uint32_t)u * kThreads;
                ret[u] = i < r1 ? returns[i] : 0.0f;
                rew[u] = i < r1 ? reward[i] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const uint32_t i = base + (uint32_t)u * kThreads;
                if (i < r1) {
                    const float r = fmaf(ret[u], gamma, rew[u]);    // vec_normalize.py: ret = ret * gamma + reward
                    returns[i] = r;
                    const double d = (double)r - (double)shr;
                    rs1 += d;
                    rs2 += d * d;
                }
            }
        }
    }
    // warp reduction that keeps the two parities apart (xor 2 .. 16), then lanes 0 / 1 hold them
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 1; o >>= 1) {
            s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
            s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, o);
        rs2 += __shfl_xor_sync(0xffffffffu, rs2, o);
    }
    const int lane = (int)(tid & 31u), warp = (int)(tid >> 5);
    if (lane < 2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            red[warp][4 * lane + k] = s1[k];
            red[warp][kObs + 4 * lane + k] = s2[k];
        }
    }
    if (lane == 0) {
        red[warp][2 * kObs] = rs1;
        red[warp][2 * kObs + 1] = rs2;
    }
    __syncthreads();
    if (tid < kAcc) {
        double s = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) s += red[w][tid];
        if (s != 0.0) atomicAdd(&st->acc[blockIdx.x % kAccSlots][tid], s);
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
    rl::rl_grid_dependency_wait();
    const int k = threadIdx.x;
    __shared__ double acc[kAcc];
    {   // fold (and clear) the kAccSlots copies of the batch sums: lane = slot, one shuffle tree per word;
        // all 18 loads are issued before the first store so that they overlap (same array: the
        // compiler must otherwise keep every load behind the previous word's store)
        static_assert(kAccSlots == 32, "one slot per lane");
        double v[kAcc];
#pragma unroll
        for (int w = 0; w < kAcc; ++w) v[w] = st->acc[k][w];
#pragma unroll
        for (int w = 0; w < kAcc; ++w) st->acc[k][w] = 0.0;
#pragma unroll
        for (int w = 0; w < kAcc; ++w) {
            double s = v[w];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (k == 0) acc[w] = s;
        }
    }
    __syncwarp();
    if (update_obs && k < kObs) {
        const double s = acc[k] / batch_count, s2 = acc[kObs + k] / batch_count;
        fold(st->obs_mean[k], st->obs_var[k], st->obs_count, (double)st->shift[k] + s, s2 - s * s, batch_count);
    }
    if (update_ret && k == kObs) {
        const double s = acc[2 * kObs] / batch_count, s2 = acc[2 * kObs + 1] / batch_count;
        fold(st->ret_mean, st->ret_var, st->ret_count, (double)st->shift[kObs] + s, s2 - s * s, batch_count);
    }
    __syncwarp();
    if (k == 0) {
        if (update_obs) st->obs_count += batch_count;
        if (update_ret) st->ret_count += batch_count;
    }
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
    rl::rl_grid_dependency_wait();                     // (programmatic dependent launch: rl_launch.cuh)
    const uint32_t tid = threadIdx.x;
    const uint32_t r0 = blockIdx.x * (uint32_t)kChunk, r1 = min(n, r0 + (uint32_t)kChunk);
    if (obs_in) {
        // flat float4 walk: this thread only meets one half of the observation row
        const int half = (int)(tid & 1u);
        float m[4], s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            m[k] = st->mean_f[4 * half + k];
            s[k] = st->inv_std_f[4 * half + k];
        }
        const size_t e1 = 2 * (size_t)r1;
        for (size_t base = 2 * (size_t)r0 + tid; base < e1; base += (size_t)kThreads * kBatch) {
            float4 v[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const size_t e = base + (size_t)u * kThreads;
                if (e < e1) v[u] = obs_in[e];
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const size_t e = base + (size_t)u * kThreads;
                if (e < e1) obs_out[e] = norm_obs ? norm4(v[u], m, s, clip_obs) : v[u];
            }
        }
    }
    const float rs = st->ret_inv_std_f;
    float rew[kRows];
    uint32_t done_bits = 0u;
#pragma unroll
    for (int u = 0; u < kRows; ++u) {                     // every load of the thread's rockets first
        const uint32_t i = r0 + tid + (uint32_t)u * kThreads;
        if (i < r1) {
            if (rew_in) rew[u] = rew_in[i];
            if (terminated) done_bits |= ((terminated[i] | truncated[i]) != 0 ? 1u : 0u) << u;
        }
    }
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
        const uint32_t i = r0 + tid + (uint32_t)u * kThreads;
        if (i >= r1) continue;
        if (rew_in) rew_out[i] = norm_reward ? fminf(fmaxf(rew[u] * rs, -clip_reward), clip_reward) : rew[u];
        if ((done_bits >> u) & 1u) {
            returns[i] = 0.0f;                                       // vec_normalize.py: self.returns[dones] = 0
            if (terminal_obs && norm_obs) {                          // rare (1 rocket in ~113): whole row by one thread
                float m[kObs], s[kObs];
#pragma unroll
                for (int k = 0; k < kObs; ++k) {
                    m[k] = st->mean_f[k];
                    s[k] = st->inv_std_f[k];
                }
                terminal_obs[2 * (size_t)i] = norm4(terminal_obs[2 * (size_t)i], m, s, clip_obs);
                terminal_obs[2 * (size_t)i + 1] = norm4(terminal_obs[2 * (size_t)i + 1], m + 4, s + 4, clip_obs);
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
    if (k < kAcc)
        for (int sl = 0; sl < kAccSlots; ++sl) st->acc[sl][k] = 0.0;
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
    int device = 0, grid = 1, mom_grid = 1, mom_chunk = kMomChunk;
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
    v->grid = (int)((n_envs + kChunk - 1) / kChunk);          // apply: one CTA per kChunk consecutive rockets
    v->mom_chunk = n_envs <= kMomSmallBatch ? kMomChunkSmall : kMomChunk;
    v->mom_grid = (int)((n_envs + v->mom_chunk - 1) / v->mom_chunk);
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
        vn_moments<<<v->mom_grid, kThreads, 0, st>>>(v->st, (const float4*)obs_in, nullptr, v->returns, (uint32_t)v->n, 0.0f, 1, 0, (uint32_t)v->mom_chunk);
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
        rl::rl_launch(vn_moments, dim3(v->mom_grid), dim3(kThreads), 0, st, v->st, (const float4*)obs_in, reward_in, v->returns,
                      (uint32_t)v->n, (float)v->gamma, norm_obs, 1, (uint32_t)v->mom_chunk);
        rl::rl_launch(vn_finalize, dim3(1), dim3(32), 0, st, v->st, (double)v->n, v->epsilon, norm_obs, 1);
        v->launches += 2;
    }
    rl::rl_launch(vn_apply, dim3(v->grid), dim3(kThreads), 0, st, (const VnState*)v->st, (const float4*)obs_in, (float4*)obs_out,
                  reward_in, reward_out, terminated, truncated, (float4*)terminal_obs, v->returns, (uint32_t)v->n,
                  (float)v->clip_obs, (float)v->clip_reward, norm_obs, norm_reward);
    v->launches += 1;
    VN_CUDA(cudaGetLastError());
    return RL_OK;
}

// normalize_obs of an arbitrary batch with the current statistics (evaluation, agent.py:181)
int rl_vecnorm_normalize_obs(RlVecNorm* v, const float* obs_in, float* obs_out, int64_t rows, void* stream) {
    if (!v || !obs_in || !obs_out || rows <= 0) return vn_fail(RL_E_INVALID, "rl_vecnorm_normalize_obs: bad argument");
    Guard g(v->device);
    const int64_t ctas = (rows + kChunk - 1) / kChunk;
    vn_apply<<<(int)ctas, kThreads, 0, (cudaStream_t)stream>>>(
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
