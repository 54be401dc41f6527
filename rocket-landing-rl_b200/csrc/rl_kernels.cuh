// Kernels of the batched environment step.  State layout in HBM (DESIGN.md "Data layout"):
//
//   plane A  float4[Np]  { x, y, angle, meta }          read + written every step
//   plane B  float4[Np]  { sx, sy, sangle, ep_return }   read + written every step
//   fuel     float [Np]                                  read + written every step
//   dry      float [Np]                                  read every step, written on reset only
//
// Np = n_envs rounded up to a 256-rocket tile; all planes live in one buffer so a checkpoint is
// one tensor.  A warp's 32 rockets are 512 contiguous bytes in planes A/B (LDG.128/STG.128),
// 128 in fuel/dry, 256 in the action array, 1 KiB in obs.
#pragma once
#include "rl_device.cuh"

namespace rl {

constexpr int kTile = 256;           // rockets per CTA iteration == threads per CTA

struct StatePlanes {
    float4* A;
    float4* B;
    float* fuel;
    float* dry;
};

__host__ __device__ inline int64_t padded_envs(int64_t n) { return (n + kTile - 1) / kTile * kTile; }

__host__ inline StatePlanes carve_state(void* base, int64_t n_envs) {
    const int64_t np = padded_envs(n_envs);
    char* p = static_cast<char*>(base);
    StatePlanes s;
    s.A = reinterpret_cast<float4*>(p);
    s.B = reinterpret_cast<float4*>(p + 16 * np);
    s.fuel = reinterpret_cast<float*>(p + 32 * np);
    s.dry = reinterpret_cast<float*>(p + 36 * np);
    return s;
}

struct StepArgs {
    StatePlanes st;
    const float2* __restrict__ actions;
    float4* __restrict__ obs;             // [N][2] float4
    float* __restrict__ reward;
    uint8_t* __restrict__ terminated;
    uint8_t* __restrict__ truncated;
    float4* __restrict__ terminal_obs;    // nullable
    float* __restrict__ episode_return;   // nullable
    int32_t* __restrict__ episode_length; // nullable
    int8_t* __restrict__ landing;         // nullable
    double* __restrict__ stats;           // [RL_STATS_WORDS]
    int64_t n;
    int64_t gid0;                         // global id of rocket 0 (Philox counter)
    uint64_t seed;
    unsigned long long* __restrict__ epoch;   // device: launch epoch (Philox counter), bumped by the last CTA
    unsigned int* __restrict__ ticket;        // device: CTAs finished in this launch
    int substeps;
};

// The launch epoch lives in device memory so that a CUDA-graph replay of the step advances the
// Philox stream like an ordinary launch does.  Every CTA reads it on entry; the last CTA to
// finish (ticket) bumps it for the next launch, which is stream-ordered after this one.
__device__ __forceinline__ void epoch_bump(unsigned long long* epoch, unsigned int* ticket) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *ticket = 0u;
            *epoch += 1ull;
        }
    }
}

// ---- block-level statistics: rare events go to shared-memory atomics, one global atomic per
// ---- slot per CTA at the end (grid is persistent, so that is a few thousand atomics per launch)
struct BlockStats {
    double sum[RL_STATS_WORDS];
};

__device__ __forceinline__ void stats_init(BlockStats& s) {
    if (threadIdx.x < RL_STATS_WORDS) s.sum[threadIdx.x] = 0.0;
    __syncthreads();
}

__device__ __forceinline__ void stats_flush(BlockStats& s, double* __restrict__ g, float reward_sum, int steps) {
    // per-thread running sums -> warp -> shared
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        reward_sum += __shfl_xor_sync(0xffffffffu, reward_sum, o);
        steps += __shfl_xor_sync(0xffffffffu, steps, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s.sum[RL_S_REWARD_SUM], (double)reward_sum);
        atomicAdd(&s.sum[RL_S_ENV_STEPS], (double)steps);
    }
    __syncthreads();
    if (threadIdx.x < RL_STATS_WORDS && s.sum[threadIdx.x] != 0.0) atomicAdd(&g[threadIdx.x], s.sum[threadIdx.x]);
}

__device__ __forceinline__ void stats_episode_end(BlockStats& s, const StepResult& o, float ep_ret, int ep_len) {
    atomicAdd(&s.sum[RL_S_EPISODES], 1.0);
    atomicAdd(&s.sum[RL_S_RETURN_SUM], (double)ep_ret);
    atomicAdd(&s.sum[RL_S_LENGTH_SUM], (double)ep_len);
    if (o.landing) atomicAdd(&s.sum[RL_S_SAFE + o.landing - 1], 1.0);
    else if (o.out_of_bounds) atomicAdd(&s.sum[RL_S_OUT_OF_BOUNDS], 1.0);
    else atomicAdd(&s.sum[RL_S_TIME_LIMIT], 1.0);
}

__device__ __forceinline__ void load_rocket(const StatePlanes& st, int64_t i, Rocket& r) {
    const float4 a = st.A[i];
    const float4 b = st.B[i];
    r.fuel = st.fuel[i];
    r.dry = st.dry[i];
    r.x = a.x; r.y = a.y; r.ang = a.z; r.meta = __float_as_uint(a.w);
    r.sx = b.x; r.sy = b.y; r.sang = b.z; r.ep_ret = b.w;
}

__device__ __forceinline__ void store_rocket(const StatePlanes& st, int64_t i, const Rocket& r, bool write_dry) {
    st.A[i] = make_float4(r.x, r.y, r.ang, __uint_as_float(r.meta));
    st.B[i] = make_float4(r.sx, r.sy, r.sang, r.ep_ret);
    st.fuel[i] = r.fuel;
    if (write_dry) st.dry[i] = r.dry;
}

// The environment step (lander.py:188-255 for every rocket) + VecEnv auto-reset.
// Persistent grid: each CTA walks tiles of 256 rockets with stride gridDim.x.
template <bool kAutoReset>
__global__ void __launch_bounds__(kTile) step_kernel(const DevParams P, const StepArgs a) {
    __shared__ BlockStats bs;
    stats_init(bs);
    const uint64_t epoch = *a.epoch;
    float reward_sum = 0.0f;
    int steps_done = 0;

    for (int64_t i = (int64_t)blockIdx.x * kTile + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * kTile) {
        Rocket r;
        load_rocket(a.st, i, r);
        const float2 act = __ldg(&a.actions[i]);

        StepResult o;
        float reward = 0.0f;
        bool done = false;
        for (int k = 0; k < a.substeps; ++k) {
            rocket_step(P, r, act.x, act.y, o);
            reward += o.reward;
            ++steps_done;
            if (o.tipped) atomicAdd(&bs.sum[RL_S_TIP_STEPS], 1.0);
            if (o.nonfinite) atomicAdd(&bs.sum[RL_S_NONFINITE], 1.0);
            done = o.terminated || o.truncated;
            if (done) break;                     // stop integrating at the first done
        }
        reward_sum += reward;

        bool write_dry = false;
        if (done) {
            const int ep_len = (int)(r.meta & RL_META_STEP_MASK);
            stats_episode_end(bs, o, r.ep_ret, ep_len);
            if (a.terminal_obs) {
                a.terminal_obs[2 * i] = make_float4(o.obs[0], o.obs[1], o.obs[2], o.obs[3]);
                a.terminal_obs[2 * i + 1] = make_float4(o.obs[4], o.obs[5], o.obs[6], o.obs[7]);
            }
            if (a.episode_return) a.episode_return[i] = r.ep_ret;
            if (a.episode_length) a.episode_length[i] = ep_len;
            if (kAutoReset) {                    // SB3 VecEnv contract: return the reset observation
                float ax0, ay0;
                sample_reset(P, a.seed, (uint64_t)(a.gid0 + i), epoch, r, ax0, ay0);
                fresh_obs(P, r, ax0, ay0, o.obs);
                write_dry = true;
            }
        }
        store_rocket(a.st, i, r, write_dry);
        a.obs[2 * i] = make_float4(o.obs[0], o.obs[1], o.obs[2], o.obs[3]);
        a.obs[2 * i + 1] = make_float4(o.obs[4], o.obs[5], o.obs[6], o.obs[7]);
        a.reward[i] = reward;
        a.terminated[i] = o.terminated ? 1 : 0;
        a.truncated[i] = o.truncated ? 1 : 0;
        if (a.landing) a.landing[i] = (int8_t)o.landing;
    }
    stats_flush(bs, a.stats, reward_sum, steps_done);
    epoch_bump(a.epoch, a.ticket);
}

// RocketLandingEnv.reset (lander.py:168-186) for all / masked rockets.
__global__ void __launch_bounds__(kTile) reset_kernel(const DevParams P, StatePlanes st, const uint8_t* __restrict__ mask,
                                                      float4* __restrict__ obs, int64_t n, int64_t gid0,
                                                      uint64_t seed, unsigned long long* __restrict__ epoch_ctr,
                                                      unsigned int* __restrict__ ticket) {
    const uint64_t epoch = *epoch_ctr;
    for (int64_t i = (int64_t)blockIdx.x * kTile + threadIdx.x; i < n; i += (int64_t)gridDim.x * kTile) {
        if (mask && !mask[i]) continue;
        Rocket r;
        float ax0, ay0, o[8];
        sample_reset(P, seed, (uint64_t)(gid0 + i), epoch, r, ax0, ay0);
        store_rocket(st, i, r, true);
        if (obs) {
            fresh_obs(P, r, ax0, ay0, o);
            obs[2 * i] = make_float4(o[0], o[1], o[2], o[3]);
            obs[2 * i + 1] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
    epoch_bump(epoch_ctr, ticket);
}

// Public state format <-> planes (rl_set_state / rl_get_state).
__global__ void __launch_bounds__(kTile) set_state_kernel(StatePlanes st, const float* __restrict__ soa,
                                                          const uint8_t* __restrict__ mask, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kTile + threadIdx.x; i < n; i += (int64_t)gridDim.x * kTile) {
        if (mask && !mask[i]) continue;
        Rocket r;
        r.x = soa[RL_W_X * n + i]; r.y = soa[RL_W_Y * n + i]; r.ang = soa[RL_W_ANGLE * n + i];
        r.sx = soa[RL_W_SX * n + i]; r.sy = soa[RL_W_SY * n + i]; r.sang = soa[RL_W_SANGLE * n + i];
        r.dry = soa[RL_W_DRY * n + i]; r.fuel = soa[RL_W_FUEL * n + i];
        r.meta = __float_as_uint(soa[RL_W_META * n + i]); r.ep_ret = soa[RL_W_EPRET * n + i];
        store_rocket(st, i, r, true);
    }
}

__global__ void __launch_bounds__(kTile) get_state_kernel(StatePlanes st, float* __restrict__ soa, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * kTile + threadIdx.x; i < n; i += (int64_t)gridDim.x * kTile) {
        Rocket r;
        load_rocket(st, i, r);
        soa[RL_W_X * n + i] = r.x; soa[RL_W_Y * n + i] = r.y; soa[RL_W_ANGLE * n + i] = r.ang;
        soa[RL_W_SX * n + i] = r.sx; soa[RL_W_SY * n + i] = r.sy; soa[RL_W_SANGLE * n + i] = r.sang;
        soa[RL_W_DRY * n + i] = r.dry; soa[RL_W_FUEL * n + i] = r.fuel;
        soa[RL_W_META * n + i] = __uint_as_float(r.meta); soa[RL_W_EPRET * n + i] = r.ep_ret;
    }
}

// rl_stats: copy (and optionally clear) the accumulators on the stream.
__global__ void stats_copy_kernel(double* __restrict__ acc, double* __restrict__ out, int reset_after) {
    const int t = threadIdx.x;
    if (t < RL_STATS_WORDS) {
        out[t] = acc[t];
        if (reset_after) acc[t] = 0.0;
    }
}

}  // namespace rl
