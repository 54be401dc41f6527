// Kernels of the batched environment step.  State layout in HBM (DESIGN.md "Data layout"):
//
//   plane A  float4[Np]  { x, y, angle, meta }          read + written every step
//   plane B  float4[Np]  { sx, sy, sangle, ep_return }   read + written every step
//   fuel     float [Np]                                  read + written every step
//   dry      float [Np]                                  read every step, written on reset only
//
// Np = n_envs rounded up to a 256-rocket tile; all planes live in one buffer so a checkpoint is
// one tensor.  A tile's slice of every plane is one contiguous, 16-byte-aligned run of bytes
// (4 KiB / 4 KiB / 1 KiB / 1 KiB, plus 2 KiB of the action array), which is what lets the step
// kernel fetch it with five TMA bulk copies (cp.async.bulk, SASS UBLKCP) into a shared-memory
// ring instead of per-thread loads: the loads of the next kStages tiles are always in flight
// while the current tile is being computed, independent of occupancy and register pressure.
#pragma once
#include "rl_device.cuh"

namespace rl {

constexpr int kTile = 256;           // rockets per tile == threads per CTA
constexpr int kStages = 3;           // depth of the shared-memory ring

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
    uint32_t n;                           // rockets on this shard (< 2^31)
    int64_t gid0;                         // global id of rocket 0 (Philox counter)
    uint64_t seed;
    unsigned long long* __restrict__ epoch;   // device: launch epoch (Philox counter), bumped by the last CTA
    unsigned int* __restrict__ ticket;        // device: CTAs finished in this launch
    int substeps;
};

// ---- TMA bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS) ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy (1-D TMA); completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct __align__(128) TileStage {     // one tile's inputs, 12 KiB
    float4 A[kTile];
    float4 B[kTile];
    float fuel[kTile];
    float dry[kTile];
    float2 act[kTile];
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
// ---- slot per CTA at the end (the grid is persistent, so that is a few thousand atomics per launch)
struct BlockStats {
    double sum[RL_STATS_WORDS];
};

__device__ __forceinline__ void stats_init(BlockStats& s) {
    if (threadIdx.x < RL_STATS_WORDS) s.sum[threadIdx.x] = 0.0;
}

__device__ __forceinline__ void stats_flush(BlockStats& s, double* __restrict__ g, float reward_sum, int steps) {
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

__device__ __noinline__ void stats_episode_end(BlockStats& s, int landing, bool out_of_bounds, float ep_ret,
                                               int ep_len) {
    atomicAdd(&s.sum[RL_S_EPISODES], 1.0);
    atomicAdd(&s.sum[RL_S_RETURN_SUM], (double)ep_ret);
    atomicAdd(&s.sum[RL_S_LENGTH_SUM], (double)ep_len);
    if (landing) atomicAdd(&s.sum[RL_S_SAFE + landing - 1], 1.0);
    else if (out_of_bounds) atomicAdd(&s.sum[RL_S_OUT_OF_BOUNDS], 1.0);
    else atomicAdd(&s.sum[RL_S_TIME_LIMIT], 1.0);
}

__device__ __forceinline__ void load_rocket(const StatePlanes& st, uint32_t i, Rocket& r) {
    const float4 a = st.A[i];
    const float4 b = st.B[i];
    r.fuel = st.fuel[i];
    r.dry = st.dry[i];
    r.x = a.x; r.y = a.y; r.ang = a.z; r.meta = __float_as_uint(a.w);
    r.sx = b.x; r.sy = b.y; r.sang = b.z; r.ep_ret = b.w;
}

__device__ __forceinline__ void store_rocket(const StatePlanes& st, uint32_t i, const Rocket& r, bool write_dry) {
    st.A[i] = make_float4(r.x, r.y, r.ang, __uint_as_float(r.meta));
    st.B[i] = make_float4(r.sx, r.sy, r.sang, r.ep_ret);
    st.fuel[i] = r.fuel;
    if (write_dry) st.dry[i] = r.dry;
}

// Warp-cooperative reset: for every done lane (uniform loop over the ballot) lanes 0..2 compute the
// three Philox blocks of that rocket in parallel and the ten words are shuffled to the lane that
// owns it.  Same counters and words as sample_reset(), one third of the serial Philox latency, and
// the rest of the warp is not dragged through three back-to-back Philox evaluations.
__device__ __forceinline__ void warp_reset(const DevParams& P, unsigned done_mask, uint64_t seed, uint64_t gid_lane0,
                                           uint64_t epoch, Rocket& r, float* obs) {
    const unsigned lane = threadIdx.x & 31u;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    while (done_mask) {
        const unsigned src = (unsigned)__ffs((int)done_mask) - 1u;
        done_mask &= done_mask - 1u;
        const uint4 blk = philox4x32_10(reset_counter(gid_lane0 + src, epoch, lane & 3u), key);
        uint32_t w[10];
        w[0] = __shfl_sync(0xffffffffu, blk.x, 0); w[1] = __shfl_sync(0xffffffffu, blk.y, 0);
        w[2] = __shfl_sync(0xffffffffu, blk.z, 0); w[3] = __shfl_sync(0xffffffffu, blk.w, 0);
        w[4] = __shfl_sync(0xffffffffu, blk.x, 1); w[5] = __shfl_sync(0xffffffffu, blk.y, 1);
        w[6] = __shfl_sync(0xffffffffu, blk.z, 1); w[7] = __shfl_sync(0xffffffffu, blk.w, 1);
        w[8] = __shfl_sync(0xffffffffu, blk.x, 2); w[9] = __shfl_sync(0xffffffffu, blk.y, 2);
        if (lane == src) {
            float ax0, ay0;
            rocket_from_words(P, w, r, ax0, ay0);
            fresh_obs(P, r, ax0, ay0, obs);          // SB3 VecEnv contract: return the reset observation
        }
    }
}

// The environment step (lander.py:188-255 for every rocket) + VecEnv auto-reset.
// Persistent grid: CTA b walks tiles b, b + gridDim.x, ...; thread t of the CTA owns rocket
// tile * 256 + t.  Inputs arrive through the TMA ring, outputs leave as coalesced 128-bit stores.
template <bool kAutoReset>
__global__ void __launch_bounds__(kTile) step_kernel(const __grid_constant__ DevParams P,
                                                     const __grid_constant__ StepArgs a) {
    __shared__ TileStage ring[kStages];
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ BlockStats bs;

    const uint32_t tid = threadIdx.x;
    const uint32_t n_tiles = (a.n + kTile - 1) / kTile;
    const uint32_t full_tiles = a.n / kTile;           // tiles whose action slice is a whole 2 KiB
    const uint32_t stride = gridDim.x;

    auto issue = [&](uint32_t tile, uint32_t s) {      // one thread: five bulk copies into stage s
        const size_t base = (size_t)tile * kTile;
        const bool act_bulk = tile < full_tiles;
        mbar_arrive_expect_tx(&full_bar[s], 4096u + 4096u + 1024u + 1024u + (act_bulk ? 2048u : 0u));
        bulk_g2s(ring[s].A, a.st.A + base, 4096u, &full_bar[s]);
        bulk_g2s(ring[s].B, a.st.B + base, 4096u, &full_bar[s]);
        bulk_g2s(ring[s].fuel, a.st.fuel + base, 1024u, &full_bar[s]);
        bulk_g2s(ring[s].dry, a.st.dry + base, 1024u, &full_bar[s]);
        if (act_bulk) bulk_g2s(ring[s].act, a.actions + base, 2048u, &full_bar[s]);
    };

    stats_init(bs);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1u);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (uint32_t s = 0; s < (uint32_t)kStages; ++s) {
            const uint32_t tile = blockIdx.x + s * stride;
            if (tile < n_tiles) issue(tile, s);
        }
    }
    const uint64_t epoch = *a.epoch;
    float reward_sum = 0.0f;
    int steps_done = 0;

    uint32_t s = 0, parity = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += stride) {
        const uint32_t i = tile * kTile + tid;
        const bool valid = i < a.n;

        mbar_wait(&full_bar[s], parity);
        Rocket r;
        {
            const float4 va = ring[s].A[tid];
            const float4 vb = ring[s].B[tid];
            r.fuel = ring[s].fuel[tid];
            r.dry = ring[s].dry[tid];
            r.x = va.x; r.y = va.y; r.ang = va.z; r.meta = __float_as_uint(va.w);
            r.sx = vb.x; r.sy = vb.y; r.sang = vb.z; r.ep_ret = vb.w;
        }
        float2 act;
        if (tile < full_tiles) act = ring[s].act[tid];
        else act = valid ? __ldg(&a.actions[i]) : make_float2(0.0f, 0.0f);   // ragged last tile
        __syncthreads();                                 // every thread has drained stage s
        if (tid == 0) {
            const uint32_t next = tile + (uint32_t)kStages * stride;
            if (next < n_tiles) issue(next, s);
        }
        if (++s == (uint32_t)kStages) { s = 0; parity ^= 1u; }

        StepResult o;
        float reward = 0.0f;
        bool done = false;
        for (int k = 0; k < a.substeps; ++k) {
            rocket_step(P, r, act.x, act.y, o);
            reward += o.reward;
            if (valid) {
                ++steps_done;
                if (o.tipped) atomicAdd(&bs.sum[RL_S_TIP_STEPS], 1.0);
                if (o.nonfinite) atomicAdd(&bs.sum[RL_S_NONFINITE], 1.0);
            }
            done = valid && (o.terminated || o.truncated);
            if (done) break;                             // stop integrating at the first done
        }
        reward_sum += valid ? reward : 0.0f;

        if (done) {
            const int ep_len = (int)(r.meta & RL_META_STEP_MASK);
            stats_episode_end(bs, o.landing, o.out_of_bounds, r.ep_ret, ep_len);
            if (a.terminal_obs) {
                a.terminal_obs[2 * (size_t)i] = make_float4(o.obs[0], o.obs[1], o.obs[2], o.obs[3]);
                a.terminal_obs[2 * (size_t)i + 1] = make_float4(o.obs[4], o.obs[5], o.obs[6], o.obs[7]);
            }
            if (a.episode_return) a.episode_return[i] = r.ep_ret;
            if (a.episode_length) a.episode_length[i] = ep_len;
        }
        bool write_dry = false;
        if (kAutoReset) {
            const unsigned done_mask = __ballot_sync(0xffffffffu, done);
            if (done_mask) {
                warp_reset(P, done_mask, a.seed, (uint64_t)a.gid0 + (uint64_t)(i - (tid & 31u)), epoch, r, o.obs);
                write_dry = done;
            }
        }
        if (valid) {
            store_rocket(a.st, i, r, write_dry);
            a.obs[2 * (size_t)i] = make_float4(o.obs[0], o.obs[1], o.obs[2], o.obs[3]);
            a.obs[2 * (size_t)i + 1] = make_float4(o.obs[4], o.obs[5], o.obs[6], o.obs[7]);
            a.reward[i] = reward;
            a.terminated[i] = o.terminated ? 1 : 0;
            a.truncated[i] = o.truncated ? 1 : 0;
            if (a.landing) a.landing[i] = (int8_t)o.landing;
        }
    }
    stats_flush(bs, a.stats, reward_sum, steps_done);
    epoch_bump(a.epoch, a.ticket);
}

// RocketLandingEnv.reset (lander.py:168-186) for all / masked rockets.
__global__ void __launch_bounds__(kTile) reset_kernel(const __grid_constant__ DevParams P, StatePlanes st,
                                                      const uint8_t* __restrict__ mask, float4* __restrict__ obs,
                                                      uint32_t n, int64_t gid0, uint64_t seed,
                                                      unsigned long long* __restrict__ epoch_ctr,
                                                      unsigned int* __restrict__ ticket) {
    const uint64_t epoch = *epoch_ctr;
    for (uint32_t i = blockIdx.x * kTile + threadIdx.x; i < n; i += gridDim.x * kTile) {
        if (mask && !mask[i]) continue;
        Rocket r;
        float ax0, ay0, o[8];
        sample_reset(P, seed, (uint64_t)gid0 + i, epoch, r, ax0, ay0);
        store_rocket(st, i, r, true);
        if (obs) {
            fresh_obs(P, r, ax0, ay0, o);
            obs[2 * (size_t)i] = make_float4(o[0], o[1], o[2], o[3]);
            obs[2 * (size_t)i + 1] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
    epoch_bump(epoch_ctr, ticket);
}

// Public state format <-> planes (rl_set_state / rl_get_state).
__global__ void __launch_bounds__(kTile) set_state_kernel(StatePlanes st, const float* __restrict__ soa,
                                                          const uint8_t* __restrict__ mask, uint32_t n) {
    const size_t N = n;
    for (uint32_t i = blockIdx.x * kTile + threadIdx.x; i < n; i += gridDim.x * kTile) {
        if (mask && !mask[i]) continue;
        Rocket r;
        r.x = soa[RL_W_X * N + i]; r.y = soa[RL_W_Y * N + i]; r.ang = soa[RL_W_ANGLE * N + i];
        r.sx = soa[RL_W_SX * N + i]; r.sy = soa[RL_W_SY * N + i]; r.sang = soa[RL_W_SANGLE * N + i];
        r.dry = soa[RL_W_DRY * N + i]; r.fuel = soa[RL_W_FUEL * N + i];
        r.meta = __float_as_uint(soa[RL_W_META * N + i]); r.ep_ret = soa[RL_W_EPRET * N + i];
        store_rocket(st, i, r, true);
    }
}

__global__ void __launch_bounds__(kTile) get_state_kernel(StatePlanes st, float* __restrict__ soa, uint32_t n) {
    const size_t N = n;
    for (uint32_t i = blockIdx.x * kTile + threadIdx.x; i < n; i += gridDim.x * kTile) {
        Rocket r;
        load_rocket(st, i, r);
        soa[RL_W_X * N + i] = r.x; soa[RL_W_Y * N + i] = r.y; soa[RL_W_ANGLE * N + i] = r.ang;
        soa[RL_W_SX * N + i] = r.sx; soa[RL_W_SY * N + i] = r.sy; soa[RL_W_SANGLE * N + i] = r.sang;
        soa[RL_W_DRY * N + i] = r.dry; soa[RL_W_FUEL * N + i] = r.fuel;
        soa[RL_W_META * N + i] = __uint_as_float(r.meta); soa[RL_W_EPRET * N + i] = r.ep_ret;
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
