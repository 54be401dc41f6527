// Kernels of the batched environment step.  State layout in HBM (DESIGN.md "Data layout"):
//
//   plane A  16 B x Np  { x i32, y i32, angle i32, meta u32 }          read + written every step
//   plane B  16 B x Np  { sx i32, sy i32, sangle f32, ep_return f32 }  read + written every step
//   fuel     i32 [Np]                                                  read + written every step
//   dry      f32 [Np]                                                  read every step, written on reset only
// (the integer words are fixed point, rl_device.cuh "PRECISION"; the planes are declared float4 /
// float and the integer words travel as raw bits)
//
// Np = n_envs rounded up to 256; all planes live in one buffer so a checkpoint is one tensor.
// Thread t of a CTA owns rocket tile * 256 + t of the tile the CTA is working on, so every
// warp-wide access is one contiguous run (512 B in planes A / B, 128 B in fuel / dry, 256 B of
// actions, 1 KiB of observations).
//
// Inputs are STAGED THROUGH SHARED MEMORY BY ASYNCHRONOUS COPIES (cp.async, SASS LDGSTS): each
// iteration a thread waits for the copies of ITS OWN rocket of the current tile, reads them back,
// and immediately issues the copies of the NEXT tile into the other half of a two-stage ring.  The
// next tile's bytes are therefore always in flight while the current one is computed, without
// holding registers for them.  Because a thread only ever reads back what it copied itself there is no
// barrier, mbarrier or proxy fence in the loop, and warps never wait for one another (a warp on
// the long reset path does not stall its CTA).  profiles/README.md records why this replaced the
// two TMA (cp.async.bulk) designs tried first.
//
// Grid: large launches give every CTA `chunk` CONSECUTIVE tiles (8 x 256 rockets) and let the
// hardware dispatch CTAs in address order, so the rockets being streamed at any moment form one
// compact, advancing window of every plane; a persistent grid striding by gridDim.x (kept for
// launches too small to fill six waves) scatters that window and loses ~20 % of HBM bandwidth.
#pragma once
#include <type_traits>

#include "rl_device.cuh"
#include "rl_launch.cuh"

namespace rl {

#ifndef RL_CTA_THREADS
#define RL_CTA_THREADS 256
#endif
constexpr int kCtaThreads = RL_CTA_THREADS;   // rockets per tile == threads per CTA
#ifndef RL_STAGES
#define RL_STAGES 2
#endif
constexpr int kStages = RL_STAGES;   // depth of the shared-memory ring (tiles in flight = kStages - 1)
constexpr int kPad = 256;            // planes are padded to a multiple of this many rockets
#ifndef RL_MIN_CTAS
#define RL_MIN_CTAS 3                // lets ptxas use 64 registers, with which 4 CTAs are still resident
#endif
struct StatePlanes {
    float4* A;
    float4* B;
    float* fuel;
    float* dry;
};

__host__ __device__ inline int64_t padded_envs(int64_t n) { return (n + kPad - 1) / kPad * kPad; }

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
    float2* __restrict__ raw_accel;       // nullable: unclipped (ax, ay) of the step (info["raw_state"], lander.py:157-166)
    double* __restrict__ stats;           // [RL_STATS_WORDS]
    uint32_t n;                           // rockets on this shard (< 2^31)
    int64_t gid0;                         // global id of rocket 0 (Philox counter)
    uint64_t seed;
    const unsigned long long* __restrict__ epoch;   // device: launch epoch to READ (Philox counter)
    unsigned long long* __restrict__ epoch_bump;    // device: epoch to advance when the launch ends, or null
    unsigned int* __restrict__ ticket;              // device: CTAs finished in this launch
    uint32_t tile_begin, tile_end;                  // tiles of 256 rockets this launch steps ([0, ceil(n/256)) = all)
    uint32_t optional;                              // bit 0: landing != null, bit 1: raw_accel != null (tested per tile:
                                                    // one bit test on a uniform register instead of two 64-bit compares)
    uint32_t chunk;                                 // consecutive tiles per CTA (grid = ceil(tiles / chunk)); 0 = persistent
                                                    // grid, CTA b walks tiles b, b + gridDim.x, ...
    int substeps;
};

// ---- asynchronous global -> shared copies (PTX cp.async; SASS LDGSTS) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {      // L2 only (.cg): streamed once
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// copies src_bytes (0 or 8) and zero-fills the rest of the 8 bytes
__device__ __forceinline__ void cp_async_8_zfill(void* dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

struct __align__(16) Stage {          // one tile's inputs, 12 KiB; thread t uses element t of every array
    float4 A[kCtaThreads];
    float4 B[kCtaThreads];
    float2 act[kCtaThreads];
    float fuel[kCtaThreads];
    float dry[kCtaThreads];
};

// The launch epoch lives in device memory so that a CUDA-graph replay of the step advances the
// Philox stream like an ordinary launch does.  Every CTA reads it on entry; the last CTA to
// finish (ticket) bumps it for the next launch, which is stream-ordered after this one.
// The step kernel splits this in two: a CTA takes its ticket as soon as all its threads have read
// the epoch (the atomic's latency hides behind the CTA's work) and the holder of the last ticket
// advances the epoch when it finishes -- by then every CTA of the launch has read the old value.
__device__ __forceinline__ unsigned int epoch_take_ticket(unsigned int* ticket) {
    __threadfence();                                   // this CTA's epoch reads (ordered by the barrier before) first
    return atomicAdd(ticket, 1u);
}
__device__ __forceinline__ void epoch_bump_if_last(unsigned long long* epoch, unsigned int* ticket, unsigned int mine) {
    if (mine == gridDim.x - 1) {
        *ticket = 0u;
        *epoch += 1ull;
    }
}
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

// ---- statistics: every thread keeps its own running sums in registers; once per CTA they are
// ---- reduced by warp shuffles into shared memory and flushed with one global atomic per slot
// ---- (the grid is persistent, so that is a few thousand global atomics per launch).  Only the
// ---- per-class episode counts use shared-memory atomics, and those are native integer adds
// ---- (a double atomicAdd on shared memory is a CAS loop -- far too slow for a per-step event).
// The global accumulators are kStatSlots copies of the RL_STATS_WORDS vector (CTA b adds into copy
// b mod kStatSlots; rl_stats sums them), so that a launch of tens of thousands of CTAs does not
// serialise its reductions on thirteen addresses.
constexpr int kStatSlots = 64;

struct ThreadStats {
    float reward_sum = 0.0f, return_sum = 0.0f;
    int env_steps = 0, tip_steps = 0, nonfinite = 0, episodes = 0, length_sum = 0;
};

struct BlockStats {
    double sum[RL_STATS_WORDS];
    unsigned int cls[8];            // [1..4] landing class, [5] out of bounds, [6] time limit
};

__device__ __forceinline__ void stats_init(BlockStats& s) {
    if (threadIdx.x < RL_STATS_WORDS) s.sum[threadIdx.x] = 0.0;
    if (threadIdx.x < 8) s.cls[threadIdx.x] = 0u;
}

__device__ __forceinline__ void stats_flush(BlockStats& s, double* __restrict__ g, ThreadStats ts) {
    double reward_sum = (double)ts.reward_sum, return_sum = (double)ts.return_sum;   // once per launch: reduce in fp64
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        reward_sum += __shfl_xor_sync(0xffffffffu, reward_sum, o);
        return_sum += __shfl_xor_sync(0xffffffffu, return_sum, o);
        ts.env_steps += __shfl_xor_sync(0xffffffffu, ts.env_steps, o);
        ts.tip_steps += __shfl_xor_sync(0xffffffffu, ts.tip_steps, o);
        ts.nonfinite += __shfl_xor_sync(0xffffffffu, ts.nonfinite, o);
        ts.episodes += __shfl_xor_sync(0xffffffffu, ts.episodes, o);
        ts.length_sum += __shfl_xor_sync(0xffffffffu, ts.length_sum, o);
    }
    if ((threadIdx.x & 31) == 0) {          // 8 warps x 7 slots, once per CTA per launch
        atomicAdd(&s.sum[RL_S_REWARD_SUM], reward_sum);
        atomicAdd(&s.sum[RL_S_RETURN_SUM], return_sum);
        atomicAdd(&s.sum[RL_S_ENV_STEPS], (double)ts.env_steps);
        atomicAdd(&s.sum[RL_S_TIP_STEPS], (double)ts.tip_steps);
        atomicAdd(&s.sum[RL_S_NONFINITE], (double)ts.nonfinite);
        atomicAdd(&s.sum[RL_S_EPISODES], (double)ts.episodes);
        atomicAdd(&s.sum[RL_S_LENGTH_SUM], (double)ts.length_sum);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s.sum[RL_S_SAFE] = (double)s.cls[1];
        s.sum[RL_S_GOOD] = (double)s.cls[2];
        s.sum[RL_S_OK] = (double)s.cls[3];
        s.sum[RL_S_UNSAFE] = (double)s.cls[4];
        s.sum[RL_S_OUT_OF_BOUNDS] = (double)s.cls[5];
        s.sum[RL_S_TIME_LIMIT] = (double)s.cls[6];
    }
    __syncthreads();
    if (threadIdx.x < RL_STATS_WORDS && s.sum[threadIdx.x] != 0.0)
        atomicAdd(&g[(blockIdx.x % kStatSlots) * RL_STATS_WORDS + threadIdx.x], s.sum[threadIdx.x]);
}

__device__ __forceinline__ void unpack_rocket(const float4& a, const float4& b, float fuel_bits, float dry, Rocket& r) {
    r.x = __float_as_int(a.x); r.y = __float_as_int(a.y); r.ang = __float_as_int(a.z); r.meta = __float_as_uint(a.w);
    r.sx = __float_as_int(b.x); r.sy = __float_as_int(b.y); r.sang = b.z; r.ep_ret = b.w;
    r.fuel = __float_as_int(fuel_bits);
    r.dry = dry;
}

__device__ __forceinline__ void load_rocket(const StatePlanes& st, uint32_t i, Rocket& r) {
    unpack_rocket(st.A[i], st.B[i], st.fuel[i], st.dry[i], r);
}

__device__ __forceinline__ void store_rocket(const StatePlanes& st, uint32_t i, const Rocket& r, bool write_dry) {
    st.A[i] = make_float4(__int_as_float(r.x), __int_as_float(r.y), __int_as_float(r.ang), __uint_as_float(r.meta));
    st.B[i] = make_float4(__int_as_float(r.sx), __int_as_float(r.sy), r.sang, r.ep_ret);
    st.fuel[i] = __int_as_float(r.fuel);
    if (write_dry) st.dry[i] = r.dry;
}

// Stores of the per-step OUTPUTS (observation, reward, flags).  They are written once and read by
// somebody else later; RL_STREAMING_STORES marks them evict-first (st.global.cs) so that they do
// not displace the state planes from L2 -- which only matters for batches whose state fits L2.
template <class T>
__device__ __forceinline__ void out_store(T* p, T v) {
#ifdef RL_STREAMING_STORES
    __stcs(p, v);
#else
    *p = v;
#endif
}

// Observation rows of the warp's 32 rockets (row i = 32 B at obs + 2 i).  Every lane holds one
// row in registers; written directly that is two 128-bit stores per lane, each instruction
// covering HALF of every one of 32 sectors.  Instead the warp transposes through 1 KiB of shared
// memory (2 STS.128 + 2 LDS.128, 16-byte chunks XOR-swizzled so neither side has bank conflicts)
// and stores two contiguous 512-byte runs: every instruction writes four whole 128-byte lines
// (+3.5 % bandwidth on the whole kernel, profiles/README.md).
struct __align__(16) ObsStage { float4 chunk[64]; };      // one warp's 32 rows of 2 x 16 B

__device__ __forceinline__ void store_obs(ObsStage& st, float4* __restrict__ obs, uint32_t i, uint32_t n, const float* o) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t w = (lane >> 2) & 1u;                  // chunk c lives at c ^ ((c >> 3) & 1); c = 2 lane (+1)
    st.chunk[(2u * lane) ^ w] = make_float4(o[0], o[1], o[2], o[3]);
    st.chunk[(2u * lane + 1u) ^ w] = make_float4(o[4], o[5], o[6], o[7]);
    __syncwarp();
    const uint32_t c = lane ^ ((lane >> 3) & 1u);
    const float4 v1 = st.chunk[c], v2 = st.chunk[32u + c];
    __syncwarp();                                         // the stage is rewritten next iteration
    const uint32_t row0 = i - lane;                       // the warp's first rocket
    const uint32_t r1 = row0 + (lane >> 1), r2 = r1 + 16u;
    if (r1 < n) out_store(&obs[2 * (size_t)row0 + lane], v1);
    if (r2 < n) out_store(&obs[2 * (size_t)row0 + 32u + lane], v2);
}

// Episode end + reset of the done lanes of one warp (rare: ~1 rocket in 113 per step; inlined
// into a cold branch -- an out-of-line call would force the rocket's registers onto the stack).
// Records the finished episode, hands the terminal observation / return / length to the
// caller's buffers, and -- with auto-reset -- resamples the rocket.  Lanes compute the three
// Philox4x32-10 blocks of a done rocket side by side and the ten words are shuffled to the lane that
// owns it: same counters and words as sample_reset(), one third of the serial Philox latency.
//   kManyDone = false (one substep per launch: ~1.3 done rockets in the 27 % of warps that have any):
//     the done lanes are walked one by one, lanes 0..2 serving each -- fewest registers, and the
//     single-step kernel must stay at 64 to keep four CTAs per SM resident;
//   kManyDone = true (fused substeps: K times as many rockets finish per launch): up to TEN done
//     rockets per pass, the j-th served by lanes 3j, 3j + 1, 3j + 2 -- a warp instruction costs one
//     issue slot however many lanes it serves, and this kernel is issue-bound.
template <bool kAutoReset, bool kManyDone, class PP>
__device__ __forceinline__ void warp_episode_end(const PP& P, const StepArgs& a, BlockStats& bs, ThreadStats& ts,
                                                 unsigned done_mask, bool done, uint32_t i, uint64_t epoch,
                                                 Rocket& r, StepResult& o) {
    if (done) {
        const int ep_len = (int)(r.meta & RL_META_STEP_MASK);
        ts.episodes += 1;
        ts.return_sum += r.ep_ret;
        ts.length_sum += ep_len;
        atomicAdd(&bs.cls[o.landing ? o.landing : (o.out_of_bounds ? 5 : 6)], 1u);
        if (a.terminal_obs) {
            a.terminal_obs[2 * (size_t)i] = make_float4(o.obs[0], o.obs[1], o.obs[2], o.obs[3]);
            a.terminal_obs[2 * (size_t)i + 1] = make_float4(o.obs[4], o.obs[5], o.obs[6], o.obs[7]);
        }
        if (a.episode_return) a.episode_return[i] = r.ep_ret;
        if (a.episode_length) a.episode_length[i] = ep_len;
    }
    if (kAutoReset) {
        const unsigned lane = threadIdx.x & 31u;
        const uint64_t gid_lane0 = (uint64_t)a.gid0 + (uint64_t)(i - lane);
        const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
        if constexpr (!kManyDone) {
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
                    rocket_from_words(P, w, r, o.obs);       // SB3 VecEnv contract: return the reset observation
                    a.st.dry[i] = r.dry;                     // dry mass only ever changes here
                }
            }
        } else {
            const unsigned group = lane / 3u, member = lane - 3u * group;      // lanes 30, 31: group 10, idle
            unsigned pending = done_mask;
            while (pending) {
                const int served = min(__popc(pending), 10);
                // the rocket this lane's group serves: the group-th set bit of `pending`
                unsigned m = pending;
                for (int t = 0; t + 1 < served; ++t)                            // (uniform trip count)
                    if ((unsigned)t < group) m &= m - 1u;
                const unsigned src = (unsigned)__ffs((int)m) - 1u;
                uint4 blk = make_uint4(0u, 0u, 0u, 0u);
                if (group < (unsigned)served) blk = philox4x32_10(reset_counter(gid_lane0 + src, epoch, member), key);
                // my own rocket, if it is served in this pass, sits in group `rank`
                const unsigned rank = (unsigned)__popc(pending & ((1u << lane) - 1u));
                const bool mine = ((pending >> lane) & 1u) != 0u && rank < 10u;
                const int g0 = (int)(3u * (rank < 10u ? rank : 0u));
                uint32_t w[10];
                w[0] = __shfl_sync(0xffffffffu, blk.x, g0); w[1] = __shfl_sync(0xffffffffu, blk.y, g0);
                w[2] = __shfl_sync(0xffffffffu, blk.z, g0); w[3] = __shfl_sync(0xffffffffu, blk.w, g0);
                w[4] = __shfl_sync(0xffffffffu, blk.x, g0 + 1); w[5] = __shfl_sync(0xffffffffu, blk.y, g0 + 1);
                w[6] = __shfl_sync(0xffffffffu, blk.z, g0 + 1); w[7] = __shfl_sync(0xffffffffu, blk.w, g0 + 1);
                w[8] = __shfl_sync(0xffffffffu, blk.x, g0 + 2); w[9] = __shfl_sync(0xffffffffu, blk.y, g0 + 2);
                if (mine) {
                    rocket_from_words(P, w, r, o.obs);
                    a.st.dry[i] = r.dry;
                }
                // drop the served (lowest) set bits
                for (int t = 0; t < served; ++t) pending &= pending - 1u;
            }
        }
    }
}

template <class PP> struct ParamSel;
template <> struct ParamSel<DevParams> {
    static __device__ __forceinline__ const DevParams& get(const DevParams& rt) { return rt; }
};
template <> struct ParamSel<DefaultParams> {
    static __device__ __forceinline__ DefaultParams get(const DevParams&) { return DefaultParams{}; }
};

// The environment step (lander.py:188-255 for every rocket) + VecEnv auto-reset.
//   kAutoReset   reset done rockets in the same launch (scripts/train.py:82-88 contract)
//   kSingleStep  substeps == 1 (the reference's step): no substep loop
//   kChunked     CTA b steps the a.chunk consecutive tiles from b * a.chunk (large launches);
//                otherwise persistent grid, CTA b walks tiles b, b + gridDim.x, ...
//   PP           DevParams (runtime constants) or DefaultParams (reference config.yaml folded in)
template <bool kAutoReset, bool kSingleStep, bool kChunked, class PP>
__global__ void __launch_bounds__(kCtaThreads, RL_MIN_CTAS) step_kernel(const __grid_constant__ DevParams Prt,
                                                                        const __grid_constant__ StepArgs a) {
    __shared__ Stage ring[kStages];
    __shared__ BlockStats bs;
    __shared__ ObsStage obs_stage[kCtaThreads / 32];

    rl_grid_dependency_wait();                             // (programmatic dependent launch: rl_launch.cuh)
    const auto& P = ParamSel<PP>::get(Prt);
    const uint32_t tid = threadIdx.x;
    const uint32_t stride = kChunked ? 1u : gridDim.x;
    const uint32_t tile0 = a.tile_begin + (kChunked ? blockIdx.x * a.chunk : blockIdx.x);   // this CTA's first tile
    const uint32_t n_tiles = kChunked ? min(a.tile_end, tile0 + a.chunk) : a.tile_end;       // one past its last

    // this thread's rocket of `tile` -> stage s (state planes are padded, so only the caller's
    // action array needs a bounds guard: out-of-range rockets get a zero action)
    auto prefetch = [&](uint32_t tile, uint32_t s) {
        const uint32_t j = tile * kCtaThreads + tid;
        cp_async_16(&ring[s].A[tid], a.st.A + j);
        cp_async_16(&ring[s].B[tid], a.st.B + j);
        cp_async_4(&ring[s].fuel[tid], a.st.fuel + j);
        cp_async_4(&ring[s].dry[tid], a.st.dry + j);
        const bool in = j < a.n;
        cp_async_8_zfill(&ring[s].act[tid], a.actions + (in ? j : 0u), in ? 8u : 0u);
    };

    stats_init(bs);
#pragma unroll
    for (uint32_t d = 0; d + 1 < (uint32_t)kStages; ++d) {   // prologue: the first kStages - 1 tiles
        const uint32_t tile = tile0 + d * stride;
        if (tile < n_tiles) prefetch(tile, d);
        cp_async_commit();
    }
    const uint64_t epoch = *a.epoch;
    __syncthreads();                                    // stats_init visible CTA-wide; every thread has read the epoch
    unsigned int my_ticket = 0u;
    if (a.epoch_bump && tid == 0) my_ticket = epoch_take_ticket(a.ticket);   // uniform over the grid
    ThreadStats ts;

    uint32_t s = 0;
    for (uint32_t tile = tile0; tile < n_tiles; tile += stride) {
        const uint32_t i = tile * kCtaThreads + tid;
        const bool valid = i < a.n;

        Rocket r;
        float2 act;
        {
            cp_async_wait<kStages - 2>();                   // this tile's group has landed
            unpack_rocket(ring[s].A[tid], ring[s].B[tid], ring[s].fuel[tid], ring[s].dry[tid], r);
            act = ring[s].act[tid];
            // the shared-memory reads above are complete (their values are in registers) before the
            // copies below are issued by the same thread, so refilling a stage is hazard-free; issuing
            // them AFTER the read-back keeps the LDS from queueing behind five LDGSTS in the MIO pipe
            const uint32_t next = tile + (uint32_t)(kStages - 1) * stride;
            const uint32_t s_next = (kStages == 2) ? (s ^ 1u) : ((s == 0u) ? (uint32_t)(kStages - 1) : s - 1u);
            if (next < n_tiles) prefetch(next, s_next);     // kStages - 1 tiles in flight while this one is computed
            cp_async_commit();                              // (possibly empty) keeps the group count uniform
            s = (s + 1u == (uint32_t)kStages) ? 0u : s + 1u;
        }

        StepResult o;
        Flight f;
        float reward;
        bool done;
        flight_begin(P, r, f);
        if constexpr (kSingleStep) {
            rocket_advance(P, r, f, act.x, act.y, o);
            reward = o.reward;
            // (padding rockets beyond n hold the all-zero state rl_create wrote and nothing ever
            // overwrites: massless, never tipped, never non-finite -- only the counts need `valid`)
            ts.env_steps += valid ? 1 : 0;
            ts.tip_steps += o.tipped ? 1 : 0;
            ts.nonfinite += o.nonfinite ? 1 : 0;
            done = valid && (o.terminated || o.truncated);
        } else {
            reward = 0.0f;
            done = false;
            int tipped = 0, bad = 0, k = 0;
            for (; k < a.substeps && !done; ++k) {       // stop integrating at the first done
                rocket_advance(P, r, f, act.x, act.y, o);  // the before-state (and its potential) is carried in f
                reward += o.reward;
                tipped += o.tipped ? 1 : 0;
                bad += o.nonfinite ? 1 : 0;
                done = valid && (o.terminated || o.truncated);
            }
            ts.env_steps += valid ? k : 0;
            ts.tip_steps += tipped;
            ts.nonfinite += bad;
        }
        flight_end(r, f);
        step_obs(P, r, f, o);                            // the observation is only needed once
        ts.reward_sum += valid ? reward : 0.0f;

        const unsigned done_mask = __ballot_sync(0xffffffffu, done);
        const bool terminated = o.terminated, truncated = o.truncated;
        const int landing = o.landing;
        if (done_mask) warp_episode_end<kAutoReset, !kSingleStep>(P, a, bs, ts, done_mask, done, i, epoch, r, o);

        store_obs(obs_stage[tid >> 5], a.obs, i, a.n, o.obs);
        if (valid) {
            store_rocket(a.st, i, r, false);
            out_store(&a.reward[i], reward);
            out_store(&a.terminated[i], (uint8_t)(terminated ? 1 : 0));
            out_store(&a.truncated[i], (uint8_t)(truncated ? 1 : 0));
            if (a.optional) {                                  // rarely asked for: one uniform test on the common path
                if (a.optional & 1u) out_store(reinterpret_cast<uint8_t*>(&a.landing[i]), (uint8_t)landing);
                if (a.optional & 2u) out_store(&a.raw_accel[i], make_float2(o.ax, o.ay));
            }
        }
    }
    cp_async_wait<0>();
    stats_flush(bs, a.stats, ts);
    if (a.epoch_bump && tid == 0) epoch_bump_if_last(a.epoch_bump, a.ticket, my_ticket);
}

// RocketLandingEnv.reset (lander.py:168-186) for all / masked rockets.
__global__ void __launch_bounds__(kCtaThreads) reset_kernel(const __grid_constant__ DevParams P, StatePlanes st,
                                                            const uint8_t* __restrict__ mask,
                                                            float4* __restrict__ obs, uint32_t n, int64_t gid0,
                                                            uint64_t seed, unsigned long long* __restrict__ epoch_ctr,
                                                            unsigned int* __restrict__ ticket) {
    const uint64_t epoch = *epoch_ctr;
    for (uint32_t i = blockIdx.x * kCtaThreads + threadIdx.x; i < n; i += gridDim.x * kCtaThreads) {
        if (mask && !mask[i]) continue;
        Rocket r;
        float o[8];
        sample_reset(P, seed, (uint64_t)gid0 + i, epoch, r, o);
        store_rocket(st, i, r, true);
        if (obs) {
            obs[2 * (size_t)i] = make_float4(o[0], o[1], o[2], o[3]);
            obs[2 * (size_t)i + 1] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
    epoch_bump(epoch_ctr, ticket);
}

// Public state format <-> planes (rl_set_state / rl_get_state): the public words ARE the kernel's
// words (raw 32-bit copies), so a get -> set round trip is lossless.
__global__ void __launch_bounds__(kCtaThreads) set_state_kernel(StatePlanes st, const uint32_t* __restrict__ words,
                                                                const uint8_t* __restrict__ mask, uint32_t n) {
    const size_t N = n;
    const float* soa = reinterpret_cast<const float*>(words);
    for (uint32_t i = blockIdx.x * kCtaThreads + threadIdx.x; i < n; i += gridDim.x * kCtaThreads) {
        if (mask && !mask[i]) continue;
        st.A[i] = make_float4(soa[RL_W_X * N + i], soa[RL_W_Y * N + i], soa[RL_W_ANGLE * N + i], soa[RL_W_META * N + i]);
        st.B[i] = make_float4(soa[RL_W_SX * N + i], soa[RL_W_SY * N + i], soa[RL_W_SANGLE * N + i], soa[RL_W_EPRET * N + i]);
        st.fuel[i] = soa[RL_W_FUEL * N + i];
        st.dry[i] = soa[RL_W_DRY * N + i];
    }
}

__global__ void __launch_bounds__(kCtaThreads) get_state_kernel(StatePlanes st, uint32_t* __restrict__ words, uint32_t n) {
    const size_t N = n;
    float* soa = reinterpret_cast<float*>(words);
    for (uint32_t i = blockIdx.x * kCtaThreads + threadIdx.x; i < n; i += gridDim.x * kCtaThreads) {
        const float4 a = st.A[i], b = st.B[i];
        soa[RL_W_X * N + i] = a.x; soa[RL_W_Y * N + i] = a.y; soa[RL_W_ANGLE * N + i] = a.z; soa[RL_W_META * N + i] = a.w;
        soa[RL_W_SX * N + i] = b.x; soa[RL_W_SY * N + i] = b.y; soa[RL_W_SANGLE * N + i] = b.z; soa[RL_W_EPRET * N + i] = b.w;
        soa[RL_W_FUEL * N + i] = st.fuel[i]; soa[RL_W_DRY * N + i] = st.dry[i];
    }
}

// rl_stats: sum the kStatSlots copies of the accumulators into out[RL_STATS_WORDS] (and optionally
// clear them), on the stream.  One CTA of RL_STATS_WORDS x 32 threads: thread (w, l) sums word w of
// copies l, l + 32, ...; a warp shuffle finishes each word.
__global__ void __launch_bounds__(RL_STATS_WORDS * 32) stats_copy_kernel(double* __restrict__ acc, double* __restrict__ out,
                                                                          int reset_after) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    double v = 0.0;
    for (int sl = l; sl < kStatSlots; sl += 32) {
        v += acc[sl * RL_STATS_WORDS + w];
        if (reset_after) acc[sl * RL_STATS_WORDS + w] = 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (l == 0) out[w] = v;
}

}  // namespace rl
