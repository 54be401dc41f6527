// Second attainable-bandwidth probe for the step kernel's access pattern (no physics).  It asks
// which properties of the PATTERN cost bandwidth, so that the step kernel can be reshaped:
//   * calibration: read-only, write-only and copy kernels (what an SM kernel can reach at all);
//   * the nine-stream mix (48 B read, 75 B written per rocket) with
//       T  observation rows written as whole 32-byte sectors per instruction (lane-pair exchange),
//       S  state in tile-major order (one contiguous 10 KiB block per 256 rockets) instead of planes,
//       N  non-persistent grid (one tile per CTA),
//       H  cache hints: loads evict-first (.cs), outputs evict-first (.cs),
//       F- without the three byte-flag streams, R- without the reward stream (marginal costs).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a stream_mix2.cu -o stream_mix2
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(256) copy_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = in[i];
}
// torch-style: one CTA per 256 x 4 elements, four independent 128-bit loads per thread
__global__ void __launch_bounds__(256) copy4_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    const size_t base = blockIdx.x * 1024ull + threadIdx.x;
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) if (base + u * 256 < n) v[u] = in[base + u * 256];
#pragma unroll
    for (int u = 0; u < 4; ++u) if (base + u * 256 < n) out[base + u * 256] = v[u];
}
__global__ void __launch_bounds__(256) read_kernel(const float4* __restrict__ in, float* __restrict__ out, size_t n) {
    float acc = 0.f;
    size_t i = blockIdx.x * 256ull + threadIdx.x;
    const size_t st = (size_t)gridDim.x * 256;
    for (; i + 3 * st < n; i += 4 * st) {
        const float4 a = in[i], b = in[i + st], c = in[i + 2 * st], d = in[i + 3 * st];
        acc += a.x + b.y + c.z + d.w;
    }
    for (; i < n; i += st) acc += in[i].x;
    if (acc == 12345.678f) out[0] = acc;
}
__global__ void __launch_bounds__(256) write_kernel(float4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
        out[i] = make_float4(1.f, 2.f, 3.f, (float)i);
}

struct Bufs {
    float4 *A, *B, *obs; float *fuel, *dry, *rew; const float2* act; uint8_t *f0, *f1, *f2;
    char* tiled;   // tile-major state: per 256 rockets { float4 A[256]; float4 B[256]; float fuel[256]; float dry[256]; }
};

template <int kHint, class T> __device__ __forceinline__ T ld(const T* p) {
    if constexpr (kHint == 1) return __ldcs(p);
    else if constexpr (kHint == 2) return __ldcg(p);
    else return *p;
}
template <int kHint, class T> __device__ __forceinline__ void st(T* p, T v) {
    if constexpr (kHint == 1) __stcs(p, v);
    else if constexpr (kHint == 2) __stcg(p, v);
    else if constexpr (kHint == 3) __stwt(p, v);
    else *p = v;
}

template <int kUnroll, bool kObsT, bool kTiled, bool kPersist, int kLdHint, int kStHint, bool kFlags, bool kReward>
__global__ void __launch_bounds__(256) mix_kernel(Bufs b, uint32_t n) {
    const uint32_t stride = kPersist ? gridDim.x * 256u * kUnroll : 0xffffffffu;
    for (uint32_t base = blockIdx.x * 256u * kUnroll + threadIdx.x; base < n; base += stride) {
        float4 a[kUnroll], bb[kUnroll]; float fu[kUnroll], d[kUnroll]; float2 ac[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t i = base + u * 256u;
            if (i < n) {
                if constexpr (kTiled) {
                    const char* t = b.tiled + (size_t)(i >> 8) * 10240u;
                    const uint32_t l = i & 255u;
                    a[u] = ld<kLdHint>(reinterpret_cast<const float4*>(t) + l);
                    bb[u] = ld<kLdHint>(reinterpret_cast<const float4*>(t + 4096) + l);
                    fu[u] = ld<kLdHint>(reinterpret_cast<const float*>(t + 8192) + l);
                    d[u] = ld<kLdHint>(reinterpret_cast<const float*>(t + 9216) + l);
                } else {
                    a[u] = ld<kLdHint>(b.A + i); bb[u] = ld<kLdHint>(b.B + i);
                    fu[u] = ld<kLdHint>(b.fuel + i); d[u] = ld<kLdHint>(b.dry + i);
                }
                ac[u] = ld<kLdHint>(b.act + i);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t i = base + u * 256u;
            if (i < n) {      // n is a multiple of 256: uniform per warp, shuffles below are safe
                const float s = a[u].x + bb[u].y + fu[u] * d[u] + ac[u].x * ac[u].y;
                const float4 nA = make_float4(a[u].y, a[u].x, a[u].w + s, a[u].z);
                const float4 nB = make_float4(bb[u].y, bb[u].x, bb[u].w, bb[u].z + s);
                if constexpr (kTiled) {
                    char* t = b.tiled + (size_t)(i >> 8) * 10240u;
                    const uint32_t l = i & 255u;
                    reinterpret_cast<float4*>(t)[l] = nA;
                    reinterpret_cast<float4*>(t + 4096)[l] = nB;
                    reinterpret_cast<float*>(t + 8192)[l] = fu[u] + 1.0f;
                } else {
                    b.A[i] = nA; b.B[i] = nB; b.fuel[i] = fu[u] + 1.0f;
                }
                const float4 lo = make_float4(s, a[u].x, a[u].y, a[u].z), hi = make_float4(bb[u].x, bb[u].y, bb[u].z, s);
                if constexpr (kObsT) {
                    const bool odd = (threadIdx.x & 1u) != 0u;
                    const float4 snd = odd ? lo : hi;
                    float4 rcv;
                    rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, 1); rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, 1);
                    rcv.z = __shfl_xor_sync(0xffffffffu, snd.z, 1); rcv.w = __shfl_xor_sync(0xffffffffu, snd.w, 1);
                    const uint32_t er = i & ~1u, orow = i | 1u;
                    st<kStHint>(&b.obs[2 * (size_t)er + (odd ? 1u : 0u)], odd ? rcv : lo);
                    st<kStHint>(&b.obs[2 * (size_t)orow + (odd ? 1u : 0u)], odd ? hi : rcv);
                } else {
                    st<kStHint>(&b.obs[2 * (size_t)i], lo);
                    st<kStHint>(&b.obs[2 * (size_t)i + 1], hi);
                }
                if constexpr (kReward) st<kStHint>(&b.rew[i], s);
                if constexpr (kFlags) {
                    st<kStHint>(&b.f0[i], (uint8_t)(s > 0.0f)); st<kStHint>(&b.f1[i], (uint8_t)(s < 1.0f));
                    st<kStHint>(&b.f2[i], (uint8_t)(s == 2.0f));
                }
            }
        }
        if (!kPersist) break;
    }
}


// The step kernel's actual skeleton: `chunk` consecutive tiles per CTA, per-thread cp.async
// two-stage ring (read back, then prefetch the next tile), observation rows transposed through
// shared memory, and `work` dependent FMAs per rocket standing in for the physics.
struct __align__(16) RingStage { float4 A[256]; float4 B[256]; float2 act[256]; float fuel[256]; float dry[256]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int kBytes> __device__ __forceinline__ void cpa(void* dst, const void* src) {
    if constexpr (kBytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst)), "l"(src), "n"(kBytes) : "memory");
}
// div_pct: percentage of warp-tiles that execute div_work extra dependent FMAs (the warps that meet a
// finished episode: Philox reset path); epilogue: the step kernel's per-CTA statistics flush (35
// shuffles, shared-memory atomics, two barriers, 13 global atomics)
template <int kStages>
__global__ void __launch_bounds__(256, 4) mix_ring(Bufs b, uint32_t n_tiles_total, uint32_t chunk, int work,
                                                   int div_pct = 0, int div_work = 0, int epilogue = 0, double* gstats = nullptr) {
    __shared__ RingStage ring[kStages];
    __shared__ float4 ostage[8][64];
    __shared__ double sstat[16];
    if (epilogue) { if (threadIdx.x < 16) sstat[threadIdx.x] = 0.0; __syncthreads(); }
    float tsum = 0.f; int tcnt = 0;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t tile0 = blockIdx.x * chunk, tile_end = min(n_tiles_total, tile0 + chunk);
    auto prefetch = [&](uint32_t tile, uint32_t s) {
        const uint32_t j = tile * 256u + tid;
        cpa<16>(&ring[s].A[tid], b.A + j); cpa<16>(&ring[s].B[tid], b.B + j);
        cpa<4>(&ring[s].fuel[tid], b.fuel + j); cpa<4>(&ring[s].dry[tid], b.dry + j);
        cpa<8>(&ring[s].act[tid], b.act + j);
    };
#pragma unroll
    for (int d = 0; d + 1 < kStages; ++d) {
        if (tile0 + d < tile_end) prefetch(tile0 + d, d);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    uint32_t s = 0;
    for (uint32_t tile = tile0; tile < tile_end; ++tile) {
        const uint32_t i = tile * 256u + tid;
        asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");
        const float4 a = ring[s].A[tid], bb = ring[s].B[tid];
        const float fu = ring[s].fuel[tid], d = ring[s].dry[tid];
        const float2 ac = ring[s].act[tid];
        const uint32_t next = tile + kStages - 1;
        const uint32_t s_next = (kStages == 2) ? (s ^ 1u) : ((s == 0u) ? (uint32_t)(kStages - 1) : s - 1u);
        if (next < tile_end) prefetch(next, s_next);
        asm volatile("cp.async.commit_group;" ::: "memory");
        s = (s + 1u == (uint32_t)kStages) ? 0u : s + 1u;
        float sacc = a.x + bb.y + fu * d + ac.x * ac.y;
        for (int k = 0; k < work; ++k) sacc = fmaf(sacc, 0.999f, 0.001f);      // dependent chain, `work` issue slots
        if (div_pct) {                                                         // warp-uniform, pseudo-random per warp-tile
            const uint32_t h = ((tile * 8u + (tid >> 5)) * 2654435761u) >> 16;
            if ((int)(h % 100u) < div_pct)
                for (int k = 0; k < div_work; ++k) sacc = fmaf(sacc, 1.001f, -0.001f);
        }
        tsum += sacc; tcnt += 1;
        const float4 nA = make_float4(a.y, a.x, a.w + sacc, a.z), nB = make_float4(bb.y, bb.x, bb.w, bb.z + sacc);
        b.A[i] = nA; b.B[i] = nB; b.fuel[i] = fu + 1.0f;
        float4* st = ostage[tid >> 5];
        const uint32_t w = (lane >> 2) & 1u;
        st[(2u * lane) ^ w] = make_float4(sacc, a.x, a.y, a.z);
        st[(2u * lane + 1u) ^ w] = make_float4(bb.x, bb.y, bb.z, sacc);
        __syncwarp();
        const uint32_t c = lane ^ ((lane >> 3) & 1u);
        const float4 v1 = st[c], v2 = st[32u + c];
        __syncwarp();
        const uint32_t row0 = i - lane;
        b.obs[2 * (size_t)row0 + lane] = v1; b.obs[2 * (size_t)row0 + 32u + lane] = v2;
        b.rew[i] = sacc;
        b.f0[i] = sacc > 0.0f; b.f1[i] = sacc < 1.0f; b.f2[i] = (uint8_t)(sacc == 2.0f);
    }
    if (epilogue) {
        double d0 = (double)tsum, d1 = (double)tsum * 0.5; int c0 = tcnt, c1 = tcnt, c2 = tcnt, c3 = tcnt, c4 = tcnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o);
            c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
            c4 += __shfl_xor_sync(0xffffffffu, c4, o);
        }
        if (lane == 0) {
            atomicAdd(&sstat[0], d0); atomicAdd(&sstat[1], d1); atomicAdd(&sstat[2], (double)c0); atomicAdd(&sstat[3], (double)c1);
            atomicAdd(&sstat[4], (double)c2); atomicAdd(&sstat[5], (double)c3); atomicAdd(&sstat[6], (double)c4);
        }
        __syncthreads();
        if (tid == 0) for (int k = 7; k < 13; ++k) sstat[k] = 1.0;
        __syncthreads();
        if (tid < 13 && sstat[tid] != 0.0) atomicAdd(&gstats[(blockIdx.x % 64) * 16 + tid], sstat[tid]);
    }
}

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atoll(argv[1]) : 16777216u;     // multiple of 512
    Bufs b; float2* act;
    CK(cudaMalloc(&b.A, 16ull * n)); CK(cudaMalloc(&b.B, 16ull * n)); CK(cudaMalloc(&b.obs, 32ull * n));
    CK(cudaMalloc(&b.fuel, 4ull * n)); CK(cudaMalloc(&b.dry, 4ull * n)); CK(cudaMalloc(&b.rew, 4ull * n));
    CK(cudaMalloc(&act, 8ull * n)); CK(cudaMalloc(&b.f0, n)); CK(cudaMalloc(&b.f1, n)); CK(cudaMalloc(&b.f2, n));
    CK(cudaMalloc(&b.tiled, 40ull * n));
    b.act = act;
    CK(cudaMemset(b.A, 0, 16ull * n)); CK(cudaMemset(b.B, 0, 16ull * n)); CK(cudaMemset(b.fuel, 0, 4ull * n));
    CK(cudaMemset(b.dry, 0, 4ull * n)); CK(cudaMemset(act, 0, 8ull * n)); CK(cudaMemset(b.tiled, 0, 40ull * n));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto time_it = [&](auto launch, const char* name, double bytes) {
        for (int w = 0; w < 5; ++w) launch();
        float best = 1e30f, tot = 0;
        for (int r = 0; r < 20; ++r) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; tot += ms;
        }
        printf("%-44s best %.3f ms %5.0f GB/s   mean %.3f ms %5.0f GB/s\n", name, best, bytes / best / 1e6, tot / 20, bytes / (tot / 20) / 1e6);
        fflush(stdout);
    };
    // calibration on a 2 n float4 array pair (obs <- A|B region is too small; use tiled (40 n B) and obs (32 n B))
    const size_t nc = 2ull * n;   // float4 elements (32 n bytes)
    const float4* src = reinterpret_cast<const float4*>(b.tiled);
    time_it([&] { read_kernel<<<sms * 8, 256>>>(src, b.rew, nc); }, "read-only float4, persistent 8/SM", 16.0 * nc);
    time_it([&] { write_kernel<<<sms * 8, 256>>>(b.obs, nc); }, "write-only float4, persistent 8/SM", 16.0 * nc);
    time_it([&] { copy_kernel<<<sms * 8, 256>>>(src, b.obs, nc); }, "copy float4, persistent 8/SM", 32.0 * nc);
    time_it([&] { copy4_kernel<<<(unsigned)((nc + 1023) / 1024), 256>>>(src, b.obs, nc); }, "copy float4 x4, one tile per CTA", 32.0 * nc);
    time_it([&] { cudaMemcpyAsync(b.obs, src, 16 * nc, cudaMemcpyDeviceToDevice); }, "cudaMemcpyAsync D2D", 32.0 * nc);

    const unsigned tiles = n / 256;
    if (argc > 2 && argv[2][0] == 'r') {
        // ring mode: the step kernel's skeleton with `work` dependent FMAs per rocket
        for (int work : {0, 100, 200, 300, 400, 600})
            for (unsigned chunk : {8u}) {
                char name[64];
                snprintf(name, sizeof name, "ring s2 chunk %u work %d", chunk, work);
                time_it([&] { mix_ring<2><<<(tiles + chunk - 1) / chunk, 256>>>(b, tiles, chunk, work); }, name, 123.0 * n);
                snprintf(name, sizeof name, "ring s3 chunk %u work %d", chunk, work);
                time_it([&] { mix_ring<3><<<(tiles + chunk - 1) / chunk, 256>>>(b, tiles, chunk, work); }, name, 123.0 * n);
            }
        for (unsigned chunk : {1u, 2u, 4u, 16u, 32u}) {
            char name[64];
            snprintf(name, sizeof name, "ring s2 chunk %u work 0", chunk);
            time_it([&] { mix_ring<2><<<(tiles + chunk - 1) / chunk, 256>>>(b, tiles, chunk, 0); }, name, 123.0 * n);
        }
        // what separates the skeleton from the step kernel: the divergent reset path and the CTA epilogue
        double* gstats; CK(cudaMalloc(&gstats, 64 * 16 * sizeof(double))); CK(cudaMemset(gstats, 0, 64 * 16 * sizeof(double)));
        const unsigned g8 = (tiles + 7) / 8;
        time_it([&] { mix_ring<2><<<g8, 256>>>(b, tiles, 8, 330); }, "ring chunk 8 work 330", 123.0 * n);
        time_it([&] { mix_ring<2><<<g8, 256>>>(b, tiles, 8, 330, 0, 0, 1, gstats); }, "ring chunk 8 work 330 + epilogue", 123.0 * n);
        time_it([&] { mix_ring<2><<<g8, 256>>>(b, tiles, 8, 330, 25, 400); }, "ring chunk 8 work 330 + 25% x 400", 123.0 * n);
        time_it([&] { mix_ring<2><<<g8, 256>>>(b, tiles, 8, 330, 25, 400, 1, gstats); }, "ring chunk 8 work 330 + 25% x 400 + epilogue", 123.0 * n);
        time_it([&] { mix_ring<2><<<g8, 256>>>(b, tiles, 8, 430, 0, 0, 1, gstats); }, "ring chunk 8 work 430 (same total, uniform) + epilogue", 123.0 * n);
        time_it([&] { mix_ring<2><<<g8, 256>>>(b, tiles, 8, 330, 25, 800, 1, gstats); }, "ring chunk 8 work 330 + 25% x 800 + epilogue", 123.0 * n);
        return 0;
    }
    if (argc > 2) {
        // sustained mode: the same kernel back to back for ~4 s, GB/s per batch of launches -- does the
        // attainable bandwidth hold once the board sits at its power cap (as it does inside bench.py)?
        auto sustain = [&](auto launch, const char* name, double bytes, int per_batch) {
            for (int b = 0; b < 16; ++b) {
                cudaEventRecord(e0);
                for (int k = 0; k < per_batch; ++k) launch();
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                printf("%-40s batch %2d: %6.0f GB/s (%.3f ms per launch)\n", name, b, bytes * per_batch / ms / 1e6, ms / per_batch);
                fflush(stdout);
            }
        };
        const int pb = (int)(250.0 / (123.0 * n / 6.0e9) + 1);      // ~0.25 s per batch
        sustain([&] { copy4_kernel<<<(unsigned)((nc + 1023) / 1024), 256>>>(src, b.obs, nc); }, "copy float4 x4, one tile per CTA", 32.0 * nc, pb * 2);
        sustain([&] { mix_kernel<1, true, false, false, 0, 0, true, true><<<tiles, 256>>>(b, n); }, "mix T N (one tile per CTA)", 123.0 * n, pb);
        sustain([&] { mix_kernel<1, true, false, true, 0, 0, true, true><<<sms * 4, 256>>>(b, n); }, "mix T 4/SM (persistent)", 123.0 * n, pb);
        return 0;
    }
#define MIX(name, bytes, grid, ...) time_it([&] { mix_kernel<__VA_ARGS__><<<(grid), 256>>>(b, n); }, name, (bytes) * (double)n)
    //                                       unroll obsT tiled persist ld st flags reward
    MIX("mix base 4/SM", 123.0, sms * 4,          1, false, false, true, 0, 0, true, true);
    MIX("mix T 4/SM", 123.0, sms * 4,             1, true, false, true, 0, 0, true, true);
    MIX("mix T S 4/SM", 123.0, sms * 4,           1, true, true, true, 0, 0, true, true);
    MIX("mix S 4/SM", 123.0, sms * 4,             1, false, true, true, 0, 0, true, true);
    MIX("mix T N (one tile per CTA)", 123.0, tiles, 1, true, false, false, 0, 0, true, true);
    MIX("mix T S N", 123.0, tiles,                1, true, true, false, 0, 0, true, true);
    MIX("mix T N unroll2", 123.0, tiles / 2,      2, true, false, false, 0, 0, true, true);
    MIX("mix T ld.cs 4/SM", 123.0, sms * 4,       1, true, false, true, 1, 0, true, true);
    MIX("mix T ld.cg 4/SM", 123.0, sms * 4,       1, true, false, true, 2, 0, true, true);
    MIX("mix T st.cs(outputs) 4/SM", 123.0, sms * 4, 1, true, false, true, 0, 1, true, true);
    MIX("mix T ld.cs st.cs 4/SM", 123.0, sms * 4, 1, true, false, true, 1, 1, true, true);
    MIX("mix T st.wt(outputs) 4/SM", 123.0, sms * 4, 1, true, false, true, 0, 3, true, true);
    MIX("mix T no flags 4/SM (120 B)", 120.0, sms * 4, 1, true, false, true, 0, 0, false, true);
    MIX("mix T no flags no reward 4/SM (116 B)", 116.0, sms * 4, 1, true, false, true, 0, 0, false, false);
    MIX("mix T 8/SM", 123.0, sms * 8,             1, true, false, true, 0, 0, true, true);
    MIX("mix T unroll2 4/SM", 123.0, sms * 4,     2, true, false, true, 0, 0, true, true);
    MIX("mix T unroll2 8/SM", 123.0, sms * 8,     2, true, false, true, 0, 0, true, true);
    MIX("mix T S unroll2 8/SM", 123.0, sms * 8,   2, true, true, true, 0, 0, true, true);
    MIX("mix base unroll2 8/SM", 123.0, sms * 8,  2, false, false, true, 0, 0, true, true);
    return 0;
}
