// Issue rates of the instructions the fused MLP kernel's epilogues are made of, per SM sub-partition:
// scalar FFMA, packed FFMA2 / FMUL2, MUFU.TANH (f32 and bf16), FMNMX, HMNMX2.BF16, and MUFU mixed with FFMA2.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 2048, kChains = 8;

template <int kMode>
__global__ void __launch_bounds__(512) rate_kernel(float* out, long long* cycles, float seed) {
    __shared__ __align__(16) float sh[256 + 4 * 512];
    for (int i = threadIdx.x; i < 256 + 4 * 512; i += blockDim.x) sh[i] = 0.f;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sh);
    float a[kChains * 2];
#pragma unroll
    for (int i = 0; i < kChains * 2; ++i) a[i] = seed + (float)(threadIdx.x + i) * 1e-3f;
    const float m = 0.999f + seed * 1e-6f, c = seed * 1e-7f;
    uint64_t mm, cc;
    asm("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            float& x = a[2 * i];
            float& y = a[2 * i + 1];
            if (kMode == 0) {                       // 2 scalar FFMA
                x = fmaf(x, m, c);
                y = fmaf(y, m, c);
            } else if (kMode == 1 || kMode == 6) {  // 1 FFMA2
                uint64_t v;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x), "f"(y));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(mm), "l"(cc));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
                if (kMode == 6 && (i & 3) == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x));   // + 1 MUFU per 4 FFMA2
            } else if (kMode == 2) {                // 1 FMUL2
                uint64_t v;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x), "f"(y));
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(mm));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
            } else if (kMode == 3) {                // 2 MUFU.TANH
                asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x));
                asm volatile("tanh.approx.f32 %0, %0;" : "+f"(y));
            } else if (kMode == 4) {                // tanh.bf16x2 = 2 MUFU.TANH.BF16 + PRMT
                uint32_t u = __float_as_uint(x);
                asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u));
                x = __uint_as_float(u);
            } else if (kMode == 5) {                // 2 FMNMX
                asm volatile("min.f32 %0, %0, %1;" : "+f"(x) : "f"(m));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(y) : "f"(c));
            } else if (kMode >= 8 && kMode <= 11) { // 4 MUFU.TANH + one shared-memory instruction (do they share a port?)
                asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x));
                asm volatile("tanh.approx.f32 %0, %0;" : "+f"(y));
                if ((i & 1) == 0) {
                    if (kMode == 8) {               // LDS.128, all lanes the same address (a broadcast, as the head's weights)
                        float4 v;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sbase + 16u * (uint32_t)(it & 63)));
                        x += v.x * 1e-30f;
                    } else if (kMode == 9) {        // STS.128, one 16-byte row per lane (512 contiguous bytes, as the A2s writes)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" :: "r"(sbase + 1024u + 16u * threadIdx.x), "f"(x) : "memory");
                    } else if (kMode == 10) {       // LDS.32 broadcast
                        float v;
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sbase + 16u * (uint32_t)(it & 63)));
                        x += v * 1e-30f;
                    }                               // 11: nothing (4 MUFU alone)
                }
            } else if (kMode == 7) {                // 2 HMNMX2.BF16
                uint32_t u = __float_as_uint(x), w = __float_as_uint(m);
                asm volatile("min.bf16x2 %0, %0, %1;" : "+r"(u) : "r"(w));
                asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(u) : "r"(__float_as_uint(c)));
                x = __uint_as_float(u);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kChains * 2; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Tensor-memory read rate: `warps` warps of one CTA (warp w reads lanes 32 (w % 4) ..) issue tcgen05.ld 32x32b.x32
// (32 fp32 columns of 32 lanes = 4 KiB per warp-instruction) back to back, one wait per `batch` loads.
__global__ void __launch_bounds__(512) ldtm_kernel(float* out, long long* cycles, int batch) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t r[32];
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < 1024; ++it) {
        const uint32_t col = (uint32_t)((it * 32 + warp * 64) & 511) & ~31u;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(base + col) : "memory");
        if ((it + 1) % batch == 0) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(r[it & 31] & 0x007fffffu);
        }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(slot) : "memory");
}

void run_ldtm(int warps, int batch, float* out, long long* cyc) {
    for (int rep = 0; rep < 2; ++rep) {
        ldtm_kernel<<<148, 32 * warps>>>(out, cyc, batch);
        cudaDeviceSynchronize();
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    printf("tcgen05.ld 32x32b.x32, %2d warps, wait every %d: %8.0f cycles  %6.1f bytes per cycle and SM  (%s)\n", warps, batch, avg,
           (double)warps * 1024 * 4096 / avg, cudaGetErrorString(cudaGetLastError()));
}

template <int kMode>
void run(const char* name, int instr_per_chain, float* out, long long* cyc) {
    rate_kernel<kMode><<<148, 512>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    rate_kernel<kMode><<<148, 512>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    // 512 threads = 16 warps = 4 warps per sub-partition
    const double warp_instr_per_smsp = 4.0 * kIters * kChains * instr_per_chain;
    printf("%-44s %8.0f cycles  %6.2f cycles per warp-instruction per sub-partition\n", name, avg, avg / warp_instr_per_smsp);
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 512 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    run<0>("FFMA (2 per pair)", 2, out, cyc);
    run<1>("FFMA2 (1 per pair)", 1, out, cyc);
    run<2>("FMUL2 (1 per pair)", 1, out, cyc);
    run<3>("MUFU.TANH f32 (2 per pair)", 2, out, cyc);
    run<4>("tanh.bf16x2 (2 MUFU + PRMT: per MUFU)", 2, out, cyc);
    run<5>("FMNMX (2 per pair)", 2, out, cyc);
    run<7>("HMNMX2.BF16 (2 per pair)", 2, out, cyc);
    run<6>("FFMA2 + 1 MUFU per 4 (per FFMA2)", 1, out, cyc);
    run<11>("MUFU.TANH alone (per MUFU)", 2, out, cyc);
    run<8>("MUFU.TANH + 1 LDS.128 broadcast per 4 (per MUFU)", 2, out, cyc);
    run<10>("MUFU.TANH + 1 LDS.32 broadcast per 4 (per MUFU)", 2, out, cyc);
    run<9>("MUFU.TANH + 1 STS.128 per 4 (per MUFU)", 2, out, cyc);
    for (int warps : {4, 8, 16})
        for (int batch : {1, 4}) run_ldtm(warps, batch, out, cyc);
    return 0;
}
