// Attainable-bandwidth probe for the step kernel's access pattern (no physics): per rocket it
// reads 48 B (two float4 planes, two float planes, one float2 action) and writes 75 B (two
// float4 planes, one float plane, 32 B observation, 4 B reward, three byte flags) -- the same
// nine streams, the same per-thread widths, the same persistent grid -- and, for reference, a
// plain float4 copy.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a stream_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(256) copy_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) out[i] = in[i];
}

template <int kUnroll>
__global__ void __launch_bounds__(256) mix_kernel(float4* A, float4* B, float* fuel, const float* dry, const float2* act,
                                                  float4* obs, float* rew, uint8_t* f0, uint8_t* f1, uint8_t* f2, uint32_t n) {
    const uint32_t stride = gridDim.x * 256u * kUnroll;
    for (uint32_t base = blockIdx.x * 256u * kUnroll + threadIdx.x; base < n; base += stride) {
        float4 a[kUnroll], b[kUnroll]; float fu[kUnroll], d[kUnroll]; float2 ac[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t i = base + u * 256u;
            if (i < n) { a[u] = A[i]; b[u] = B[i]; fu[u] = fuel[i]; d[u] = dry[i]; ac[u] = act[i]; }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const uint32_t i = base + u * 256u;
            if (i < n) {
                const float s = a[u].x + b[u].y + fu[u] * d[u] + ac[u].x * ac[u].y;
                A[i] = make_float4(a[u].y, a[u].x, a[u].w + s, a[u].z);
                B[i] = make_float4(b[u].y, b[u].x, b[u].w, b[u].z + s);
                fuel[i] = fu[u] + 1.0f;
                obs[2 * (size_t)i] = make_float4(s, a[u].x, a[u].y, a[u].z);
                obs[2 * (size_t)i + 1] = make_float4(b[u].x, b[u].y, b[u].z, s);
                rew[i] = s;
                f0[i] = s > 0.0f; f1[i] = s < 1.0f; f2[i] = (uint8_t)(s == 2.0f);
            }
        }
    }
}

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atoll(argv[1]) : 16777216u;
    float4 *A, *B, *obs; float *fuel, *dry, *rew; float2* act; uint8_t *f0, *f1, *f2;
    CK(cudaMalloc(&A, 16ull * n)); CK(cudaMalloc(&B, 16ull * n)); CK(cudaMalloc(&obs, 32ull * n));
    CK(cudaMalloc(&fuel, 4ull * n)); CK(cudaMalloc(&dry, 4ull * n)); CK(cudaMalloc(&rew, 4ull * n));
    CK(cudaMalloc(&act, 8ull * n)); CK(cudaMalloc(&f0, n)); CK(cudaMalloc(&f1, n)); CK(cudaMalloc(&f2, n));
    CK(cudaMemset(A, 0, 16ull * n)); CK(cudaMemset(B, 0, 16ull * n)); CK(cudaMemset(fuel, 0, 4ull * n));
    CK(cudaMemset(dry, 0, 4ull * n)); CK(cudaMemset(act, 0, 8ull * n));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto time_it = [&](auto launch, const char* name, double bytes) {
        for (int w = 0; w < 5; ++w) launch();
        float best = 1e30f, tot = 0;
        for (int r = 0; r < 20; ++r) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; tot += ms;
        }
        printf("%-34s best %.3f ms  %.0f GB/s   mean %.3f ms  %.0f GB/s\n", name, best, bytes / best / 1e6, tot / 20, bytes / (tot / 20) / 1e6);
    };
    const size_t nc = (size_t)n * 2;   // float4 elements: 32 B read + 32 B written per rocket-equivalent
    time_it([&] { copy_kernel<<<sms * 8, 256>>>((const float4*)obs, (float4*)obs + 0, 0); }, "empty launch", 1.0);
    time_it([&] { copy_kernel<<<sms * 8, 256>>>(A, obs, (size_t)n); }, "float4 copy (16B r + 16B w / elem)", 32.0 * n);
    for (int ctas : {4, 5, 6, 8}) {
        char name[64];
        snprintf(name, sizeof name, "mix 48r/75w unroll1 %d CTA/SM", ctas);
        time_it([&] { mix_kernel<1><<<sms * ctas, 256>>>(A, B, fuel, dry, act, obs, rew, f0, f1, f2, n); }, name, 123.0 * n);
        snprintf(name, sizeof name, "mix 48r/75w unroll2 %d CTA/SM", ctas);
        time_it([&] { mix_kernel<2><<<sms * ctas, 256>>>(A, B, fuel, dry, act, obs, rew, f0, f1, f2, n); }, name, 123.0 * n);
    }
    (void)nc;
    return 0;
}
