// Fused inference of the reference's actor-critic MLPs for device-resident rollout collection
// (BASELINE.json configs[4]).  The reference trains
// PPO("MlpPolicy", policy_kwargs=dict(net_arch=dict(pi=[256, 256], vf=[256, 256])))
// (scripts/train.py:117-134): two separate tanh MLPs 8 -> 256 -> 256 with a linear head (2 action
// means / 1 value), 136 965 parameters as in assets/model/v3.  Run layer by layer (PyTorch / cuBLAS)
// the 256-wide activations of a million rockets go to HBM and back after every layer and the
// collection loop spends 3.3 of its 3.6 ms per step there.  Here one launch evaluates one whole net:
// a 16-rocket row block stays in the registers of one warp from the observation to the head.
//
//   layer 1  (8 -> 256)    tensor cores, K = 16 = [bf16 high part | bf16 low part] of the fp32
//                          observation against W1 stacked twice: the input keeps ~16 mantissa bits
//   tanh                   MUFU tanh.approx on the fp32 accumulators
//   layer 2  (256 -> 256)  tensor cores (mma.sync m16n8k16, bf16 x bf16 -> fp32).  The accumulator
//                          fragments of layer 1 ARE the A fragments of layer 2 (same lane / register
//                          map once packed to bf16x2), so the activations never leave the register
//                          file; W2 (bf16, 128 KiB) sits in shared memory for the life of the CTA, its
//                          rows padded to 528 B so that ldmatrix is bank-conflict free
//   tanh, head             fp32 on the CUDA cores from the accumulator fragments, quad shuffle reduce
//
// Weights are bf16 (what torch.autocast(bfloat16) would use), accumulation and biases fp32; the
// test states the tolerance against the fp32 PyTorch module.  mma.sync, not tcgen05: the shape
// (K = 256, activations produced and consumed in registers) suits the warp-level instruction.
// Measured (profiles/mlp_forward_r1_summary.txt, 1 Mi rockets): 0.37 ms per net = 395 TFLOP/s on
// the tensor pipe, which is 70 % active (the warp-level HMMA path peaks near 565 TFLOP/s on this
// part, a quarter of tcgen05), shared-memory wavefronts at 84 % of peak.  A variant in which each
// warp keeps two row blocks and shares every W2 fragment between them (half the shared-memory
// traffic, 255 registers) measured the same 0.38-0.40 ms, and so did 12 and 16 warps per CTA with
// layer 2 accumulated in quarters / eighths (168 / 128 registers: 0.36 ms), so it is the warp-level
// tensor path itself that bounds it; a tcgen05 / TMEM version (M = 128 row blocks,
// activations staged in swizzled shared memory) would be MUFU-bound on the 2 x 256 tanh per rocket
// at roughly half the time -- the next step if this neighbour of the hot path becomes the
// bottleneck again.
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/rocket_landing_b200.h"

namespace {

thread_local char g_po_err[256] = "";
int po_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_po_err, sizeof(g_po_err), fmt, ap);
    va_end(ap);
    return code;
}

constexpr int kHidden = RL_MLP_HIDDEN;     // 256
constexpr int kObs = 8;
constexpr int kThreads = 256;              // 8 warps x 16 rows = 128 rows per CTA iteration
constexpr int kRowsPerWarp = 16;
constexpr int kRowsPerCta = (kThreads / 32) * kRowsPerWarp;

// packed weights in global memory (RL_MLP_BLOB_BYTES, built by the host side: rollout.py pack_mlp)
constexpr int kOffW1 = 0;                               // bf16 [256][16]: W1[n][0..7] twice
constexpr int kOffB1 = kOffW1 + kHidden * 16 * 2;       // f32 [256]
constexpr int kOffW2 = kOffB1 + kHidden * 4;            // bf16 [256][256]  (row = output unit, as nn.Linear.weight)
constexpr int kOffB2 = kOffW2 + kHidden * kHidden * 2;  // f32 [256]
constexpr int kOffW3 = kOffB2 + kHidden * 4;            // f32 [2][256]     (second row unused by a 1-wide head)
constexpr int kOffB3 = kOffW3 + 2 * kHidden * 4;        // f32 [2]
static_assert(kOffB3 + 16 == RL_MLP_BLOB_BYTES, "blob layout");

// shared memory: padded copies
constexpr int kW2Stride = kHidden * 2 + 16;             // 528 B: 8 consecutive rows hit 8 different 16-byte bank groups
constexpr int kW1Stride = 48;                           // 16 bf16 + 16 B pad, same reason
constexpr int kSmW2 = 0;
constexpr int kSmW1 = kSmW2 + kHidden * kW2Stride;
constexpr int kSmB1 = kSmW1 + kHidden * kW1Stride;
constexpr int kSmB2 = kSmB1 + kHidden * 4;
constexpr int kSmW3 = kSmB2 + kHidden * 4;
constexpr int kSmB3 = kSmW3 + 2 * kHidden * 4;
constexpr int kSmemBytes = kSmB3 + 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

// D (16x8, fp32) += A (16x16, bf16, row) x B (16x8, bf16, col)
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {       // lo -> bits 0..15 (the lower column index)
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// bf16 high part of x and bf16 of the remainder: x ~ hi + lo to ~16 mantissa bits
__device__ __forceinline__ void split_bf16(float x, float& hi, float& lo) {
    hi = __bfloat162float(__float2bfloat16_rn(x));
    lo = x - hi;
}

template <int kOut>
__global__ void __launch_bounds__(kThreads, 1) mlp_forward_kernel(const unsigned char* __restrict__ blob,
                                                                   const float* __restrict__ obs, float* __restrict__ out,
                                                                   int64_t n) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- weights -> shared memory, once per CTA (persistent grid) --------------------------------
    for (int c = tid; c < kHidden * 32; c += kThreads) {          // W2: 256 rows x 32 chunks of 16 B
        const int row = c >> 5, ch = c & 31;
        *reinterpret_cast<uint4*>(smem + kSmW2 + row * kW2Stride + ch * 16) =
            *reinterpret_cast<const uint4*>(blob + kOffW2 + row * (kHidden * 2) + ch * 16);
    }
    for (int c = tid; c < kHidden * 2; c += kThreads) {           // W1 (stacked twice): 256 rows x 2 chunks
        const int row = c >> 1, ch = c & 1;
        *reinterpret_cast<uint4*>(smem + kSmW1 + row * kW1Stride + ch * 16) =
            *reinterpret_cast<const uint4*>(blob + kOffW1 + row * 32 + ch * 16);
    }
    for (int c = tid; c < kHidden * 4 / 16; c += kThreads)        // b1
        *reinterpret_cast<uint4*>(smem + kSmB1 + c * 16) = *reinterpret_cast<const uint4*>(blob + kOffB1 + c * 16);
    for (int c = tid; c < (kSmemBytes - kSmB2) / 16; c += kThreads)   // b2 | W3 | b3: contiguous in the blob and here
        *reinterpret_cast<uint4*>(smem + kSmB2 + c * 16) = *reinterpret_cast<const uint4*>(blob + kOffB2 + c * 16);
    __syncthreads();

    const float* b1s = reinterpret_cast<const float*>(smem + kSmB1);
    const float* b2s = reinterpret_cast<const float*>(smem + kSmB2);
    const float* w3s = reinterpret_cast<const float*>(smem + kSmW3);
    const float* b3s = reinterpret_cast<const float*>(smem + kSmB3);

    // fragment coordinates of this lane (PTX ISA, mma.m16n8k16): rows g and g + 8, column pair 2 q
    const int g = lane >> 2, q = lane & 3;
    // ldmatrix.x4 row address of this lane: matrix lane / 8 = (n-tile pair member, k half), row lane % 8
    const int lm_mat = lane >> 3, lm_row = lane & 7;
    const uint32_t w2_lane = smem_u32(smem + kSmW2) + (uint32_t)((8 * (lm_mat >> 1) + lm_row) * kW2Stride + 16 * (lm_mat & 1));
    const uint32_t w1_lane = smem_u32(smem + kSmW1) + (uint32_t)((8 * (lm_mat >> 1) + lm_row) * kW1Stride + 16 * (lm_mat & 1));

    const int64_t n_blocks = (n + kRowsPerCta - 1) / kRowsPerCta;
    for (int64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const int64_t row0 = blk * kRowsPerCta + warp * kRowsPerWarp;       // this warp's 16 rows
        if (row0 >= n) continue;
        const int64_t ra = row0 + g, rb = row0 + g + 8;

        // ---- layer 1: A = [obs_hi | obs_lo] (16 x 16), B = W1 stacked twice ------------------------
        uint32_t a1[4];
        {
            float2 xa = make_float2(0.f, 0.f), xb = make_float2(0.f, 0.f);
            if (ra < n) xa = *reinterpret_cast<const float2*>(obs + ra * kObs + 2 * q);
            if (rb < n) xb = *reinterpret_cast<const float2*>(obs + rb * kObs + 2 * q);
            float h0, l0, h1, l1, h2, l2, h3, l3;
            split_bf16(xa.x, h0, l0); split_bf16(xa.y, h1, l1);
            split_bf16(xb.x, h2, l2); split_bf16(xb.y, h3, l3);
            a1[0] = pack_bf16(h0, h1);      // row g,     k = 2q, 2q+1        (high parts)
            a1[1] = pack_bf16(h2, h3);      // row g + 8
            a1[2] = pack_bf16(l0, l1);      // row g,     k = 8 + 2q, 9 + 2q  (low parts)
            a1[3] = pack_bf16(l2, l3);      // row g + 8
        }
        uint32_t a2[kHidden / 16][4];       // layer-2 A fragments: 16 k-tiles of the 16 x 256 activations
#pragma unroll
        for (int np = 0; np < kHidden / 16; ++np) {                 // pairs of n-tiles 2 np, 2 np + 1 = k-tile np of layer 2
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(w1_lane + (uint32_t)(np * 16 * kW1Stride), b0, b1, b2, b3);
            float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
            mma_bf16(d0, a1, b0, b1);
            mma_bf16(d1, a1, b2, b3);
            const float2 ba = *reinterpret_cast<const float2*>(b1s + 16 * np + 2 * q);
            const float2 bb = *reinterpret_cast<const float2*>(b1s + 16 * np + 8 + 2 * q);
            a2[np][0] = pack_bf16(tanh_fast(d0[0] + ba.x), tanh_fast(d0[1] + ba.y));     // row g,     k = 16 np + 2q ..
            a2[np][1] = pack_bf16(tanh_fast(d0[2] + ba.x), tanh_fast(d0[3] + ba.y));     // row g + 8
            a2[np][2] = pack_bf16(tanh_fast(d1[0] + bb.x), tanh_fast(d1[1] + bb.y));     // row g,     k = 16 np + 8 + 2q ..
            a2[np][3] = pack_bf16(tanh_fast(d1[2] + bb.x), tanh_fast(d1[3] + bb.y));     // row g + 8
        }

        // ---- layer 2 in two halves of 128 output units, head accumulated on the fly -------------------
        float sa[kOut], sb[kOut];           // head partial sums of rows g and g + 8
#pragma unroll
        for (int o = 0; o < kOut; ++o) sa[o] = sb[o] = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float acc[16][4];
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
            for (int kt = 0; kt < kHidden / 16; ++kt) {
#pragma unroll
                for (int jp = 0; jp < 8; ++jp) {                    // n-tiles 16 half + 2 jp, + 1
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4(w2_lane + (uint32_t)((128 * half + 16 * jp) * kW2Stride + 32 * kt), b0, b1, b2, b3);
                    mma_bf16(acc[2 * jp], a2[kt], b0, b1);
                    mma_bf16(acc[2 * jp + 1], a2[kt], b2, b3);
                }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int col = 128 * half + 8 * j + 2 * q;
                const float2 bias = *reinterpret_cast<const float2*>(b2s + col);
                const float va0 = tanh_fast(acc[j][0] + bias.x), va1 = tanh_fast(acc[j][1] + bias.y);
                const float vb0 = tanh_fast(acc[j][2] + bias.x), vb1 = tanh_fast(acc[j][3] + bias.y);
#pragma unroll
                for (int o = 0; o < kOut; ++o) {
                    const float2 w = *reinterpret_cast<const float2*>(w3s + o * kHidden + col);
                    sa[o] = fmaf(va0, w.x, fmaf(va1, w.y, sa[o]));
                    sb[o] = fmaf(vb0, w.x, fmaf(vb1, w.y, sb[o]));
                }
            }
        }
#pragma unroll
        for (int o = 0; o < kOut; ++o) {                            // the four lanes of a quad hold one row's partial sums
            sa[o] += __shfl_xor_sync(0xffffffffu, sa[o], 1);
            sa[o] += __shfl_xor_sync(0xffffffffu, sa[o], 2);
            sb[o] += __shfl_xor_sync(0xffffffffu, sb[o], 1);
            sb[o] += __shfl_xor_sync(0xffffffffu, sb[o], 2);
        }
        if (q == 0) {
#pragma unroll
            for (int o = 0; o < kOut; ++o) {
                if (ra < n) out[ra * kOut + o] = sa[o] + b3s[o];
                if (rb < n) out[rb * kOut + o] = sb[o] + b3s[o];
            }
        }
    }
}

}  // namespace

extern "C" {

const char* rl_policy_last_error(void) { return g_po_err; }

int64_t rl_mlp_blob_bytes(void) { return RL_MLP_BLOB_BYTES; }

// out[n][out_dim] = head(tanh(W2 tanh(W1 obs + b1) + b2)) for the packed net `blob` (DEVICE memory,
// RL_MLP_BLOB_BYTES, 16-byte aligned); obs float32 [n][8] (8-byte aligned), out float32.  Enqueue only.
int rl_mlp_forward(const void* blob, const float* obs, float* out, int64_t n, int out_dim, void* stream) {
    if (!blob || !obs || !out) return po_fail(RL_E_INVALID, "rl_mlp_forward: null argument");
    if (n < 1) return po_fail(RL_E_INVALID, "rl_mlp_forward: n must be positive");
    if (out_dim != 1 && out_dim != 2) return po_fail(RL_E_INVALID, "rl_mlp_forward: out_dim must be 1 or 2 (got %d)", out_dim);
    if ((reinterpret_cast<uintptr_t>(blob) & 15) || (reinterpret_cast<uintptr_t>(obs) & 7))
        return po_fail(RL_E_INVALID, "rl_mlp_forward: blob needs 16-byte, obs 8-byte alignment");
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t blocks = (n + kRowsPerCta - 1) / kRowsPerCta;
    const int grid = (int)(blocks < sms ? blocks : sms);
    cudaError_t e;
    if (out_dim == 2) {
        e = cudaFuncSetAttribute(mlp_forward_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess)
            mlp_forward_kernel<2><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>((const unsigned char*)blob, obs, out, n);
    } else {
        e = cudaFuncSetAttribute(mlp_forward_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess)
            mlp_forward_kernel<1><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>((const unsigned char*)blob, obs, out, n);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return po_fail(RL_E_CUDA, "rl_mlp_forward launch failed: %s", cudaGetErrorString(e));
    return RL_OK;
}

}  // extern "C"
