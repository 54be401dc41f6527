// Fused inference of the reference's actor-critic MLPs for device-resident rollout collection
// (BASELINE.json configs[4]).  The reference trains
// PPO("MlpPolicy", policy_kwargs=dict(net_arch=dict(pi=[256, 256], vf=[256, 256])))
// (scripts/train.py:117-134): two separate tanh MLPs 8 -> 256 -> 256 with a linear head (2 action
// means / 1 value), 136 965 parameters as in assets/model/v3.  Run layer by layer (PyTorch / cuBLAS)
// the 256-wide activations of a million rockets go to HBM and back after every layer and the
// collection loop spends 3.3 of its 3.6 ms per step there.  Here one launch evaluates one whole net
// and nothing but the observation (32 B) and the head output (4-8 B) of a rocket touches HBM.
//
// tcgen05 / TMEM: one persistent CTA per SM steps 128-rocket row blocks; both layers are
// single-thread-issued tcgen05.mma (M = 128, K = 16 per instruction, bf16 x bf16 -> fp32 in tensor
// memory).  What bounds the kernel once the GEMMs are on tcgen05 is the MUFU pipe: 2 x 256 tanh per
// rocket = 4096 MUFU cycles per row block and scheduler, twice what the tensor core needs.  The
// kernel is therefore built around keeping the MUFU pipe fed (see "schedule" below).
//
//   shared memory (operands in the canonical K-major no-swizzle layout: 8 x 16-byte "core matrices",
//   element (row r, k) at (k / 8) LBO + (r / 8) 128 + (r % 8) 16 + (k % 8) 2 bytes):
//     W2s  256 x 256 bf16, LBO 4096   (one TMA bulk copy per CTA)  W1s  256 x 16, LBO 4096
//     A2s  128 x 256 bf16, LBO 2048   (layer-1 activations)        A1s  2 x (128 x 16), LBO 2048 (obs hi | lo)
//     Bb1, Bb2, Ones                   bias tiles: b1 and b2 come out of the tensor core too
//   tensor memory (512 columns): [0, 128) layer-1 ring (2 slots of 64), [128, 512) three layer-2
//   accumulators of 128 columns used in rotation
//
//   layer 1 (8 -> 256)   K = 16 = [bf16 high part | bf16 low part] of the fp32 observation against W1
//                        stacked twice: the input keeps ~16 mantissa bits; four N = 64 slices
//   epilogue 1  (team X) ring slot -> tanh -> bf16 -> A2s, one K slice of 64 units at a time
//   layer 2 (256 -> 256) two N = 128 halves (A: units 0..127, B: 128..255), each 16 x tcgen05.mma
//                        issued slice by slice behind team X
//   epilogue 2  (team Y) accumulator -> tanh -> head dot product -> out (or: sample the action, see below)
//
// Schedule.  The first tcgen05 version moved all 16 warps in lockstep through epilogue 1 -> layer 2
// -> epilogue 2 of a row block: whenever they met at a CTA barrier, waited for a TMEM load or for
// the last MMAs to drain, the MUFU pipe had nothing to do (58 % busy, 0.22 ms per net at 1 Mi
// rockets).  Now the two epilogues belong to two teams of 8 warps that run about one row block apart
// and only ever meet through mbarriers, so that every scheduler holds warps of both kinds, and a
// 17th warp does nothing but issue the MMAs:
//   * tensor memory cannot hold two whole layer-2 accumulators next to layer 1, so half A of block i
//     goes to accumulator 2 i mod 3 (drained by team Y one block ago) and half B to 2 i + 1 mod 3
//     (which held half A of block i - 1: free about mid-way through epilogue 1 of block i) -- no MMA
//     waits for the END of an epilogue;
//   * layer 1 of the slice two ahead is issued into a ring slot as soon as team X has drained it;
//   * team X may start the next block at once: A2s slice 0 has been read by then (barrier e0),
//     slices 1..2 / 3 are released by the commits of half A / B;
//   * both bias adds are MMAs against a tile of ones (no LDS + FADD in either epilogue);
//   * the period of the pipeline cannot be shorter than the issuing warp's own instruction stream
//     (42 MMAs, 10 commits, 10 waits per block), so the observations are staged by team Y and the
//     descriptors are 32-bit words that only need an add.
//   * W2 (128 KiB) is stored in the blob as the image of its shared-memory operand and staged by one
//     bulk copy of the TMA engine that only the issuing warp ever waits for.
//   * one pair of units in four of epilogue 1 (and of the value net's epilogue 2) evaluates its tanh on the
//     FMA pipe instead ("tanh on the FMA pipe" below), and the MMAs are issued behind elect.sync, which
//     lets ptxas issue them from uniform registers without a broadcast loop per instruction.
// Measured (profiles/mlp_forward_r2c_summary.txt): 0.140 ms (value net) / 0.157-0.163 ms (policy net) at
// 1 Mi rockets, MUFU pipe 71 % busy, tensor pipe 55 % (all tanh on the MUFU pipe: 0.160 / 0.172 ms, 83 %).
//
// Weights are bf16 (what torch.autocast(bfloat16) would use), accumulation fp32, biases bf16 high +
// low parts (16 mantissa bits); measured against the fp32 module: rms error 8e-4, max 4e-3 on
// outputs of mean magnitude 0.24 (autocast: 1.2e-3 / 7e-3; tools/mlp_error_probe.py), and the test
// states the tolerance.  Earlier versions are in the history of this file: warp-level mma.sync with
// the activations chained through registers (0.37 ms), lockstep tcgen05 (0.22 ms).
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/rocket_landing_b200.h"
#include "rl_device.cuh"
#include "rl_launch.cuh"

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
// packed weights in global memory (RL_MLP_BLOB_BYTES, built by the host side: rollout.py pack_mlp)
constexpr int kOffW1 = 0;                               // bf16 [256][16]: W1[n][0..7] twice
constexpr int kOffB1 = kOffW1 + kHidden * 16 * 2;       // f32 [256]
constexpr int kOffW2 = kOffB1 + kHidden * 4;            // bf16 [32][256][8]: W2[n][k] at [k / 8][n][k % 8] = the image of W2s
constexpr int kOffB2 = kOffW2 + kHidden * kHidden * 2;  // f32 [256]
constexpr int kOffW3 = kOffB2 + kHidden * 4;            // f32 [2][256]     (second row unused by a 1-wide head)
constexpr int kOffB3 = kOffW3 + 2 * kHidden * 4;        // f32 [2]
static_assert(kOffB3 + 16 == RL_MLP_BLOB_BYTES, "blob layout");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

// Optional epilogue of the policy net (out_dim 2): draw the action the way SB3's DiagGaussian
// distribution does -- action = mean + exp(log_std) eps, eps ~ N(0, 1); log-probability summed over
// the two components; the Box-clipped copy that goes to env.step (SB3 clips before stepping and
// stores the unclipped sample) -- with counter-based noise: Philox4x32-10 keyed by `seed` xor a domain tag,
// counter (global row, draw), Box-Muller.  All null / zero: plain forward.
struct SampleArgs {
    const float* log_std;      // device [2]
    float* actions;            // [n][2] unclipped samples (null: no sampling)
    float* clipped;            // [n][2]
    float* log_prob;           // [n]
    float* mean;               // [n][2], nullable
    unsigned long long seed, draw, row_offset;
    float lo0, lo1, hi0, hi1;
};

namespace tc5 {

constexpr int kM = 128;                       // rows per block = TMEM lanes
constexpr int kLboW = 4096, kLboA = 2048, kSbo = 128;
#ifndef RL_MLP_Y_WARPS
#define RL_MLP_Y_WARPS 8
#endif
constexpr int kXWarps = 8, kYWarps = RL_MLP_Y_WARPS;      // team sizes: multiples of 4 (a warp reads TMEM lanes 32 (w % 4) .. + 31)
constexpr int kYSlices = kYWarps / 4;                     // team Y warps per lane quarter = column slices of an N half
constexpr int kYUnits = 128 / kYSlices;                   // units of an N half per team-Y thread
static_assert(kYUnits % 32 == 0, "team Y reads 32 columns at a time");
constexpr int kThreads = 32 * (kXWarps + kYWarps + 1);

constexpr int kSmW2 = 0;                                  // 131072
constexpr int kSmW1 = kSmW2 + kHidden * kHidden * 2;      // 8192
constexpr int kSmA2 = kSmW1 + kHidden * 16 * 2;           // 65536
constexpr int kSmA1 = kSmA2 + kM * kHidden * 2;           // 4096: observations of even blocks
constexpr int kSmA1b = kSmA1 + kM * 16 * 2;               // ... of odd blocks
// Both biases come out of the tensor core as well: D = Ones (128 x 16: columns 0, 1 = 1) Bb^T + ..., with
// Bb (256 x 16) = [high part of b | remainder | 0 ...].  Only the first 8 k's of a bias tile are stored: its
// second K chunk (LBO = 4096 further on) meets the zero columns of Ones, so it may alias anything FINITE --
// the next bias tile, and for the last one the first chunk of Ones.
constexpr int kSmBb1 = kSmA1b + kM * 16 * 2;              // 256 x 8 bf16
constexpr int kSmBb2 = kSmBb1 + kHidden * 8 * 2;
constexpr int kSmOnes = kSmBb2 + kHidden * 8 * 2;         // 128 x 16 bf16, LBO 2048
static_assert(kSmBb2 - kSmBb1 == kLboW && kSmOnes - kSmBb2 == kLboW, "aliased second K chunks");
constexpr int kSmW3 = kSmOnes + kM * 16 * 2;
constexpr int kSmB3 = kSmW3 + 2 * kHidden * 4;
constexpr int kRedBufs = kYSlices > 2 ? 1 : 2;            // one buffer (and a second barrier after it has been read) when two do not fit
constexpr int kSmRed = kSmB3 + 16;                        // kRedBufs x [kYSlices - 1][128 rows][2] f32: partial sums handed to the last slice's warp
constexpr int kSmBar = kSmRed + kRedBufs * (kYSlices - 1) * kM * 2 * 4;
constexpr int kNumBars = 12;                              // x_ready[4] | l1bar[2] | barA | e0 | barB | y_doneA | y_doneB | wbar
constexpr int kSmemBytes = kSmBar + 8 * kNumBars + 8 + 16; // ... + the TMEM base address
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A / B bf16, both K-major, M = 128, N = n_cols
__host__ __device__ constexpr uint32_t idesc_n(int n_cols) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), no swizzle, version 1 (Blackwell): the low word
// holds (address >> 4) | (LBO >> 4) << 16, the high word SBO >> 4 and the version.  The issuing warp keeps one
// low word per operand array and adds (byte offset >> 4) per instruction.
constexpr uint32_t kDescHi = (uint32_t)(kSbo >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ void mma_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .b64 a, b;\n\t.reg .pred p;\n\tmov.b64 a, {%1, %3};\n\tmov.b64 b, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %4, p;\n\t}\n"
        :: "r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 consecutive fp32 columns of this thread's TMEM lane: issue, and (later) wait.  The wait names
// the registers as in/out operands so that no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}
// acc.x += a0 b0, acc.y += a1 b1 in one instruction (FFMA2)
__device__ __forceinline__ void ffma2(float2& acc, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}\n"
        : "+f"(acc.x), "+f"(acc.y) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// tanh of two bf16 values (two MUFU operations: there is no packed form in hardware)
__device__ __forceinline__ uint32_t tanh_bf16x2(uint32_t x) {
    uint32_t y;
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

// ---- tanh on the FMA pipe (part of both epilogues) ---------------------------------------------
// The MUFU pipe (16 tanh per clock and SM) is the kernel's roofline while two thirds of the issue slots and
// nearly all of the FMA pipe idle, so one pair in four of the layer-1 epilogue (RL_MLP_X_POLY) -- and of the
// layer-2 epilogue of the value net (RL_MLP_Y_POLY_VALUE; the policy net's team Y has no issue slots to spare:
// it carries a second head row) -- takes this route instead: tanh(x) ~ x R(x^2), fitted on |x| <= 3.8, with R of
// degree 7 in packed FFMA2 (two units per instruction, coefficients as immediates), clamped beyond; maximum
// absolute error 1.3e-3 over all x (tools/tanh_poly_fit.py).  The MUFU route of epilogue 1 rounds its INPUT
// to bf16 first and is no more accurate: the output error against the fp32 module goes DOWN (rms 8.3e-4 ->
// 7.8e-4).  Measured (profiles/tools/pipe_rates.cu): FFMA2 / FMUL2 issue every 2 cycles per scheduler, MUFU
// every 8, so a pair costs ~22 cycles of the FMA pipe here against 16 of the MUFU pipe there, and team X, which
// sets the pace, gets slower with every pair it converts: more than one pair in four loses again
// (profiles/README.md has the sweep).
#ifndef RL_MLP_X_POLY
#define RL_MLP_X_POLY 1
#endif
#ifndef RL_MLP_Y_POLY_VALUE
#define RL_MLP_Y_POLY_VALUE 1
#endif
#ifndef RL_MLP_Y_POLY_POLICY
#define RL_MLP_Y_POLY_POLICY 0
#endif
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// Inputs are clamped where the fitted polynomial has its maximum, x R(x^2) = 1.00016 at |x| = 3.727 (beyond the
// fitted range it turns over), so a saturated unit reads 1.0002 and the bound of 1.3e-3 holds for every input.
constexpr float kTanhClamp = 3.727f;
// x R(u) for the pair x with u = x^2 <= c^2 (the (c, c) pairs become FFMA2 immediates)
__device__ __forceinline__ uint64_t tanh_poly_core(uint64_t x, uint64_t u) {
    const float c[8] = {-9.573820403e-08f, 5.802626095e-06f, -1.463067778e-04f, 2.001198300e-03f,
                        -1.635971380e-02f, 8.460201472e-02f, -3.004610678e-01f, 9.930788304e-01f};
    uint64_t r = fma2(u, pk2(c[0], c[0]), pk2(c[1], c[1]));
#pragma unroll
    for (int i = 2; i < 8; ++i) r = fma2(r, u, pk2(c[i], c[i]));
    return mul2(x, r);
}
// epilogue 1: bf16x2 of tanh(x0) (bits 0..15), tanh(x1).  u is clamped, and beyond c the linear continuation
// is clamped at +-1 after the rounding to bf16 (two packed instructions for the pair).
__device__ __forceinline__ uint32_t tanh_poly_bf16x2(float x0, float x1) {
    const uint64_t x = pk2(x0, x1);
    float u0, u1, t0, t1;
    upk2(mul2(x, x), u0, u1);
    upk2(tanh_poly_core(x, pk2(fminf(u0, kTanhClamp * kTanhClamp), fminf(u1, kTanhClamp * kTanhClamp))), t0, t1);
    uint32_t b = pack_bf16(t0, t1);
    asm("min.bf16x2 %0, %0, %1;" : "+r"(b) : "r"(0x3F803F80u));
    asm("max.bf16x2 %0, %0, %1;" : "+r"(b) : "r"(0xBF80BF80u));
    return b;
}
// epilogue 2: fp32 results; the inputs are clamped to +-c instead (c R(c^2) = 1.00016: no clamp of the result)
__device__ __forceinline__ void tanh_poly_f32x2(float x0, float x1, float& t0, float& t1) {
    const uint64_t x = pk2(fminf(fmaxf(x0, -kTanhClamp), kTanhClamp), fminf(fmaxf(x1, -kTanhClamp), kTanhClamp));
    upk2(tanh_poly_core(x, mul2(x, x)), t0, t1);
}

// One thread of a converged warp (elect.sync): ptxas then knows that what the branch guards runs on a single
// thread and issues the tcgen05 instructions from uniform registers directly -- behind `lane == 0` every one of
// them sits in a loop that elects and broadcasts (ELECT / R2UR.BROADCAST / BRA.U.ANY, ~11 instructions per MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t is_leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(is_leader));
    return is_leader != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void quarter_barrier(int q) {          // the team-Y warps that share TMEM lanes 32 q .. 32 q + 31
    asm volatile("bar.sync %0, %1;" :: "r"(q + 1), "n"(32 * kYSlices) : "memory");
}

// One net over the row blocks cta, cta + n_ctas, ... (the whole CTA: 17 warps).  A launch runs it on every CTA
// (mlp_forward_tc5_kernel) or the policy net on some CTAs and the value net on the others (mlp_pair_kernel).
template <int kOut>
__device__ __forceinline__ void mlp_body(const unsigned char* __restrict__ blob, const float* __restrict__ obs,
                                         float* __restrict__ out, int64_t n, const SampleArgs& sa, int64_t cta, int64_t n_ctas) {
    extern __shared__ __align__(128) unsigned char smem[];
    rl::rl_grid_dependency_wait();                                // (programmatic dependent launch: rl_launch.cuh)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + kSmBar;
    const uint32_t x_ready = bars, l1bar = bars + 32, barA = bars + 48, e0 = bars + 56, barB = bars + 64, y_doneA = bars + 72, wbar = bars + 88,
                   y_doneB = bars + 80;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmBar + 8 * kNumBars + 8);

    // ---- one-time setup: weights in the canonical layout, barriers, tensor memory ------------------
    for (int c = tid; c < kHidden * 2; c += kThreads) {                  // W1 stacked twice: 2 chunks per row
        const int n_row = c & 255, kc = c >> 8;
        *reinterpret_cast<uint4*>(smem + kSmW1 + kc * kLboW + (n_row >> 3) * kSbo + (n_row & 7) * 16) =
            *reinterpret_cast<const uint4*>(blob + kOffW1 + n_row * 32 + kc * 16);
    }
    for (int c = tid; c < kM * 2; c += kThreads) {
        const int r = c & 127, kc = c >> 7;
        *reinterpret_cast<uint4*>(smem + kSmOnes + kc * kLboA + (r >> 3) * kSbo + (r & 7) * 16) =
            make_uint4(kc == 0 ? 0x3F803F80u : 0u, 0u, 0u, 0u);
    }
    for (int c = tid; c < kHidden * 2; c += kThreads) {                  // b1 | b2
        const int n_row = c & 255, layer = c >> 8;
        float hi, lo;
        split_bf16(*reinterpret_cast<const float*>(blob + (layer ? kOffB2 : kOffB1) + n_row * 4), hi, lo);
        *reinterpret_cast<uint4*>(smem + (layer ? kSmBb2 : kSmBb1) + (n_row >> 3) * kSbo + (n_row & 7) * 16) =
            make_uint4(pack_bf16(hi, lo), 0u, 0u, 0u);
    }
    for (int c = tid; c < (kSmRed - kSmW3) / 16; c += kThreads)        // W3 | b3
        *reinterpret_cast<uint4*>(smem + kSmW3 + c * 16) = *reinterpret_cast<const uint4*>(blob + kOffW3 + c * 16);
    if (tid == 0) {
        // W2 (128 KiB) is stored in the blob as the image of W2s: one bulk copy by the TMA engine, complete
        // (mbarrier wbar) long before the first layer-2 MMA needs it -- nothing else waits for it
        mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(wbar), "r"(kHidden * kHidden * 2) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(sbase + kSmW2), "l"(blob + kOffW2), "r"(kHidden * kHidden * 2), "r"(wbar) : "memory");
#pragma unroll
        for (int c = 0; c < 4; ++c) mbar_init(x_ready + 8 * c, kXWarps);
        mbar_init(l1bar, 1);
        mbar_init(l1bar + 8, 1);
        mbar_init(barA, 1);
        mbar_init(e0, 1);
        mbar_init(barB, 1);
        mbar_init(y_doneA, kYWarps);
        mbar_init(y_doneB, kYWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                              // one warp allocates all 512 columns (1 CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    const int64_t n_blocks = (n + kM - 1) / kM;
    const int64_t stride = n_ctas;
    const int q = warp & 3;                                       // TMEM lane quarter of this warp
    const int row = 32 * q + lane;                                // this thread's row of the block = its TMEM lane
    const int yh = (warp - kXWarps) >> 2;                         // team Y: which kYUnits of a half's 128 units

    // Observations: the first four warps of team Y hold one row each, two blocks ahead, and stage it as
    // A1 = [bf16 high parts | bf16 low parts] (two buffers, block i -> buffer i & 1).
    const bool stager = warp >= kXWarps && warp < kXWarps + 4;
    float4 oa = make_float4(0.f, 0.f, 0.f, 0.f), ob = oa;
    auto load_obs = [&](int64_t blk) {
        oa = ob = make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t r = blk * kM + row;
        if (stager && blk < n_blocks && r < n) {
            oa = *reinterpret_cast<const float4*>(obs + r * kObs);
            ob = *reinterpret_cast<const float4*>(obs + r * kObs + 4);
        }
    };
    auto stage_obs = [&](int buf) {
        if (stager) {
            const float x[8] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
            float hi[8], lo[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) split_bf16(x[k], hi[k], lo[k]);
            unsigned char* dst = smem + (buf ? kSmA1b : kSmA1) + (row >> 3) * kSbo + (row & 7) * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[3]), pack_bf16(hi[4], hi[5]), pack_bf16(hi[6], hi[7]));
            *reinterpret_cast<uint4*>(dst + kLboA) = make_uint4(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]), pack_bf16(lo[4], lo[5]), pack_bf16(lo[6], lo[7]));
            fence_async_proxy();
        }
    };
    load_obs(cta);
    stage_obs(0);
    load_obs(cta + stride);
    stage_obs(1);
    load_obs(cta + 2 * stride);

    fence_async_proxy();                                          // the weights were written through the generic proxy
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_lane = tmem + ((uint32_t)(32 * q) << 16);

    if (warp < kXWarps) {
        // =============================== team X: epilogue 1 ===========================================
        const int xh = warp >> 2;                                 // which 32 of a slice's 64 units
        int it = 0;
        for (int64_t blk = cta; blk < n_blocks; blk += stride, ++it) {
            const uint32_t ppar = (uint32_t)(it & 1) ^ 1u;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t cur[32];
                mbar_wait(l1bar + 8 * (c & 1), (uint32_t)(c >> 1));   // two completions per slot and block: parity = c / 2
                fence_after_sync();
                tmem_ld32_issue(t_lane + (uint32_t)(64 * (c & 1) + 32 * xh), cur);
                tmem_ld_wait(cur);
                // A2s slice c is free when the previous block's MMAs of both halves have read it
                if (it > 0 && c == 0) mbar_wait(e0, ppar);
                if (it > 0 && c == 1) mbar_wait(barA, ppar);
                if (it > 0 && c == 3) mbar_wait(barB, ppar);
                const int col0 = 64 * c + 32 * xh;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {                  // 8 units = one 16-byte chunk of the K-major operand
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float x0 = __uint_as_float(cur[8 * k4 + 2 * j]), x1 = __uint_as_float(cur[8 * k4 + 2 * j + 1]);
                        w[j] = (j < RL_MLP_X_POLY) ? tanh_poly_bf16x2(x0, x1) : tanh_bf16x2(pack_bf16(x0, x1));
                    }
                    const int kc = (col0 >> 3) + k4;
                    *reinterpret_cast<uint4*>(smem + kSmA2 + kc * kLboA + (row >> 3) * kSbo + (row & 7) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                }
                fence_async_proxy();                              // A2s writes -> visible to the tensor core
                fence_before_sync();                              // this thread's reads of ring slot c & 1 are complete
                __syncwarp();
                if (lane == 0) mbar_arrive(x_ready + 8 * c);
            }
        }
    } else if (warp < kXWarps + kYWarps) {
        // =============================== team Y: epilogue 2 and the head ==============================
        constexpr int kYPoly = kOut == 1 ? RL_MLP_Y_POLY_VALUE : RL_MLP_Y_POLY_POLICY;   // pairs in four on the FMA pipe
        const float* w3s = reinterpret_cast<const float*>(smem + kSmW3);
        const float* b3s = reinterpret_cast<const float*>(smem + kSmB3);
        float* red = reinterpret_cast<float*>(smem + kSmRed);
        int it = 0;
        for (int64_t blk = cta; blk < n_blocks; blk += stride, ++it) {
            const uint32_t par = (uint32_t)(it & 1);
            float2 s2[kOut];
#pragma unroll
            for (int o = 0; o < kOut; ++o) s2[o] = make_float2(0.f, 0.f);
#pragma unroll
            for (int nh = 0; nh < 2; ++nh) {
                // (the wait for this half was done before the arrival that released the previous one, see below)
                if (it == 0 && nh == 0) mbar_wait(barA, 0u);
                fence_after_sync();
                const uint32_t col = 128u + 128u * (uint32_t)((2 * it + nh) % 3) + (uint32_t)(kYUnits * yh);
#pragma unroll
                for (int c = 0; c < kYUnits / 32; ++c) {
                    uint32_t cur[32];
                    tmem_ld32_issue(t_lane + col + 32u * (uint32_t)c, cur);
                    tmem_ld_wait(cur);
                    const int unit0 = 128 * nh + kYUnits * yh + 32 * c;
#pragma unroll
                    for (int k = 0; k < 32; k += 4) {
                        float h0, h1, h2, h3;                             // pairs (k, k + 1), (k + 2, k + 3): pair index k / 2, k / 2 + 1
                        if (((k >> 1) & 3) < kYPoly) tanh_poly_f32x2(__uint_as_float(cur[k]), __uint_as_float(cur[k + 1]), h0, h1);
                        else h0 = tanh_fast(__uint_as_float(cur[k])), h1 = tanh_fast(__uint_as_float(cur[k + 1]));
                        if ((((k >> 1) + 1) & 3) < kYPoly) tanh_poly_f32x2(__uint_as_float(cur[k + 2]), __uint_as_float(cur[k + 3]), h2, h3);
                        else h2 = tanh_fast(__uint_as_float(cur[k + 2])), h3 = tanh_fast(__uint_as_float(cur[k + 3]));
#pragma unroll
                        for (int o = 0; o < kOut; ++o) {                  // even / odd units accumulate in the two halves of a packed FMA
                            const float4 w = *reinterpret_cast<const float4*>(w3s + o * kHidden + unit0 + k);
                            ffma2(s2[o], h0, h1, w.x, w.y);
                            ffma2(s2[o], h2, h3, w.z, w.w);
                        }
                    }
                }
                fence_before_sync();                              // this thread's reads of the accumulator are complete
                // Wait for the NEXT half before releasing this one: the commit that completes the next
                // half's barrier a second time is issued only after this release, so a parity wait can
                // never see its barrier two phases ahead.
                if (nh == 0) {
                    mbar_wait(barB, par);
                    // Team X has finished this block (barA), so layer 1 has read its A1 buffer: the observations of
                    // the block two ahead go there, announced by the arrival below (warp 16 waits for it before it
                    // issues that block's layer 1).
                    stage_obs(it & 1);
                    load_obs(blk + 3 * stride);
                } else if (blk + stride < n_blocks) {
                    mbar_wait(barA, par ^ 1u);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(nh ? y_doneB : y_doneA);
            }
            float s[kOut];
#pragma unroll
            for (int o = 0; o < kOut; ++o) s[o] = s2[o].x + s2[o].y;
            float* rbuf = red + (kRedBufs == 2 ? (it & 1) : 0) * ((kYSlices - 1) * kM * 2);   // the other slices' warps hand their sums to the last one
            if (yh < kYSlices - 1) {
#pragma unroll
                for (int o = 0; o < kOut; ++o) rbuf[(yh * kM + row) * 2 + o] = s[o];
            }
            quarter_barrier(q);
            float v[kOut];
            if (yh == kYSlices - 1) {
#pragma unroll
                for (int o = 0; o < kOut; ++o) {
                    v[o] = b3s[o] + s[o];
#pragma unroll
                    for (int sl = 0; sl < kYSlices - 1; ++sl) v[o] += rbuf[(sl * kM + row) * 2 + o];
                }
            }
            if (kRedBufs == 1) quarter_barrier(q);                // the single buffer has been read: the next block may overwrite it
            if (yh == kYSlices - 1) {
                const int64_t r = blk * kM + row;
                if (r < n) {
                    if (kOut == 2 && sa.actions != nullptr) {
                        // counter (global row, draw); the KEY is the seed xor a domain tag, so that this stream is
                        // independent of the environment's reset stream (key = seed, counter = (global id, epoch))
                        // even when both are given the same seed
                        const uint64_t gr = (uint64_t)r + sa.row_offset;
                        const uint4 bits = rl::philox4x32_10(make_uint4((uint32_t)gr, (uint32_t)(gr >> 32), (uint32_t)sa.draw,
                                                                        (uint32_t)(sa.draw >> 32)),
                                                             make_uint2((uint32_t)sa.seed ^ 0x706F6C69u, (uint32_t)(sa.seed >> 32) ^ 0x63792E61u));
                        const float u1 = (float)((bits.x >> 8) + 1u) * 5.9604644775390625e-8f;      // (0, 1]
                        const float u2 = (float)(bits.y >> 8) * 5.9604644775390625e-8f;             // [0, 1)
                        const float rad = sqrtf(-2.0f * logf(u1));
                        float sn, cs;
                        sincospif(2.0f * u2, &sn, &cs);
                        const float z0 = rad * cs, z1 = rad * sn;
                        const float ls0 = sa.log_std[0], ls1 = sa.log_std[1];
                        const float a0 = fmaf(expf(ls0), z0, v[0]), a1 = fmaf(expf(ls1), z1, v[kOut - 1]);
                        *reinterpret_cast<float2*>(sa.actions + 2 * r) = make_float2(a0, a1);
                        *reinterpret_cast<float2*>(sa.clipped + 2 * r) =
                            make_float2(fminf(fmaxf(a0, sa.lo0), sa.hi0), fminf(fmaxf(a1, sa.lo1), sa.hi1));
                        sa.log_prob[r] = (-0.5f * z0 * z0 - ls0 - 0.91893853320467274f) + (-0.5f * z1 * z1 - ls1 - 0.91893853320467274f);
                        if (sa.mean) *reinterpret_cast<float2*>(sa.mean + 2 * r) = make_float2(v[0], v[kOut - 1]);
                    } else {
#pragma unroll
                        for (int o = 0; o < kOut; ++o) out[r * kOut + o] = v[o];
                    }
                }
            }
            // (rbuf is rewritten two blocks later: all warps of the quarter pass another quarter_barrier in between)
        }
    } else {
        // =============================== the last warp: every MMA =============================================
        // The period of the whole pipeline cannot be shorter than this warp's own instruction stream
        // (42 MMAs, 10 commits, 10 waits per block), so it does nothing else and keeps its descriptors as
        // 32-bit low words that only need an add.
        const uint32_t a2_lo = desc_lo(sbase + kSmA2, kLboA), w2_lo = desc_lo(sbase + kSmW2, kLboW);
        const uint32_t a1_lo0 = desc_lo(sbase + kSmA1, kLboA), a1_lo1 = desc_lo(sbase + kSmA1b, kLboA);
        const uint32_t ones_lo = desc_lo(sbase + kSmOnes, kLboA), w1_lo = desc_lo(sbase + kSmW1, kLboW);
        const uint32_t bb1_lo = desc_lo(sbase + kSmBb1, kLboW), bb2_lo = desc_lo(sbase + kSmBb2, kLboW);
        // lane 0: layer 1 (bias, then A1 W1^T), units [64 u, 64 u + 64) -> ring slot u & 1
        auto issue_layer1 = [&](uint32_t a1_lo, int u) {
            const uint32_t d = tmem + 64u * (uint32_t)(u & 1), off = (uint32_t)(u * (64 / 8) * kSbo) >> 4;
            mma_lo(d, ones_lo, bb1_lo + off, idesc_n(64), 0u);
            mma_lo(d, a1_lo, w1_lo + off, idesc_n(64), 1u);
            mma_commit(l1bar + 8 * (u & 1));
        };
        // lane 0: layer 2, K slice c (64 units) of N half nh -> accumulator column `acc` (slice 0 starts from b2)
        auto issue_layer2 = [&](uint32_t acc, int nh, int c) {
            const uint32_t n_off = (uint32_t)(nh * (128 / 8) * kSbo) >> 4;
            if (c == 0) mma_lo(tmem + acc, ones_lo, bb2_lo + n_off, idesc_n(128), 0u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ks = 4 * c + j;
                mma_lo(tmem + acc, a2_lo + ((uint32_t)(ks * 2 * kLboA) >> 4), w2_lo + ((uint32_t)(ks * 2 * kLboW) >> 4) + n_off, idesc_n(128), 1u);
            }
        };
        if (elect_one()) {
            fence_after_sync();
            issue_layer1(a1_lo0, 0);
            issue_layer1(a1_lo0, 1);
        }
        mbar_wait(wbar, 0u);                                      // W2 has landed (the bulk copy was the kernel's first instruction)
        int it = 0;
        for (int64_t blk = cta; blk < n_blocks; blk += stride, ++it) {
            const bool has_next = blk + stride < n_blocks;
            const uint32_t par = (uint32_t)(it & 1), ppar = par ^ 1u;
            const uint32_t accA = 128u + 128u * (uint32_t)((2 * it) % 3), accB = 128u + 128u * (uint32_t)((2 * it + 1) % 3);
            const uint32_t a1_cur = (it & 1) ? a1_lo1 : a1_lo0, a1_nxt = (it & 1) ? a1_lo0 : a1_lo1;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                mbar_wait(x_ready + 8 * c, par);                  // team X: A2s slice c written, ring slot c & 1 drained
                // team Y: half A of the previous block drained -> accB is free; observations of the next block staged
                if (c == 2 && it > 0) mbar_wait(y_doneA, ppar);
                if (elect_one()) {
                    fence_after_sync();
                    if (c < 2) issue_layer1(a1_cur, c + 2);       // team X waits for this one first
                    else if (has_next) issue_layer1(a1_nxt, c - 2);
                    issue_layer2(accA, 0, c);                     // (accA was released by y_doneB two blocks ago: waited for below)
                    if (c == 3) mma_commit(barA);
                    if (c == 2) {
                        issue_layer2(accB, 1, 0);
                        mma_commit(e0);
                        issue_layer2(accB, 1, 1);
                        issue_layer2(accB, 1, 2);
                    }
                    if (c == 3) {
                        issue_layer2(accB, 1, 3);
                        mma_commit(barB);
                    }
                }
                __syncwarp();
            }
            if (it > 0) mbar_wait(y_doneB, ppar);                  // half B of the previous block drained -> free for half A of the next
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

template <int kOut>
__global__ void __launch_bounds__(kThreads, 1) mlp_forward_tc5_kernel(const unsigned char* __restrict__ blob,
                                                                const float* __restrict__ obs, float* __restrict__ out,
                                                                int64_t n, const SampleArgs sa) {
    mlp_body<kOut>(blob, obs, out, n, sa, (int64_t)blockIdx.x, (int64_t)gridDim.x);
}

// Both nets of the actor-critic in ONE launch: CTAs [0, n_pi) run the policy net (with or without the sampling
// epilogue), the others the value net -- one prologue (W2 staging, barriers, tensor-memory allocation) and one
// launch instead of two back to back; every CTA computes exactly what it would in a launch of its own net.
__global__ void __launch_bounds__(kThreads, 1) mlp_pair_kernel(const unsigned char* __restrict__ blob_pi,
                                                               const unsigned char* __restrict__ blob_vf,
                                                               const float* __restrict__ obs, float* __restrict__ out_pi,
                                                               float* __restrict__ values, int64_t n, const SampleArgs sa, int n_pi) {
    if ((int)blockIdx.x < n_pi) {
        mlp_body<2>(blob_pi, obs, out_pi, n, sa, (int64_t)blockIdx.x, (int64_t)n_pi);
    } else {
        const SampleArgs none = {};
        mlp_body<1>(blob_vf, obs, values, n, none, (int64_t)blockIdx.x - n_pi, (int64_t)gridDim.x - n_pi);
    }
}


}  // namespace tc5

}  // namespace

extern "C" {

const char* rl_policy_last_error(void) { return g_po_err; }

int64_t rl_mlp_blob_bytes(void) { return RL_MLP_BLOB_BYTES; }

namespace {
int launch_mlp(const void* blob, const float* obs, float* out, int64_t n, int out_dim, const SampleArgs& sa, cudaStream_t stream) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t blocks = (n + tc5::kM - 1) / tc5::kM;
    const int grid = (int)(blocks < sms ? blocks : sms);          // persistent: the weights are staged once per CTA
    cudaError_t e;
    if (out_dim == 2) {
        e = cudaFuncSetAttribute(tc5::mlp_forward_tc5_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc5::kSmemBytes);
        if (e == cudaSuccess)
            e = rl::rl_launch(tc5::mlp_forward_tc5_kernel<2>, dim3(grid), dim3(tc5::kThreads), tc5::kSmemBytes, stream, (const unsigned char*)blob, obs, out, n, sa);
    } else {
        e = cudaFuncSetAttribute(tc5::mlp_forward_tc5_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc5::kSmemBytes);
        if (e == cudaSuccess)
            e = rl::rl_launch(tc5::mlp_forward_tc5_kernel<1>, dim3(grid), dim3(tc5::kThreads), tc5::kSmemBytes, stream, (const unsigned char*)blob, obs, out, n, sa);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return po_fail(RL_E_CUDA, "rl_mlp_forward launch failed: %s", cudaGetErrorString(e));
    return RL_OK;
}
}  // namespace

namespace {
// policy net (sampling epilogue iff sa.actions) + value net in one launch
int launch_pair(const void* blob_pi, const void* blob_vf, const float* obs, float* out_pi, float* values, int64_t n,
                const SampleArgs& sa, cudaStream_t stream) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t blocks = (n + tc5::kM - 1) / tc5::kM;
    // CTAs so that both sides finish together: a row block of the policy net (second head row, sampling epilogue)
    // takes `ratio` = 1.18 times one of the value net (measured in the pair kernel, profiles/README.md); choose the
    // split that minimises max(ceil(blocks / n_pi) ratio, ceil(blocks / n_vf)) -- at most one CTA per row block
    // on either side.  RL_B200_PI_RATIO=<percent> overrides the ratio (experiment knob).
    const int64_t total = 2 * blocks < sms ? 2 * blocks : sms;
    static const double ratio = [] {
        const char* t = getenv("RL_B200_PI_RATIO");
        const long v = t ? strtol(t, nullptr, 10) : 0;
        return (v >= 50 && v <= 300) ? (double)v / 100.0 : 1.18;
    }();
    int64_t n_pi = 1, n_vf = 1;
    double best = 1e300;
    for (int64_t p = 1; p < total || p == 1; ++p) {
        const int64_t a_pi = p > blocks ? blocks : p;
        int64_t a_vf = total - p < 1 ? 1 : total - p;
        a_vf = a_vf > blocks ? blocks : a_vf;
        const double t_pi = (double)((blocks + a_pi - 1) / a_pi) * ratio, t_vf = (double)((blocks + a_vf - 1) / a_vf);
        const double t = t_pi > t_vf ? t_pi : t_vf;
        if (t < best - 1e-9) { best = t; n_pi = a_pi; n_vf = a_vf; }
    }
    cudaError_t e = cudaFuncSetAttribute(tc5::mlp_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc5::kSmemBytes);
    if (e == cudaSuccess)
        e = rl::rl_launch(tc5::mlp_pair_kernel, dim3((unsigned)(n_pi + n_vf)), dim3(tc5::kThreads), tc5::kSmemBytes, stream,
                          (const unsigned char*)blob_pi, (const unsigned char*)blob_vf, obs, out_pi, values, n, sa, (int)n_pi);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return po_fail(RL_E_CUDA, "rl_policy_value_sample launch failed: %s", cudaGetErrorString(e));
    return RL_OK;
}
}  // namespace

// out[n][out_dim] = head(tanh(W2 tanh(W1 obs + b1) + b2)) for the packed net `blob` (DEVICE memory,
// RL_MLP_BLOB_BYTES, 16-byte aligned); obs float32 [n][8] (16-byte aligned), out float32.  Enqueue only.
int rl_mlp_forward(const void* blob, const float* obs, float* out, int64_t n, int out_dim, void* stream) {
    if (!blob || !obs || !out) return po_fail(RL_E_INVALID, "rl_mlp_forward: null argument");
    if (n < 1) return po_fail(RL_E_INVALID, "rl_mlp_forward: n must be positive");
    if (out_dim != 1 && out_dim != 2) return po_fail(RL_E_INVALID, "rl_mlp_forward: out_dim must be 1 or 2 (got %d)", out_dim);
    if ((reinterpret_cast<uintptr_t>(blob) & 15) || (reinterpret_cast<uintptr_t>(obs) & 15))
        return po_fail(RL_E_INVALID, "rl_mlp_forward: blob and obs need 16-byte alignment");
    SampleArgs sa = {};
    return launch_mlp(blob, obs, out, n, out_dim, sa, (cudaStream_t)stream);
}

// The policy net with SB3's action sampling fused into its epilogue: actions[n][2] = mean +
// exp(log_std) eps (unclipped, what the rollout buffer stores), clipped[n][2] = the same clipped to
// [low, high] (what env.step receives), log_prob[n]; mean[n][2] optional.  eps is Philox4x32-10 keyed
// by `seed` (xor a domain tag) at counter (row_offset + row, draw): pass a different `draw` every call and the
// shard's global id offset as row_offset.  DEVICE pointers, enqueue only.
int rl_policy_sample(const void* blob, const float* obs, const float* log_std, uint64_t seed, uint64_t draw,
                     int64_t row_offset, const float* low2, const float* high2, float* actions, float* clipped,
                     float* log_prob, float* mean, int64_t n, void* stream) {
    if (!blob || !obs || !log_std || !low2 || !high2 || !actions || !clipped || !log_prob)
        return po_fail(RL_E_INVALID, "rl_policy_sample: null argument");
    if (n < 1) return po_fail(RL_E_INVALID, "rl_policy_sample: n must be positive");
    if (row_offset < 0) return po_fail(RL_E_INVALID, "rl_policy_sample: row_offset must be >= 0");
    if ((reinterpret_cast<uintptr_t>(blob) & 15) || (reinterpret_cast<uintptr_t>(obs) & 15) ||
        (reinterpret_cast<uintptr_t>(actions) & 7) || (reinterpret_cast<uintptr_t>(clipped) & 7) ||
        (mean && (reinterpret_cast<uintptr_t>(mean) & 7)))
        return po_fail(RL_E_INVALID, "rl_policy_sample: blob / obs need 16-byte, actions / clipped / mean 8-byte alignment");
    SampleArgs sa = {};
    sa.log_std = log_std; sa.actions = actions; sa.clipped = clipped; sa.log_prob = log_prob; sa.mean = mean;
    sa.seed = seed; sa.draw = draw; sa.row_offset = (unsigned long long)row_offset;
    sa.lo0 = low2[0]; sa.lo1 = low2[1]; sa.hi0 = high2[0]; sa.hi1 = high2[1];      // HOST pointers: two floats each
    return launch_mlp(blob, obs, actions, n, 2, sa, (cudaStream_t)stream);
}

// rl_policy_sample and the value net (values[n] = rl_mlp_forward(blob_vf, obs, ..., out_dim 1)) in ONE launch: what
// a rollout step asks of the actor-critic.  Results are bit-identical to the two separate calls.
int rl_policy_value_sample(const void* blob_pi, const void* blob_vf, const float* obs, const float* log_std, uint64_t seed,
                           uint64_t draw, int64_t row_offset, const float* low2, const float* high2, float* actions,
                           float* clipped, float* log_prob, float* mean, float* values, int64_t n, void* stream) {
    if (!blob_pi || !blob_vf || !obs || !log_std || !low2 || !high2 || !actions || !clipped || !log_prob || !values)
        return po_fail(RL_E_INVALID, "rl_policy_value_sample: null argument");
    if (n < 1) return po_fail(RL_E_INVALID, "rl_policy_value_sample: n must be positive");
    if (row_offset < 0) return po_fail(RL_E_INVALID, "rl_policy_value_sample: row_offset must be >= 0");
    if ((reinterpret_cast<uintptr_t>(blob_pi) & 15) || (reinterpret_cast<uintptr_t>(blob_vf) & 15) ||
        (reinterpret_cast<uintptr_t>(obs) & 15) || (reinterpret_cast<uintptr_t>(actions) & 7) ||
        (reinterpret_cast<uintptr_t>(clipped) & 7) || (mean && (reinterpret_cast<uintptr_t>(mean) & 7)))
        return po_fail(RL_E_INVALID, "rl_policy_value_sample: blobs / obs need 16-byte, actions / clipped / mean 8-byte alignment");
    SampleArgs sa = {};
    sa.log_std = log_std; sa.actions = actions; sa.clipped = clipped; sa.log_prob = log_prob; sa.mean = mean;
    sa.seed = seed; sa.draw = draw; sa.row_offset = (unsigned long long)row_offset;
    sa.lo0 = low2[0]; sa.lo1 = low2[1]; sa.hi0 = high2[0]; sa.hi1 = high2[1];      // HOST pointers: two floats each
    return launch_pair(blob_pi, blob_vf, obs, actions, values, n, sa, (cudaStream_t)stream);
}

}  // extern "C"
