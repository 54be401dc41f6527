// TEST INFRASTRUCTURE ONLY -- host build of the DEVICE math (rl_device.cuh) for a logic check
// that runs without a GPU.  It is not part of the product library, is never loaded by the
// package, and is not a CPU fallback: it exists so the CPU test tier can compare the fp32
// step formulation (delta-form Verlet, wrap bookkeeping, reward branches, Philox reset)
// against the fp64 oracle on the committed golden vectors.  CUDA intrinsics are shimmed
// below; results agree with the GPU to rounding (sincospif / implicit FMA contraction
// differ in the last bits), not bit-for-bit.
#define RL_HOST_EMULATION 1
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return {x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __fdividef(float a, float b) { return a / b; }
using std::min;

#include "../../rocket-landing-rl_b200/csrc/rl_device.cuh"

// soa: float32 [10][n] public state format, updated in place.  Outputs as rl_step (no reset).
// folded != 0 runs the instantiation with the reference's default constants folded in
// (rl::DefaultParams), which the library only selects when matches_default() holds.
template <class PP>
static void emu_step_impl(const PP& P, float* soa, int64_t n, const float* actions, float* obs, float* reward,
                          uint8_t* terminated, uint8_t* truncated, int8_t* landing, int substeps) {
    for (int64_t i = 0; i < n; ++i) {
        rl::Rocket r;
        r.x = soa[RL_W_X * n + i]; r.y = soa[RL_W_Y * n + i]; r.ang = soa[RL_W_ANGLE * n + i];
        r.sx = soa[RL_W_SX * n + i]; r.sy = soa[RL_W_SY * n + i]; r.sang = soa[RL_W_SANGLE * n + i];
        r.dry = soa[RL_W_DRY * n + i]; r.fuel = soa[RL_W_FUEL * n + i];
        std::memcpy(&r.meta, &soa[RL_W_META * n + i], 4); r.ep_ret = soa[RL_W_EPRET * n + i];
        rl::StepResult o;
        float rew = 0.0f;
        for (int k = 0; k < substeps; ++k) {
            rl::rocket_step(P, r, actions[2 * i], actions[2 * i + 1], o);
            rew += o.reward;
            if (o.terminated || o.truncated) break;
        }
        soa[RL_W_X * n + i] = r.x; soa[RL_W_Y * n + i] = r.y; soa[RL_W_ANGLE * n + i] = r.ang;
        soa[RL_W_SX * n + i] = r.sx; soa[RL_W_SY * n + i] = r.sy; soa[RL_W_SANGLE * n + i] = r.sang;
        soa[RL_W_DRY * n + i] = r.dry; soa[RL_W_FUEL * n + i] = r.fuel;
        std::memcpy(&soa[RL_W_META * n + i], &r.meta, 4); soa[RL_W_EPRET * n + i] = r.ep_ret;
        std::memcpy(obs + 8 * i, o.obs, 32);
        reward[i] = rew;
        terminated[i] = o.terminated; truncated[i] = o.truncated; landing[i] = (int8_t)o.landing;
    }
}

extern "C" {

int emu_step(const RlParams* p, float* soa, int64_t n, const float* actions, float* obs, float* reward,
             uint8_t* terminated, uint8_t* truncated, int8_t* landing, int substeps, int folded) {
    rl::DevParams P;
    char err[256];
    if (rl::derive_params(*p, P, err, sizeof(err)) != RL_OK) return -1;
    if (folded) {
        if (!rl::matches_default(P)) return -2;
        emu_step_impl(rl::DefaultParams{}, soa, n, actions, obs, reward, terminated, truncated, landing, substeps);
    } else {
        emu_step_impl(P, soa, n, actions, obs, reward, terminated, truncated, landing, substeps);
    }
    return 0;
}

// Precise-altitude mode (rocket_step<PP, true>): `lo` is float32 [2][n] = y_lo, sy_lo, in/out.
int emu_step_precise(const RlParams* p, float* soa, float* lo, int64_t n, const float* actions, float* obs,
                     float* reward, uint8_t* terminated, uint8_t* truncated, int8_t* landing) {
    rl::DevParams P;
    char err[256];
    if (rl::derive_params(*p, P, err, sizeof(err)) != RL_OK) return -1;
    for (int64_t i = 0; i < n; ++i) {
        rl::Rocket r;
        r.x = soa[RL_W_X * n + i]; r.y = soa[RL_W_Y * n + i]; r.ang = soa[RL_W_ANGLE * n + i];
        r.sx = soa[RL_W_SX * n + i]; r.sy = soa[RL_W_SY * n + i]; r.sang = soa[RL_W_SANGLE * n + i];
        r.dry = soa[RL_W_DRY * n + i]; r.fuel = soa[RL_W_FUEL * n + i];
        std::memcpy(&r.meta, &soa[RL_W_META * n + i], 4); r.ep_ret = soa[RL_W_EPRET * n + i];
        r.y_lo = lo[i]; r.sy_lo = lo[n + i];
        rl::StepResult o;
        rl::rocket_step<rl::DevParams, true>(P, r, actions[2 * i], actions[2 * i + 1], o);
        soa[RL_W_X * n + i] = r.x; soa[RL_W_Y * n + i] = r.y; soa[RL_W_ANGLE * n + i] = r.ang;
        soa[RL_W_SX * n + i] = r.sx; soa[RL_W_SY * n + i] = r.sy; soa[RL_W_SANGLE * n + i] = r.sang;
        soa[RL_W_DRY * n + i] = r.dry; soa[RL_W_FUEL * n + i] = r.fuel;
        std::memcpy(&soa[RL_W_META * n + i], &r.meta, 4); soa[RL_W_EPRET * n + i] = r.ep_ret;
        lo[i] = r.y_lo; lo[n + i] = r.sy_lo;
        std::memcpy(obs + 8 * i, o.obs, 32);
        reward[i] = o.reward;
        terminated[i] = o.terminated; truncated[i] = o.truncated; landing[i] = (int8_t)o.landing;
    }
    return 0;
}

// One simulation-mode tick (rl_sim.cuh semantics) on the host: soa in/out, telemetry [n][16].
int emu_sim_step(const RlParams* p, float* soa, int64_t n, const float* actions, float* telemetry) {
    rl::DevParams P;
    char err[256];
    if (rl::derive_params(*p, P, err, sizeof(err)) != RL_OK) return -1;
    const uint32_t kLanded = 0x40000000u;
    for (int64_t i = 0; i < n; ++i) {
        rl::Rocket r;
        r.x = soa[RL_W_X * n + i]; r.y = soa[RL_W_Y * n + i]; r.ang = soa[RL_W_ANGLE * n + i];
        r.sx = soa[RL_W_SX * n + i]; r.sy = soa[RL_W_SY * n + i]; r.sang = soa[RL_W_SANGLE * n + i];
        r.dry = soa[RL_W_DRY * n + i]; r.fuel = soa[RL_W_FUEL * n + i];
        std::memcpy(&r.meta, &soa[RL_W_META * n + i], 4); r.ep_ret = soa[RL_W_EPRET * n + i];
        float* rec = telemetry + 16 * i;
        if (r.meta & kLanded) {
            for (int k = 0; k < 16; ++k) rec[k] = 0.0f;
            rec[11] = std::nanf("");
            rec[14] = (float)((r.meta >> 27) & 7u);
            continue;
        }
        const float th = std::fmin(std::fmax(actions[2 * i], 0.0f), 1.0f);
        const float cg = std::fmin(std::fmax(actions[2 * i + 1], -1.0f), 1.0f);
        rl::StepResult o;
        rl::rocket_step(P, r, th, cg, o);
        const bool landed = o.terminated;
        if (landed) r.meta |= kLanded | ((uint32_t)o.landing << 27);
        soa[RL_W_X * n + i] = r.x; soa[RL_W_Y * n + i] = r.y; soa[RL_W_ANGLE * n + i] = r.ang;
        soa[RL_W_SX * n + i] = r.sx; soa[RL_W_SY * n + i] = r.sy; soa[RL_W_SANGLE * n + i] = r.sang;
        soa[RL_W_FUEL * n + i] = r.fuel; std::memcpy(&soa[RL_W_META * n + i], &r.meta, 4);
        soa[RL_W_EPRET * n + i] = r.ep_ret;
        const float rv[16] = {r.x, r.y, r.sx * P.inv_dt, r.sy * P.inv_dt, o.ax, o.ay, r.ang, r.sang * P.inv_dt,
                              o.alpha, r.dry, r.fuel, o.reward, landed ? 0.0f : actions[2 * i],
                              landed ? 0.0f : actions[2 * i + 1], (float)o.landing, 1.0f};
        std::memcpy(rec, rv, sizeof(rv));
    }
    return 0;
}

int emu_matches_default(const RlParams* p) {
    rl::DevParams P;
    char err[256];
    if (rl::derive_params(*p, P, err, sizeof(err)) != RL_OK) return -1;
    return rl::matches_default(P) ? 1 : 0;
}

// Philox reset of n rockets with global ids gid0.. at `epoch`; writes soa [10][n] and obs [n][8].
int emu_reset(const RlParams* p, float* soa, int64_t n, int64_t gid0, uint64_t seed, uint64_t epoch, float* obs) {
    rl::DevParams P;
    char err[256];
    if (rl::derive_params(*p, P, err, sizeof(err)) != RL_OK) return -1;
    for (int64_t i = 0; i < n; ++i) {
        rl::Rocket r;
        float ax0, ay0;
        rl::sample_reset(P, seed, (uint64_t)(gid0 + i), epoch, r, ax0, ay0);
        soa[RL_W_X * n + i] = r.x; soa[RL_W_Y * n + i] = r.y; soa[RL_W_ANGLE * n + i] = r.ang;
        soa[RL_W_SX * n + i] = r.sx; soa[RL_W_SY * n + i] = r.sy; soa[RL_W_SANGLE * n + i] = r.sang;
        soa[RL_W_DRY * n + i] = r.dry; soa[RL_W_FUEL * n + i] = r.fuel;
        std::memcpy(&soa[RL_W_META * n + i], &r.meta, 4); soa[RL_W_EPRET * n + i] = r.ep_ret;
        rl::fresh_obs(P, r, ax0, ay0, obs + 8 * i);
    }
    return 0;
}

void emu_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint4 r = rl::philox4x32_10(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

}  // extern "C"
