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
using std::max;
using std::fma;
using std::floor;

#include "../../rocket-landing-rl_b200/csrc/rl_device.cuh"

static inline uint32_t w32(const float* soa, int64_t n, int w, int64_t i) { uint32_t u; std::memcpy(&u, &soa[w * n + i], 4); return u; }
static inline void s32(float* soa, int64_t n, int w, int64_t i, uint32_t u) { std::memcpy(&soa[w * n + i], &u, 4); }
static inline float wf(const float* soa, int64_t n, int w, int64_t i) { return soa[w * n + i]; }

static void load_rocket(const float* soa, int64_t n, int64_t i, rl::Rocket& r) {
    r.x = (int32_t)w32(soa, n, RL_W_X, i); r.y = (int32_t)w32(soa, n, RL_W_Y, i); r.ang = (int32_t)w32(soa, n, RL_W_ANGLE, i);
    r.sx = (int32_t)w32(soa, n, RL_W_SX, i); r.sy = (int32_t)w32(soa, n, RL_W_SY, i); r.sang = wf(soa, n, RL_W_SANGLE, i);
    r.dry = wf(soa, n, RL_W_DRY, i); r.fuel = (int32_t)w32(soa, n, RL_W_FUEL, i);
    r.meta = w32(soa, n, RL_W_META, i); r.ep_ret = wf(soa, n, RL_W_EPRET, i);
}
static void store_rocket(float* soa, int64_t n, int64_t i, const rl::Rocket& r) {
    s32(soa, n, RL_W_X, i, (uint32_t)r.x); s32(soa, n, RL_W_Y, i, (uint32_t)r.y); s32(soa, n, RL_W_ANGLE, i, (uint32_t)r.ang);
    s32(soa, n, RL_W_SX, i, (uint32_t)r.sx); s32(soa, n, RL_W_SY, i, (uint32_t)r.sy); soa[RL_W_SANGLE * n + i] = r.sang;
    soa[RL_W_DRY * n + i] = r.dry; s32(soa, n, RL_W_FUEL, i, (uint32_t)r.fuel);
    s32(soa, n, RL_W_META, i, r.meta); soa[RL_W_EPRET * n + i] = r.ep_ret;
}

// soa: 32-bit words [10][n] public state format, updated in place.  Outputs as rl_step (no reset).
// folded != 0 runs the instantiation with the reference's default constants folded in
// (rl::DefaultParams), which the library only selects when matches_default() holds.
template <class PP>
static void emu_step_impl(const PP& P, float* soa, int64_t n, const float* actions, float* obs, float* reward,
                          uint8_t* terminated, uint8_t* truncated, int8_t* landing, int substeps) {
    for (int64_t i = 0; i < n; ++i) {
        rl::Rocket r;
        load_rocket(soa, n, i, r);
        rl::StepResult o;
        rl::Flight f;
        float rew = 0.0f;
        rl::flight_begin(P, r, f);
        for (int k = 0; k < substeps; ++k) {           // as the fused-substep loop of the step kernel
            rl::rocket_advance(P, r, f, actions[2 * i], actions[2 * i + 1], o);
            rew += o.reward;
            if (o.terminated || o.truncated) break;
        }
        rl::flight_end(r, f);
        rl::step_obs(P, r, f, o);
        store_rocket(soa, n, i, r);
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

// One simulation-mode tick (rl_sim.cuh semantics) on the host: soa in/out, telemetry [n][16].
int emu_sim_step(const RlParams* p, float* soa, int64_t n, const float* actions, float* telemetry) {
    rl::DevParams P;
    char err[256];
    if (rl::derive_params(*p, P, err, sizeof(err)) != RL_OK) return -1;
    const uint32_t kLanded = 0x40000000u;
    for (int64_t i = 0; i < n; ++i) {
        rl::Rocket r;
        load_rocket(soa, n, i, r);
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
        rl::rocket_step<rl::DevParams, true>(P, r, th, cg, o);
        const bool landed = o.terminated;
        if (landed) r.meta |= kLanded | ((uint32_t)o.landing << 27);
        store_rocket(soa, n, i, r);
        const float rv[16] = {(float)r.x * P.pos_unit, (float)r.y * P.pos_unit, o.vx, o.vy, o.ax, o.ay, rl::angle_deg(r.ang), o.om,
                              o.alpha, r.dry, (float)((double)r.fuel * P.fuel_unit_d), o.reward, landed ? 0.0f : actions[2 * i],
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
        rl::sample_reset(P, seed, (uint64_t)(gid0 + i), epoch, r, obs + 8 * i);
        store_rocket(soa, n, i, r);
    }
    return 0;
}

void emu_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint4 r = rl::philox4x32_10(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

}  // extern "C"
