// Device-side rocket step: physics + reward/termination + observation + Philox reset.
//
// One rocket lives entirely in registers for the duration of a launch.  The integrator is the
// reference's position-Verlet rewritten in DELTA FORM (DESIGN.md "fp32 formulation"):
//
//   reference (engine.py:181-222)                       here
//   x' = 2x - x_prev + ax dt^2                          dx' = dx + ax dt^2 ; x' = x + dx'
//   vx' = (x' - x) / dt                                 vx' = dx' / dt
//   th' = th + (th - th_prev) damp + alpha dt^2         dth' = (dth - 360 k) damp + alpha dt^2
//   om' = (th' - th) / dt   (before normalising)        om' = dth' / dt
//   th' = normalise(th')                                th' = th + dth' - 360 k'
//
// where k is the number of 360-degree wraps the PREVIOUS normalisation applied: the reference
// keeps the normalised angle as the next step's "previous" angle (base.py:171), so after a
// wrap its (th - th_prev) contains a spurious -+360; subtracting 360 k reproduces that exactly.
//
// PRECISION (DESIGN.md "Long-horizon accuracy").  Forces, reward and observation are fp32.  The
// words that INTEGRATE -- and whose rounding errors therefore persist for the rest of an episode
// of up to 1000 steps -- carry 32 significant bits instead of fp32's 24, at no extra bytes:
//   angle     int32, 360 / 2^32 deg per unit: the int32 range IS [-180, 180), wrap-around is the
//             normalisation, resolution 8.4e-8 deg everywhere;
//   x, y      int32, P.pos_unit m per unit (2^-15 m for the reference's +-50 km world): fp32 positions
//             at 2 km round by 1.2e-4 m every step, a random walk of 4e-3 m over 1000 steps;
//   dx, dy    int32, pos_unit / 256 per unit (2^-23 m: +-256 m per step);
//   fuel      int32, P.fuel_unit kg per unit (2^-12 kg for the reference's tanks);
//   dtheta    fp32 + 8 low-order bits kept in the meta word (a 32-bit mantissa), updated in fp64
//             together with 1 / m and alpha dt^2: an error e in dtheta decays only by the 0.995
//             damping, i.e. moves the angle by 200 e, and the angle steers the thrust.
// Measured on 1000-step hover episodes against the fp64 oracle the plain-fp32 version of these
// five words drifts to 3.1x the 1e-5 scale-relative tolerance (vx); this version stays at 0.2x.
//
// The always-executed path is written with selects instead of branches (the kernel is issue- as
// much as bandwidth-limited, and BSSY/BSYNC/BRA around short divergent regions cost more than
// the few flops they skip); only genuinely rare events branch: the first step of a FRESH rocket,
// touchdown, an angle wrap, a zero-mass rocket.
#pragma once
#ifndef RL_HOST_EMULATION
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "rl_params.h"

namespace rl {

struct Rocket {                // the ten persistent words (RL_W_*)
    int32_t x, y;              // units of P.pos_unit m
    int32_t ang;               // units of kAngleUnit deg
    int32_t sx, sy;            // x - x_prev, y - y_prev in units of P.delta_unit m (v dt while FRESH: bootstrap pending)
    float sang;                // carried pre-wrap angle delta, high part (omega dt while FRESH)
    float dry;
    int32_t fuel;              // units of P.fuel_unit_d kg (signed: an injected negative load acts like the
                               // reference's, base.py:138-144)
    float ep_ret;
    uint32_t meta;
};

struct StepResult {
    float obs[8];
    float reward;
    float ax, ay, alpha;       // raw accelerations of this step (telemetry: protocol.py:9-14)
    float vx, vy, om;          // velocities after the step (unclipped)
    float phi;                 // potential of the state after the step (reward.py:231-264)
    bool terminated, truncated, out_of_bounds, tipped, nonfinite;
    int landing;               // 0 none, 1 safe, 2 good, 3 ok, 4 unsafe (protocol.py:31-41)
};

// What a step needs of the state BEFORE it, derived from the persistent words: computed once when a
// rocket is loaded (flight_begin) and handed from one fused substep to the next by rocket_advance --
// every field is produced by the very expression flight_begin would evaluate on the stored words, so
// carrying it changes no bit of the results (the fused and the repeated single step are tested to be
// identical).  The meta word is unpacked here and packed back by flight_end.
struct Flight {
    float x, y;                // position in metres (fp32 image of the fixed-point words)
    float vx, vy, om;          // velocities: delta / dt (engine.py:193-194, :220-222)
    float ang;                 // angle in degrees
    float phi;                 // potential of this state (reward.py:231-264)
    int steps;                 // episode step count (lander.py:229)
    int k;                     // 360-degree turns the last normalisation removed (base.py:171)
    int lo;                    // low-order 8 bits of the carried angle delta
    bool fresh;                // bootstrap term still to be subtracted (base.py:65-130)
    uint32_t sim_bits;         // simulation-mode bits of the meta word, passed through
};

// ---- small math helpers (plain-C versions in the host build of tests/hostemu) ------------------
#ifndef RL_HOST_EMULATION
// 1/x: MUFU.RCP + one Newton step (< 1 ulp); __frcp_rn costs ~10 instructions for the last half ulp
__device__ __forceinline__ float rl_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
// sqrt(x), ~1 ulp (MUFU); only feeds the drag term, ~1e-7 of the acceleration
__device__ __forceinline__ float rl_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// 1/x for x >= 1 (no denormal / overflow handling needed): bare MUFU.RCP, 1 ulp
__device__ __forceinline__ float rl_rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// clamp to [0, 1] as an instruction modifier (FMUL.SAT / FFMA.SAT): min(1, x) for free when x >= 0
// is known, max(0, x) for free when x <= 1 is known
__device__ __forceinline__ float rl_sat(float x) { return __saturatef(x); }
__device__ __forceinline__ uint32_t rl_f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float rl_u2f(uint32_t u) { return __uint_as_float(u); }
// 1/m to ~1e-14 relative: MUFU.RCP of the fp32-rounded mass + one Newton step in fp64
__device__ __forceinline__ double rl_rcp_d(double m) {
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"((float)m));
    const double r = (double)r0;
    return fma(r, fma(-m, r, 1.0), r);
}
__device__ __forceinline__ int32_t rl_f2i(float f) { return __float2int_rn(f); }          // saturating, NaN -> 0

__device__ __forceinline__ int32_t rl_d2i(double d) { return __double2int_rn(d); }
__device__ __forceinline__ long long rl_d2ll(double d) { return __double2ll_rn(d); }
__device__ __forceinline__ double rl_pow2_bits(uint32_t biased) { return __hiloint2double((int)(biased << 20), 0); }
#else
inline float rl_rcp(float x) { return 1.0f / x; }
inline float rl_sqrt(float x) { return std::sqrt(x); }
inline float rl_rcp_fast(float x) { return 1.0f / x; }
inline float rl_sat(float x) { return x != x ? 0.0f : std::fmin(std::fmax(x, 0.0f), 1.0f); }
inline uint32_t rl_f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float rl_u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline double rl_rcp_d(double m) { const double r = (double)(1.0f / (float)m); return std::fma(r, std::fma(-m, r, 1.0), r); }
inline int32_t rl_f2i(float f) {
    if (f != f) return 0;
    const double r = std::nearbyint((double)f);
    return r >= 2147483647.0 ? INT32_MAX : (r <= -2147483648.0 ? INT32_MIN : (int32_t)r);
}
inline int32_t rl_d2i(double d) {
    if (d != d) return 0;
    const double r = std::nearbyint(d);
    return r >= 2147483647.0 ? INT32_MAX : (r <= -2147483648.0 ? INT32_MIN : (int32_t)r);
}
inline long long rl_d2ll(double d) {
    if (d != d) return 0;
    const double r = std::nearbyint(d);
    return r >= 9.2e18 ? INT64_MAX : (r <= -9.2e18 ? INT64_MIN : (long long)r);
}
inline double rl_pow2_bits(uint32_t biased) { const uint64_t b = (uint64_t)biased << 52; double d; std::memcpy(&d, &b, 8); return d; }
#endif

// int32 add / subtract with two's-complement wrap-around (signed overflow is undefined in C++)
__device__ __forceinline__ int32_t rl_wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
__device__ __forceinline__ int32_t rl_wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

// ---- the 32-bit-mantissa angle delta: sang (fp32) + lo * ulp(sang) / 256, lo in [-128, 127] ----
// unit of the low part of a value whose high part has the biased fp32 exponent E: 2^(E - 127 - 31)
__device__ __forceinline__ double dtheta_join(float hi, int lo) {
    const uint32_t E = (rl_f2u(hi) >> 23) & 0xffu;
    return fma((double)lo, rl_pow2_bits(E + (1023u - 158u)), (double)hi);
}
// (angle - previous angle) when the previous normalisation removed k turns; kept out of line of the
// compiler's if-conversion: wraps are rare and the fp64 work should not run on every step
__device__ __forceinline__ double rl_unwrap(double dang, int k) {
#ifndef RL_HOST_EMULATION
    asm volatile("");
#endif
    return dang - 360.0 * (double)k;
}
__device__ __forceinline__ void dtheta_split(double v, float& hi, int& lo) {
    hi = (float)v;
    const uint32_t E = (rl_f2u(hi) >> 23) & 0xffu;
    lo = rl_d2i((v - (double)hi) * rl_pow2_bits((1023u + 158u) - E));
    lo = min(max(lo, -128), 127);
}

// sin and cos of the fixed-point angle (the reference does np.sin(np.deg2rad(a)), engine.py:64-73).
// The quadrant is the top two bits of the integer (exact), the remainder in [-45, 45] deg goes to
// radians with one multiply, then degree-7 / degree-8 minimax polynomials on [-pi/4, pi/4]; max
// abs error 8.7e-8 (measured against fp64), i.e. that of sincospif, in ~22 branch-free instructions.
__device__ __forceinline__ void sincos_units(int32_t a, float& s, float& c) {
    const uint32_t qu = ((uint32_t)a + 0x20000000u) >> 30;             // quadrant mod 4 (wraps at +-180)
    const int32_t rem = (int32_t)((uint32_t)a - (qu << 30));           // [-2^29, 2^29) = [-45, 45) deg
    const float t = (float)rem * 1.4629180792671596e-9f;               // pi / 2^31 rad per unit
    const float u = t * t;
    const float sn = fmaf(t * u, fmaf(fmaf(-0.0001958794891834259f, u, 0.008332748897373676f), u, -0.166666641831398f), t);
    const float cs = fmaf(u, fmaf(fmaf(fmaf(2.446382313792128e-05f, u, -0.001388759003020823f), u, 0.04166664928197861f), u, -0.5f), 1.0f);
    const bool swap = (qu & 1u) != 0u;
    const float ss = swap ? cs : sn, cc = swap ? sn : cs;
    s = rl_u2f(rl_f2u(ss) ^ ((qu & 2u) << 30));                        // quadrants 2, 3: sin flips sign
    c = rl_u2f(rl_f2u(cc) ^ (((qu + 1u) & 2u) << 30));                 // quadrants 1, 2: cos flips sign
}

__device__ __forceinline__ float angle_deg(int32_t a) { return (float)a * (float)kAngleUnit; }   // 45 * 2^-29: exact constant

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based: no per-rocket generator state ----
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// Counter of Philox block `blk` (0..2) of the reset of global rocket `gid` at launch `epoch`.
__device__ __forceinline__ uint4 reset_counter(uint64_t gid, uint64_t epoch, uint32_t blk) {
    return make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)epoch, ((uint32_t)(epoch >> 32) << 2) | blk);
}

// 24 random bits -> [0, 1) exactly representable in fp32
__device__ __forceinline__ float u01(uint32_t bits) { return (float)(bits >> 8) * 5.9604644775390625e-8f; }

__device__ __forceinline__ float clipf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// _get_obs (lander.py:124-150)
template <class PP>
__device__ __forceinline__ void clip_obs(const PP& P, float x, float y, float vx, float vy,
                                         float ax, float ay, float ang, float om, float* obs) {
    obs[0] = clipf(x, P.obs_lo(0), P.obs_hi(0));
    obs[1] = clipf(y, P.obs_lo(1), P.obs_hi(1));
    obs[2] = clipf(vx, P.obs_lo(2), P.obs_hi(2));
    obs[3] = clipf(vy, P.obs_lo(3), P.obs_hi(3));
    obs[4] = clipf(ax, P.obs_lo(4), P.obs_hi(4));
    obs[5] = clipf(ay, P.obs_lo(5), P.obs_hi(5));
    obs[6] = clipf(ang, P.obs_lo(6), P.obs_hi(6));
    obs[7] = clipf(om, P.obs_lo(7), P.obs_hi(7));
}

// get_initial_state (simulation/config.py:26-53): ten independent uniforms lo + span u, in the
// reference's draw order x, y, vx, vy, ax, ay, angle, angular velocity, dry mass, fuel mass, from
// the ten random words w[0..9].  Writes the clipped reset observation (lander.py:182: the sampled
// ax, ay only ever appear there) when obs != nullptr.
// A FRESH rocket carries v dt in its delta slots (so v = delta / dt like any other rocket) and
// omega dt in the angle-delta slot; the Verlet bootstrap correction -a0 dt^2 / 2 (base.py:65-130) is
// applied by its first step, which evaluates the zero-control acceleration a0 anyway.
template <class PP>
__device__ __forceinline__ void rocket_from_words(const PP& P, const uint32_t* w, Rocket& r, float* obs) {
    const float vx = fmaf(P.reset_span(2), u01(w[2]), P.reset_lo(2));
    const float vy = fmaf(P.reset_span(3), u01(w[3]), P.reset_lo(3));
    const float ax0 = fmaf(P.reset_span(4), u01(w[4]), P.reset_lo(4));
    const float ay0 = fmaf(P.reset_span(5), u01(w[5]), P.reset_lo(5));
    const float ang = fmaf(P.reset_span(6), u01(w[6]), P.reset_lo(6));
    const float om = fmaf(P.reset_span(7), u01(w[7]), P.reset_lo(7));
    r.x = rl_f2i(fmaf(P.reset_span(0), u01(w[0]), P.reset_lo(0)) * P.pos_inv_unit);
    r.y = rl_f2i(fmaf(P.reset_span(1), u01(w[1]), P.reset_lo(1)) * P.pos_inv_unit);
    r.sx = rl_f2i(vx * P.dt_units);
    r.sy = rl_f2i(vy * P.dt_units);
    r.ang = rl_f2i(ang * (float)kAngleInvUnit);
    r.sang = om * P.dt;
    r.dry = fmaf(P.reset_span(8), u01(w[8]), P.reset_lo(8));
    r.fuel = rl_f2i(fmaf(P.reset_span(9), u01(w[9]), P.reset_lo(9)) * P.fuel_inv_unit);
    r.ep_ret = 0.0f;
    r.meta = RL_META_FRESH;                                      // current_step = 0 (lander.py:180)
    // the observation reads the STORED words back (as lander.py:124-150 reads Rocket.state), so it is
    // exactly the state the first step starts from
    if (obs)
        clip_obs(P, (float)r.x * P.pos_unit, (float)r.y * P.pos_unit, ((float)r.sx * P.delta_unit) * P.inv_dt,
                 ((float)r.sy * P.delta_unit) * P.inv_dt, ax0, ay0, angle_deg(r.ang), r.sang * P.inv_dt, obs);
}

// One thread draws all three Philox blocks keyed by (seed, global rocket id, epoch).
template <class PP>
__device__ __forceinline__ void sample_reset(const PP& P, uint64_t seed, uint64_t gid,
                                             uint64_t epoch, Rocket& r, float* obs) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint4 a = philox4x32_10(reset_counter(gid, epoch, 0u), key);
    const uint4 b = philox4x32_10(reset_counter(gid, epoch, 1u), key);
    const uint4 c = philox4x32_10(reset_counter(gid, epoch, 2u), key);
    const uint32_t w[10] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y};
    rocket_from_words(P, w, r, obs);
}

// Potential of reward.py:231-264.
__device__ __forceinline__ float potential(float y, float vx, float vy, float ang, float om) {
    const float yp = fmaxf(0.0f, y);
    const float w = 2.0f - rl_sat(y * (1.0f / 2000.0f));          // 2 - min(1, max(0, y) / 2000)
    const float near_ground = fmaf(0.015f, fabsf(vy), fmaf(0.01f, fabsf(ang), 0.05f * fabsf(om)));
    return -fmaf(w, near_ground, fmaf(0.005f, yp, 0.005f * fabsf(vx)));
}

// Landing step (rare): evaluate_landing (utils.py:1-37) + landing reward (reward.py:77-100).
template <class PP>
__device__ __forceinline__ float landing_reward(const PP& P, float avx, float avy, float aa, int& cls) {
    cls = 4;
    if (avx < P.land_lt(2, 0) && avy < P.land_lt(2, 1) && aa < P.land_lt(2, 2)) cls = 3;
    if (avx < P.land_lt(1, 0) && avy < P.land_lt(1, 1) && aa < P.land_lt(1, 2)) cls = 2;
    if (avx < P.land_lt(0, 0) && avy < P.land_lt(0, 1) && aa < P.land_lt(0, 2)) cls = 1;
    const float q = 0.6f * fmaxf(0.0f, 1.0f - aa * 0.1f) + 0.4f * fmaxf(0.0f, 1.0f - avy * 0.2f);
    if (cls == 1) return P.r_perfect * (1.0f + 0.5f * q);
    if (cls == 2) return P.r_good * (0.8f + 0.2f * q);
    if (cls == 3) return P.r_ok;
    return P.r_crash * (0.7f + 0.3f * fminf(1.0f, (avy * 0.05f + aa * (1.0f / 45.0f)) * 0.5f));
}

// Unpack the meta word and derive the before-state of the next step from the persistent words.
template <class PP>
__device__ __forceinline__ void flight_begin(const PP& P, const Rocket& r, Flight& f) {
    f.fresh = (r.meta & RL_META_FRESH) != 0u;
    f.steps = (int)(r.meta & RL_META_STEP_MASK);
    f.k = ((int)(r.meta << 5)) >> 29;                          // bits [24,27) sign-extended
    f.lo = ((int)(r.meta << 8)) >> 24;                         // bits [16,24) sign-extended
    f.sim_bits = r.meta & RL_META_SIM_MASK;
    f.x = (float)r.x * P.pos_unit;                             // (the unit scalings are exact: powers of two)
    f.y = (float)r.y * P.pos_unit;
    f.vx = ((float)r.sx * P.delta_unit) * P.inv_dt;
    f.vy = ((float)r.sy * P.delta_unit) * P.inv_dt;
    f.om = r.sang * P.inv_dt;
    f.ang = angle_deg(r.ang);
    f.phi = potential(f.y, f.vx, f.vy, f.ang, f.om);
}

__device__ __forceinline__ void flight_end(Rocket& r, const Flight& f) {
    r.meta = (f.fresh ? RL_META_FRESH : 0u) | f.sim_bits | (((uint32_t)f.k & 7u) << RL_META_WRAP_SHIFT) |
             (((uint32_t)f.lo & 0xffu) << RL_META_DLO_SHIFT) | (uint32_t)f.steps;
}

// One env step (lander.py:188-255) of one rocket, no reset.  th_raw / cg_raw are the raw
// action: physics clips it (base.py:135-136), the reward does not (reward.py:69-70).
//
// rocket_advance does everything except the observation and the meta word: `f` describes the state
// before the step on entry and the state after it on return (ready for the next fused substep).
// kClampPos (simulation mode only): that mode has no bounds check (controls.py:89 ignores truncation),
// so a rocket can reach the edge of the fixed-point position grid; it then stays there instead of
// wrapping around.  The gym step does not need the check: reward.py:274-276 ends the episode at
// 1 / 1.3 of the grid.
template <class PP, bool kClampPos = false>
__device__ __forceinline__ void rocket_advance(const PP& P, Rocket& r, Flight& f, float th_raw, float cg_raw,
                                               StepResult& o) {
    const bool fresh = f.fresh;
    const int steps = min(f.steps + 1, (int)RL_META_STEP_MASK);                 // lander.py:229
    const int k_prev = f.k, lo_prev = f.lo;

    // ---- state before the action (Rocket.get_state, base.py:220) --------------------------
    const float y_b = f.y, ang_b = f.ang;
    const float vx_b = f.vx, vy_b = f.vy, om_b = f.om;

    // ---- Rocket.apply_action (base.py:132-184) -----------------------------------------------
    const int32_t fuel_now = r.fuel;
    const float th = (fuel_now <= 0) ? 0.0f : clipf(th_raw, 0.0f, 1.0f);          // base.py:135-141
    const float cg = clipf(cg_raw, -1.0f, 1.0f);
    const double m_d = fma((double)fuel_now, P.fuel_unit_d, (double)r.dry);       // pre-burn mass (exact), base.py:143-144

    float ax = 0.0f, ay = 0.0f, alpha = 0.0f, vx_a = vx_b, vy_a = vy_b, om_a = om_b;
    float speed_chk = 0.0f;
    if (m_d > 1e-6) {                                          // base.py:145-147: else no-op
        const double inv_m_d = rl_rcp_d(m_d);
        const float inv_m = (float)inv_m_d;
        // drag (engine.py:78-98): 0.5 rho Cd A v^2 (-v/|v|) = -drag_k |v| v
        const float v2 = fmaf(vx_b, vx_b, vy_b * vy_b);
        const float kd = (v2 > P.speed2_gate) ? P.drag_k * rl_sqrt(v2) * inv_m : 0.0f;
        // zero-control acceleration: gravity + drag (engine.py:48-52, :117-123)
        const float a0x = -kd * vx_b;
        const float a0y = fmaf(-kd, vy_b, P.gravity);
        // thrust along the rocket axis, only above the throttle gate (engine.py:63-74)
        float s, c;
        sincos_units(r.ang, s, c);
        const float t = (th > P.throttle_gate) ? th * P.thrust_power * inv_m : 0.0f;
        ax = fmaf(t, s, a0x);
        ay = fmaf(t, c, a0y);
        // cold-gas torque (engine.py:125-155): alpha dt^2 in fp64 -- it accumulates into the angle delta
        const bool torque = (P.alpha_min_mass_d <= 1e-6) || (m_d >= P.alpha_min_mass_d);   // (folds away: m_d > 1e-6 here)
        const double adt2_d = torque ? (double)cg * P.alpha_dt2_d * inv_m_d : 0.0;
        alpha = torque ? cg * P.alpha_k * inv_m : 0.0f;        // (telemetry only)

        // carried deltas (x - x_prev, y - y_prev in fixed point; angle - angle_prev in fp64)
        int32_t dx = r.sx, dy = r.sy;
        double dang_d = dtheta_join(r.sang, lo_prev);
        if (k_prev != 0) dang_d = rl_unwrap(dang_d, k_prev);  // both angles normalised (base.py:171); rare: a real branch
        if (fresh) {
            // Verlet bootstrap (base.py:65-130): x - x_prev = v dt - a0 dt^2 / 2 (the slots hold v dt)
            // and, alpha0 being 0, angle - angle_prev = omega dt with the previous angle normalised
            // (base.py:117), which only matters within one step of +-180
            dx = rl_wrap_sub(dx, rl_f2i(a0x * P.half_dt2_units));
            dy = rl_wrap_sub(dy, rl_f2i(a0y * P.half_dt2_units));
            if (!(fabsf(ang_b) + fabsf(r.sang) < 179.0f)) {
                const double prev = (double)r.ang * kAngleUnit - dang_d;
                dang_d += 360.0 * floor((prev + 180.0) * (1.0 / 360.0));
            }
        }
        // engine.py:181-222
        const int32_t ndx = rl_wrap_add(dx, rl_f2i(ax * P.dt2_units));        // (two's-complement wrap is well defined
        const int32_t ndy = rl_wrap_add(dy, rl_f2i(ay * P.dt2_units));        // here; speed_chk below reports it)
        const double ndang_d = fma(dang_d, P.damp_d, adt2_d);
        const float sxf = (float)ndx * P.delta_unit, syf = (float)ndy * P.delta_unit;
        // x' = x + dx', rounded to the position grid (ties up).  No overflow check in the gym step: the grid
        // spans 1.3 x the bounds at which reward.py:274-276 ends the episode (kClampPos: the simulation mode)
        const int32_t x0 = r.x, y0 = r.y;
        r.x = rl_wrap_add(x0, rl_wrap_add(ndx, 1 << (kDeltaShift - 1)) >> kDeltaShift);
        r.y = rl_wrap_add(y0, rl_wrap_add(ndy, 1 << (kDeltaShift - 1)) >> kDeltaShift);
        if constexpr (kClampPos) {                             // sign flip far from zero = wrap-around
            if (((x0 ^ r.x) < 0) && (x0 > 0x40000000 || x0 < -0x40000000)) r.x = x0 > 0 ? INT32_MAX : INT32_MIN;
            if (((y0 ^ r.y) < 0) && (y0 > 0x40000000 || y0 < -0x40000000)) r.y = y0 > 0 ? INT32_MAX : INT32_MIN;
        }
        vx_a = sxf * P.inv_dt;
        vy_a = syf * P.inv_dt;
        speed_chk = fmaxf(fabsf(sxf), fabsf(syf));
        float hi;
        int lo;
        dtheta_split(ndang_d, hi, lo);
        om_a = hi * P.inv_dt;                                  // before normalisation, engine.py:220-222
        // angle + delta, normalised (engine.py:226): integer wrap-around; k = turns removed
        const double du_d = ndang_d * kAngleInvUnit;
        int k = 0;
        if (fabsf(hi) < 90.0f) {
            const int32_t du = rl_d2i(du_d);
            const int32_t na = (int32_t)((uint32_t)r.ang + (uint32_t)du);
            if ((~(r.ang ^ du) & (r.ang ^ na)) < 0) k = du > 0 ? 1 : -1;      // signed overflow: wrapped once
            r.ang = na;
        } else {                                               // >= 900 deg/s: rare, general path
            const long long sum = (long long)r.ang + rl_d2ll(du_d);
            const int32_t na = (int32_t)(uint32_t)(unsigned long long)sum;
            k = (int)((sum - (long long)na) >> 32);
            r.ang = na;
        }
        r.sx = ndx;
        r.sy = ndy;
        r.sang = hi;
        const int32_t burn = rl_f2i(th * P.burn_units);        // engine.py:239-243, base.py:174-175 (th >= 0)
        r.fuel = max(fuel_now - burn, 0);
        f.k = k;
        f.lo = lo;
        f.fresh = false;
    } else {
        r.fuel = max(fuel_now, 0);                             // base.py:139-141 happens before the early return
    }
    f.steps = steps;

    const float x_a = (float)r.x * P.pos_unit, y_a = (float)r.y * P.pos_unit, ang_a = angle_deg(r.ang);
    const float aa = fabsf(ang_a), avy = fabsf(vy_a);

    // ---- calculate_reward (reward.py:27-281), shaping path ------------------------------------
    const float ab = fabsf(ang_b), wb = fabsf(om_b), wa = fabsf(om_a);
    const float d_ang = ab - aa, d_om = wb - wa;
    // 1. angular control (reward.py:108-124)
    float total = fmaf(d_ang, 0.5f, d_om * 0.1f);
    // cold-gas direction term (reward.py:130-167); cold gas is the RAW action
    const bool need = (ab > P.dead_band) || (wb > P.dead_band);
    // (angle > 0 and cg < 0) or (angle < 0 and cg > 0): both non-zero with opposite signs, i.e. a
    // strictly negative product (|angle| >= 8.4e-8 where non-zero, so only a cold-gas command below
    // 1e-38 could underflow it -- where the term it gates is below 1e-36 anyway)
    const bool correct = ang_b * cg_raw < 0.0f;
    const float effect = d_ang + d_om;
    const float amp = (need && effect > 0.0f) ? fmaf(effect, effect, 1.0f) : 1.0f;
    const float gain = need ? (ab + wb) * (correct ? P.k_direction : -P.k_direction) : -0.3f;
    total = fmaf(fabsf(cg_raw) * gain * P.k_cold_gas, amp, total);
    const bool airborne = y_a > P.ground_y;
    const bool descending = airborne && (vy_a < 0.0f);
    // 2. throttle while descending (reward.py:171-186)
    const float af = rl_sat(fmaf(aa, -1.0f / 45.0f, 1.0f));              // max(0, 1 - |angle| / 45), aa >= 0
    const float t2 = th_raw * (-vy_a) * af * (1.0f + rl_sat(avy * 0.1f)) * P.k_throttle_descent;
    total += descending ? t2 : 0.0f;
    // 3. free fall (reward.py:190-201)
    const float prox = fmaf(8.0f, rl_rcp_fast(fmaxf(y_a, 1.0f)), 1.0f);
    const float t3 = P.k_free_fall * (-vy_a) * prox * (1.0f + rl_sat(avy * (1.0f / 15.0f)));
    total -= (descending && th_raw < P.throttle_lo_lt) ? t3 : 0.0f;
    // 4. throttle while tilted (reward.py:205-219)
    const float tilt = (aa > 10.0f) ? aa * (1.0f / 90.0f) : ((aa > 5.0f) ? (aa - 5.0f) * 0.1f : 0.0f);
    total -= (airborne && th_raw > P.throttle_lo_gt) ? th_raw * tilt * P.k_angle_throttle : 0.0f;
    // 5. climbing (reward.py:222-227)
    // (both factors only count when y_a > 0.1 and vy_a > 0, where min(1, x) == sat(x))
    const float t5 = vy_a * 0.5f * rl_sat(y_a * 1e-3f) * (1.0f + rl_sat(vy_a * 0.2f));
    total -= (airborne && vy_a > 0.0f) ? t5 : 0.0f;
    // 6. potential shaping (reward.py:267-270)
    const float phi_b = f.phi;
    const float phi_a = potential(y_a, vx_a, vy_a, ang_a, om_a);
    total += fmaf(P.gamma, phi_a, -phi_b);
    // truncation penalties (reward.py:274-279); tip-over never ends the episode
    bool oob = (fabsf(x_a) > P.max_x) || (y_a > P.max_y);
    bool tipped = aa > P.tip_angle;
    total += oob ? P.r_oob : 0.0f;
    total += tipped ? P.r_tip : 0.0f;

    // ---- ground contact (reward.py:73-102): replaces the shaping reward, skips bounds / tip ----
    const bool terminated = (y_a <= P.ground_y) && (y_b > P.ground_y);
    int landing = 0;
    if (terminated) {
        total = landing_reward(P, fabsf(vx_a), avy, aa, landing);
        oob = false;
        tipped = false;
    }
    o.terminated = terminated;
    o.out_of_bounds = oob;
    o.tipped = tipped;
    o.landing = landing;
    o.truncated = oob || (steps >= P.max_episode_steps);       // lander.py:237-241
    o.reward = total;
    o.ax = ax;
    o.ay = ay;
    o.alpha = alpha;
    o.vx = vx_a;
    o.vy = vy_a;
    o.om = om_a;
    o.phi = phi_a;
    f.x = x_a;
    f.y = y_a;
    f.vx = vx_a;
    f.vy = vy_a;
    f.om = om_a;
    f.ang = ang_a;
    f.phi = phi_a;
    // NaN / Inf, or a position delta about to leave the fixed-point range (+-2^31 delta units per step:
    // +-256 m for the reference's world)
    o.nonfinite = !(fabsf(total) <= 3.0e38f) || (speed_chk > P.delta_unit * 2.09e9f);
    r.ep_ret += total;
}

// _get_obs of the state rocket_advance left behind (lander.py:243)
template <class PP>
__device__ __forceinline__ void step_obs(const PP& P, const Rocket& r, const Flight& f, StepResult& o) {
    clip_obs(P, f.x, f.y, f.vx, f.vy, o.ax, o.ay, f.ang, f.om, o.obs);
}

template <class PP, bool kClampPos = false>
__device__ __forceinline__ void rocket_step(const PP& P, Rocket& r, float th_raw, float cg_raw,
                                            StepResult& o) {
    Flight f;
    flight_begin(P, r, f);
    rocket_advance<PP, kClampPos>(P, r, f, th_raw, cg_raw, o);
    flight_end(r, f);
    step_obs(P, r, f, o);
}

}  // namespace rl
