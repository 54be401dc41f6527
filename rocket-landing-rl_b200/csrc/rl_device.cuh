// Device-side rocket step: physics + reward/termination + observation + Philox reset.
//
// One rocket lives entirely in registers for the duration of a launch.  Arithmetic is fp32;
// the integrator is the reference's position-Verlet rewritten in DELTA FORM so that fp32
// loses nothing the reference's fp64 keeps (DESIGN.md "fp32 formulation"):
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
// The always-executed path is written with selects instead of branches (the kernel is issue- as
// much as bandwidth-limited, and BSSY/BSYNC/BRA around short divergent regions cost more than
// the few flops they skip); only genuinely rare events branch: touchdown, an angle leaving
// [-180, 180), a zero-mass rocket.
#pragma once
#ifndef RL_HOST_EMULATION
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "rl_params.h"

namespace rl {

struct Rocket {                // the ten persistent words (RL_W_*)
    float x, y, ang;
    float sx, sy, sang;
    float dry, fuel;
    float ep_ret;
    uint32_t meta;
    // precise-altitude mode only (RL_W_YLO, RL_W_SYLO): low-order words of y and of its delta,
    // so that y = y + y_lo and sy = sy + sy_lo carry ~48 bits; zero and unused otherwise
    float y_lo = 0.0f, sy_lo = 0.0f;
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
#else
inline float rl_rcp(float x) { return 1.0f / x; }
inline float rl_sqrt(float x) { return std::sqrt(x); }
inline float rl_rcp_fast(float x) { return 1.0f / x; }
inline float rl_sat(float x) { return x != x ? 0.0f : std::fmin(std::fmax(x, 0.0f), 1.0f); }
inline uint32_t rl_f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float rl_u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
#endif

// sin and cos of an angle given in DEGREES (the reference does np.sin(np.deg2rad(a)),
// engine.py:64-73).  Quadrant reduction is exact in degrees (a - 90 q), then degree-7 / degree-8
// minimax polynomials on [-pi/4, pi/4]; max abs error 8.7e-8 over [-180, 180) (measured against
// fp64), i.e. that of sincospif, in ~24 branch-free instructions.
__device__ __forceinline__ void sincos_deg(float a, float& s, float& c) {
    const float q = rintf(a * (1.0f / 90.0f));
    const float r = fmaf(q, -90.0f, a);
    const float t = r * 0.017453292519943295f;
    const float u = t * t;
    const float sn = fmaf(t * u, fmaf(fmaf(-0.0001958794891834259f, u, 0.008332748897373676f), u, -0.166666641831398f), t);
    const float cs = fmaf(u, fmaf(fmaf(fmaf(2.446382313792128e-05f, u, -0.001388759003020823f), u, 0.04166664928197861f), u, -0.5f), 1.0f);
    const int qi = (int)q;
    const bool swap = (qi & 1) != 0;
    const float ss = swap ? cs : sn, cc = swap ? sn : cs;
    s = rl_u2f(rl_f2u(ss) ^ (((uint32_t)qi & 2u) << 30));             // quadrants 2, 3: sin flips sign
    c = rl_u2f(rl_f2u(cc) ^ ((((uint32_t)qi + 1u) & 2u) << 30));      // quadrants 1, 2: cos flips sign
}

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

// Angle normalisation to [-180, 180) (engine.py:230-237).  In-range values pass through
// bit-untouched (the reference's `% 360` round trip perturbs them by ~1e-14, below fp32
// resolution); k is the number of 360-degree turns removed.
__device__ __forceinline__ float normalize_angle(float a, int& k) {
    k = 0;
    if (!(a >= -180.0f && a < 180.0f)) {          // rare path (also taken for NaN, harmlessly)
        const float turns = floorf((a + 180.0f) * (1.0f / 360.0f));
        a = fmaf(-360.0f, turns, a);
        k = (int)turns;
        if (a >= 180.0f) { a -= 360.0f; ++k; }
        if (a < -180.0f) { a += 360.0f; --k; }
    }
    return a;
}

// get_initial_state (simulation/config.py:26-53): ten independent uniforms lo + span u, in the
// reference's draw order x, y, vx, vy, ax, ay, angle, angular velocity, dry mass, fuel mass, from
// the ten random words w[0..9].  Returns the sampled ax, ay (they only ever appear in the reset
// observation, lander.py:182).
template <class PP>
__device__ __forceinline__ void rocket_from_words(const PP& P, const uint32_t* w, Rocket& r,
                                                  float& ax0, float& ay0) {
    r.x = fmaf(P.reset_span(0), u01(w[0]), P.reset_lo(0));
    r.y = fmaf(P.reset_span(1), u01(w[1]), P.reset_lo(1));
    r.sx = fmaf(P.reset_span(2), u01(w[2]), P.reset_lo(2));      // vx while FRESH
    r.sy = fmaf(P.reset_span(3), u01(w[3]), P.reset_lo(3));      // vy while FRESH
    ax0 = fmaf(P.reset_span(4), u01(w[4]), P.reset_lo(4));
    ay0 = fmaf(P.reset_span(5), u01(w[5]), P.reset_lo(5));
    r.ang = fmaf(P.reset_span(6), u01(w[6]), P.reset_lo(6));
    r.sang = fmaf(P.reset_span(7), u01(w[7]), P.reset_lo(7));    // angular velocity while FRESH
    r.dry = fmaf(P.reset_span(8), u01(w[8]), P.reset_lo(8));
    r.fuel = fmaf(P.reset_span(9), u01(w[9]), P.reset_lo(9));
    r.ep_ret = 0.0f;
    r.meta = RL_META_FRESH;                                      // current_step = 0 (lander.py:180)
}

// One thread draws all three Philox blocks keyed by (seed, global rocket id, epoch).
template <class PP>
__device__ __forceinline__ void sample_reset(const PP& P, uint64_t seed, uint64_t gid,
                                             uint64_t epoch, Rocket& r, float& ax0, float& ay0) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint4 a = philox4x32_10(reset_counter(gid, epoch, 0u), key);
    const uint4 b = philox4x32_10(reset_counter(gid, epoch, 1u), key);
    const uint4 c = philox4x32_10(reset_counter(gid, epoch, 2u), key);
    const uint32_t w[10] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y};
    rocket_from_words(P, w, r, ax0, ay0);
}

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

// Observation of a FRESH rocket (reset observation).
template <class PP>
__device__ __forceinline__ void fresh_obs(const PP& P, const Rocket& r, float ax0, float ay0,
                                          float* obs) {
    clip_obs(P, r.x, r.y, r.sx, r.sy, ax0, ay0, r.ang, r.sang, obs);
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

// One env step (lander.py:188-255) of one rocket, no reset.  th_raw / cg_raw are the raw
// action: physics clips it (base.py:135-136), the reward does not (reward.py:69-70).
//
// kPreciseY: the altitude recurrence (dy' = dy + ay dt^2, y' = y + dy') and the touchdown test run
// in fp64 on (hi, lo) float pairs; everything else, the forces included, stays fp32.  Altitude is
// the one word whose error decides a discrete outcome (the touchdown step): in fp32 it reaches
// ~1e-3 m by the end of a descent from 2000 m, i.e. a +-1-step disagreement in ~1 of 10^4 episodes.
//
// rocket_advance does everything except the observation.  With have_phi, o.phi must hold the
// potential of the CURRENT state (it does after a previous rocket_advance on the same rocket):
// the fused-substep loop then evaluates the potential once per substep instead of twice, with
// bit-identical results (same expression, same operands).
template <class PP, bool kPreciseY = false>
__device__ __forceinline__ void rocket_advance(const PP& P, Rocket& r, float th_raw, float cg_raw,
                                               StepResult& o, bool have_phi) {
    const bool fresh = (r.meta & RL_META_FRESH) != 0u;
    const int steps = min((int)(r.meta & RL_META_STEP_MASK) + 1, (int)RL_META_STEP_MASK);   // lander.py:229
    const int k_prev = ((int)(r.meta << 5)) >> 29;             // bits [24,27) sign-extended

    // ---- state before the action (Rocket.get_state, base.py:220) --------------------------
    const float y_b = r.y, ang_b = r.ang;
    const double y_b_d = (double)r.y + (double)r.y_lo;          // (precise mode only)
    const float vscale = fresh ? 1.0f : P.inv_dt;               // slots hold v while FRESH, deltas after
    const float vx_b = r.sx * vscale, om_b = r.sang * vscale;
    float vy_b = r.sy * vscale;
    if constexpr (kPreciseY) {
        if (!fresh) vy_b = (float)(((double)r.sy + (double)r.sy_lo) * (1.0 / (double)P.dt_d));
    }

    // ---- Rocket.apply_action (base.py:132-184) -----------------------------------------------
    const float fuel_now = r.fuel;
    const float th = (fuel_now <= 0.0f) ? 0.0f : clipf(th_raw, 0.0f, 1.0f);       // base.py:135-141
    const float cg = clipf(cg_raw, -1.0f, 1.0f);
    const float m = r.dry + fuel_now;                          // pre-burn mass, base.py:143-144

    float ax = 0.0f, ay = 0.0f, alpha = 0.0f, vx_a = vx_b, vy_a = vy_b, om_a = om_b;
    double y_a_d = y_b_d;
    if (m > P.mass_gate) {                                     // base.py:145-147: else no-op
        const float inv_m = rl_rcp(m);
        // drag (engine.py:78-98): 0.5 rho Cd A v^2 (-v/|v|) = -drag_k |v| v
        const float v2 = fmaf(vx_b, vx_b, vy_b * vy_b);
        const float kd = (v2 > P.speed2_gate) ? P.drag_k * rl_sqrt(v2) * inv_m : 0.0f;
        // zero-control acceleration: gravity + drag (engine.py:48-52, :117-123)
        const float a0x = -kd * vx_b;
        const float a0y = fmaf(-kd, vy_b, P.gravity);
        // thrust along the rocket axis, only above the throttle gate (engine.py:63-74)
        float s, c;
        sincos_deg(ang_b, s, c);
        const float t = (th > P.throttle_gate) ? th * P.thrust_power * inv_m : 0.0f;
        ax = fmaf(t, s, a0x);
        ay = fmaf(t, c, a0y);
        alpha = (m >= P.alpha_min_mass) ? cg * P.alpha_k * inv_m : 0.0f;              // engine.py:125-155

        // carried deltas; a FRESH rocket gets the Verlet bootstrap (base.py:65-130):
        // x - x_prev = v dt - a0 dt^2 / 2 and, alpha0 being 0, angle - angle_prev = omega dt
        const float boot_x = fmaf(-P.half_dt2, a0x, vx_b * P.dt);
        const float boot_y = fmaf(-P.half_dt2, a0y, vy_b * P.dt);
        const float back = om_b * P.dt;
        int kp = 0;
        if (fresh) (void)normalize_angle(ang_b - back, kp);    // base.py:117 normalises the previous angle
        const float dx = fresh ? boot_x : r.sx;
        const float dy = fresh ? boot_y : r.sy;
        // (angle - previous angle) with both angles normalised (base.py:171)
        const float dang = fresh ? fmaf(360.0f, (float)kp, back) : fmaf(-360.0f, (float)k_prev, r.sang);
        // engine.py:181-222
        const float ndx = fmaf(ax, P.dt2, dx);
        const float ndy = fmaf(ay, P.dt2, dy);
        const float ndang = fmaf(dang, P.damp, alpha * P.dt2);
        r.x += ndx;
        vx_a = ndx * P.inv_dt;
        if constexpr (kPreciseY) {
            const double dtd = (double)P.dt_d;
            const double dy_d = fresh ? ((double)vy_b * dtd - 0.5 * (double)a0y * dtd * dtd)
                                      : ((double)r.sy + (double)r.sy_lo);
            const double ndy_d = dy_d + (double)ay * (dtd * dtd);
            y_a_d = y_b_d + ndy_d;
            r.y = (float)y_a_d;
            r.y_lo = (float)(y_a_d - (double)r.y);
            r.sy = (float)ndy_d;
            r.sy_lo = (float)(ndy_d - (double)r.sy);
            vy_a = (float)(ndy_d * (1.0 / dtd));
        } else {
            r.y += ndy;
            vy_a = ndy * P.inv_dt;
            r.sy = ndy;
        }
        om_a = ndang * P.inv_dt;                               // before normalisation, engine.py:220-222
        int k;
        r.ang = normalize_angle(ang_b + ndang, k);             // engine.py:226
        r.sx = ndx;
        r.sang = ndang;
        r.fuel = fmaxf(0.0f, fuel_now - fmaxf(0.0f, th * P.burn_per_step));   // engine.py:239-243, base.py:174-175
        r.meta = (((uint32_t)k & 7u) << RL_META_WRAP_SHIFT) | (uint32_t)steps;   // FRESH cleared
    } else {
        r.fuel = fmaxf(0.0f, fuel_now);
        r.meta = (r.meta & ~RL_META_STEP_MASK) | (uint32_t)steps;
    }

    const float x_a = r.x, y_a = r.y, ang_a = r.ang;
    const float aa = fabsf(ang_a), avy = fabsf(vy_a);

    // ---- calculate_reward (reward.py:27-281), shaping path ------------------------------------
    const float ab = fabsf(ang_b), wb = fabsf(om_b), wa = fabsf(om_a);
    const float d_ang = ab - aa, d_om = wb - wa;
    // 1. angular control (reward.py:108-124)
    float total = fmaf(d_ang, 0.5f, d_om * 0.1f);
    // cold-gas direction term (reward.py:130-167); cold gas is the RAW action
    const bool need = (ab > P.dead_band) || (wb > P.dead_band);
    // (angle > 0 and cg < 0) or (angle < 0 and cg > 0): both non-zero with opposite signs
    const bool correct = (ang_b != 0.0f) && (cg_raw != 0.0f) && ((int32_t)(rl_f2u(ang_b) ^ rl_f2u(cg_raw)) < 0);
    const float effect = d_ang + d_om;
    const float amp = (need && effect > 0.0f) ? fmaf(effect, effect, 1.0f) : 1.0f;
    const float gain = need ? (ab + wb) * (correct ? P.k_direction : -P.k_direction) : -0.3f;
    total = fmaf(fabsf(cg_raw) * gain * P.k_cold_gas, amp, total);
    const bool airborne = kPreciseY ? (y_a_d > 0.1) : (y_a > P.ground_y);
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
    const float phi_b = have_phi ? o.phi : potential(y_b, vx_b, vy_b, ang_b, om_b);
    const float phi_a = potential(y_a, vx_a, vy_a, ang_a, om_a);
    total += fmaf(P.gamma, phi_a, -phi_b);
    // truncation penalties (reward.py:274-279); tip-over never ends the episode
    bool oob = (fabsf(x_a) > P.max_x) || (y_a > P.max_y);
    bool tipped = aa > P.tip_angle;
    total += oob ? P.r_oob : 0.0f;
    total += tipped ? P.r_tip : 0.0f;

    // ---- ground contact (reward.py:73-102): replaces the shaping reward, skips bounds / tip ----
    const bool terminated = kPreciseY ? ((y_a_d <= 0.1) && (y_b_d > 0.1))
                                      : ((y_a <= P.ground_y) && (y_b > P.ground_y));
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
    o.nonfinite = !(fabsf(total) <= 3.0e38f) || !(fabsf(y_a) <= 3.0e38f);
    r.ep_ret += total;
}

// _get_obs of the state rocket_advance left behind (lander.py:243)
template <class PP>
__device__ __forceinline__ void step_obs(const PP& P, const Rocket& r, StepResult& o) {
    clip_obs(P, r.x, r.y, o.vx, o.vy, o.ax, o.ay, r.ang, o.om, o.obs);
}

template <class PP, bool kPreciseY = false>
__device__ __forceinline__ void rocket_step(const PP& P, Rocket& r, float th_raw, float cg_raw,
                                            StepResult& o) {
    rocket_advance<PP, kPreciseY>(P, r, th_raw, cg_raw, o, false);
    step_obs(P, r, o);
}

}  // namespace rl
