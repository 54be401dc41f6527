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
};

struct StepResult {
    float obs[8];
    float reward;
    bool terminated, truncated, out_of_bounds, tipped, nonfinite;
    int landing;               // 0 none, 1 safe, 2 good, 3 ok, 4 unsafe (protocol.py:31-41)
};

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
// reference's draw order, from three Philox blocks keyed by (seed, global rocket id, epoch).
// Returns the sampled ax, ay (they only ever appear in the reset observation, lander.py:182).
__device__ __forceinline__ void sample_reset(const DevParams& P, uint64_t seed, uint64_t gid,
                                             uint64_t epoch, Rocket& r, float& ax0, float& ay0) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t c0 = (uint32_t)gid, c1 = (uint32_t)(gid >> 32), c2 = (uint32_t)epoch;
    const uint32_t c3 = (uint32_t)(epoch >> 32) << 2;
    const uint4 a = philox4x32_10(make_uint4(c0, c1, c2, c3 | 0u), key);
    const uint4 b = philox4x32_10(make_uint4(c0, c1, c2, c3 | 1u), key);
    const uint4 c = philox4x32_10(make_uint4(c0, c1, c2, c3 | 2u), key);
    r.x = fmaf(P.reset_span[0], u01(a.x), P.reset_lo[0]);
    r.y = fmaf(P.reset_span[1], u01(a.y), P.reset_lo[1]);
    r.sx = fmaf(P.reset_span[2], u01(a.z), P.reset_lo[2]);      // vx while FRESH
    r.sy = fmaf(P.reset_span[3], u01(a.w), P.reset_lo[3]);      // vy while FRESH
    ax0 = fmaf(P.reset_span[4], u01(b.x), P.reset_lo[4]);
    ay0 = fmaf(P.reset_span[5], u01(b.y), P.reset_lo[5]);
    r.ang = fmaf(P.reset_span[6], u01(b.z), P.reset_lo[6]);
    r.sang = fmaf(P.reset_span[7], u01(b.w), P.reset_lo[7]);    // angular velocity while FRESH
    r.dry = fmaf(P.reset_span[8], u01(c.x), P.reset_lo[8]);
    r.fuel = fmaf(P.reset_span[9], u01(c.y), P.reset_lo[9]);
    r.ep_ret = 0.0f;
    r.meta = RL_META_FRESH;                                      // current_step = 0 (lander.py:180)
}

__device__ __forceinline__ float clipf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// _get_obs (lander.py:124-150)
__device__ __forceinline__ void clip_obs(const DevParams& P, float x, float y, float vx, float vy,
                                         float ax, float ay, float ang, float om, float* obs) {
    obs[0] = clipf(x, P.obs_lo[0], P.obs_hi[0]);
    obs[1] = clipf(y, P.obs_lo[1], P.obs_hi[1]);
    obs[2] = clipf(vx, P.obs_lo[2], P.obs_hi[2]);
    obs[3] = clipf(vy, P.obs_lo[3], P.obs_hi[3]);
    obs[4] = clipf(ax, P.obs_lo[4], P.obs_hi[4]);
    obs[5] = clipf(ay, P.obs_lo[5], P.obs_hi[5]);
    obs[6] = clipf(ang, P.obs_lo[6], P.obs_hi[6]);
    obs[7] = clipf(om, P.obs_lo[7], P.obs_hi[7]);
}

// Observation of a FRESH rocket (reset observation).
__device__ __forceinline__ void fresh_obs(const DevParams& P, const Rocket& r, float ax0, float ay0,
                                          float* obs) {
    clip_obs(P, r.x, r.y, r.sx, r.sy, ax0, ay0, r.ang, r.sang, obs);
}

// Potential of reward.py:231-264.
__device__ __forceinline__ float potential(float y, float vx, float vy, float ang, float om) {
    const float yp = fmaxf(0.0f, y);
    const float w = 2.0f - fminf(1.0f, yp * (1.0f / 2000.0f));
    return -0.005f * yp - (0.015f * w) * fabsf(vy) - 0.005f * fabsf(vx) - (0.01f * w) * fabsf(ang) -
           (0.05f * w) * fabsf(om);
}

// One env step (lander.py:188-255) of one rocket, no reset.  th_raw / cg_raw are the raw
// action: physics clips it (base.py:135-136), the reward does not (reward.py:69-70).
__device__ __forceinline__ void rocket_step(const DevParams& P, Rocket& r, float th_raw, float cg_raw,
                                            StepResult& o) {
    const bool fresh = (r.meta & RL_META_FRESH) != 0u;
    int steps = (int)(r.meta & RL_META_STEP_MASK);
    const int k_prev = ((int)(r.meta << 5)) >> 29;             // bits [24,27) sign-extended

    // ---- state before the action (Rocket.get_state, base.py:220) --------------------------
    const float y_b = r.y, ang_b = r.ang;
    const float vx_b = fresh ? r.sx : r.sx * P.inv_dt;
    const float vy_b = fresh ? r.sy : r.sy * P.inv_dt;
    const float om_b = fresh ? r.sang : r.sang * P.inv_dt;

    // ---- Rocket.apply_action (base.py:132-184) -----------------------------------------------
    float th = clipf(th_raw, 0.0f, 1.0f);
    const float cg = clipf(cg_raw, -1.0f, 1.0f);
    const float fuel_now = r.fuel;
    if (fuel_now <= 0.0f) th = 0.0f;                           // base.py:138-141
    const float m = r.dry + fuel_now;                          // pre-burn mass, base.py:143-144

    float ax = 0.0f, ay = 0.0f, vx_a = vx_b, vy_a = vy_b, om_a = om_b;
    if (m > P.mass_gate) {                                     // base.py:145-147: else no-op
        const float inv_m = 1.0f / m;
        // drag (engine.py:78-98): 0.5 rho Cd A v^2 (-v/|v|) = -drag_k |v| v
        const float v2 = fmaf(vx_b, vx_b, vy_b * vy_b);
        const float kd = (v2 > P.speed2_gate) ? P.drag_k * sqrtf(v2) * inv_m : 0.0f;
        // zero-control acceleration: gravity + drag (engine.py:48-52, :117-123)
        const float a0x = -kd * vx_b;
        const float a0y = fmaf(-kd, vy_b, P.gravity);
        ax = a0x;
        ay = a0y;
        if (th > P.throttle_gate) {                            // thrust, engine.py:63-74
            float s, c;
            sincospif(ang_b * (1.0f / 180.0f), &s, &c);
            const float t = th * P.thrust_power * inv_m;
            ax = fmaf(t, s, ax);
            ay = fmaf(t, c, ay);
        }
        const float alpha = (m >= P.alpha_min_mass) ? cg * P.alpha_k * inv_m : 0.0f;  // engine.py:125-155

        float dx, dy, dang;
        if (fresh) {
            // Verlet bootstrap (base.py:65-130): x - x_prev = v dt - a0 dt^2 / 2, alpha0 = 0
            dx = fmaf(-P.half_dt2, a0x, vx_b * P.dt);
            dy = fmaf(-P.half_dt2, a0y, vy_b * P.dt);
            const float back = om_b * P.dt;
            int kp;
            (void)normalize_angle(ang_b - back, kp);          // base.py:117: previous angle is normalised
            dang = fmaf(360.0f, (float)kp, back);
        } else {
            dx = r.sx;
            dy = r.sy;
            dang = fmaf(-360.0f, (float)k_prev, r.sang);       // (angle - previous angle), base.py:171
        }
        // engine.py:181-222
        const float ndx = fmaf(ax, P.dt2, dx);
        const float ndy = fmaf(ay, P.dt2, dy);
        const float ndang = fmaf(dang, P.damp, alpha * P.dt2);
        r.x += ndx;
        r.y += ndy;
        vx_a = ndx * P.inv_dt;
        vy_a = ndy * P.inv_dt;
        om_a = ndang * P.inv_dt;                               // before normalisation, engine.py:220-222
        int k;
        r.ang = normalize_angle(ang_b + ndang, k);             // engine.py:226
        r.sx = ndx;
        r.sy = ndy;
        r.sang = ndang;
        r.fuel = fmaxf(0.0f, fuel_now - fmaxf(0.0f, th * P.burn_per_step));   // engine.py:239-243, base.py:174-175
        r.meta = ((uint32_t)k & 7u) << RL_META_WRAP_SHIFT;     // FRESH cleared; step count set below
    } else {
        r.fuel = fmaxf(0.0f, fuel_now);
        r.meta &= ~RL_META_STEP_MASK;
    }
    steps = min(steps + 1, (int)RL_META_STEP_MASK);            // lander.py:229
    r.meta |= (uint32_t)steps;

    const float x_a = r.x, y_a = r.y, ang_a = r.ang;
    const float aa = fabsf(ang_a), avy = fabsf(vy_a);

    // ---- calculate_reward (reward.py:27-281) -----------------------------------------------
    float total = 0.0f;
    o.terminated = (y_a <= P.ground_y) && (y_b > P.ground_y);  // reward.py:73
    o.truncated = false;
    o.out_of_bounds = false;
    o.tipped = false;
    o.landing = 0;
    if (o.terminated) {
        // evaluate_landing (utils.py:1-37)
        const float avx = fabsf(vx_a);
        int cls = 4;
        if (avx < P.land_lt[2][0] && avy < P.land_lt[2][1] && aa < P.land_lt[2][2]) cls = 3;
        if (avx < P.land_lt[1][0] && avy < P.land_lt[1][1] && aa < P.land_lt[1][2]) cls = 2;
        if (avx < P.land_lt[0][0] && avy < P.land_lt[0][1] && aa < P.land_lt[0][2]) cls = 1;
        o.landing = cls;
        const float q = 0.6f * fmaxf(0.0f, 1.0f - aa * 0.1f) + 0.4f * fmaxf(0.0f, 1.0f - avy * 0.2f);
        if (cls == 1) total = P.r_perfect * (1.0f + 0.5f * q);
        else if (cls == 2) total = P.r_good * (0.8f + 0.2f * q);
        else if (cls == 3) total = P.r_ok;
        else total = P.r_crash * (0.7f + 0.3f * fminf(1.0f, (avy * 0.05f + aa * (1.0f / 45.0f)) * 0.5f));
        // reward.py:102 returns here: no bounds / tip-over check on the landing step
    } else {
        const float ab = fabsf(ang_b), wb = fabsf(om_b), wa = fabsf(om_a);
        // 1. angular control (reward.py:108-124)
        total = -(aa - ab) * 0.5f - (wa - wb) * 0.1f;
        // cold-gas direction term (reward.py:130-167); cold gas is the RAW action
        const float acg = fabsf(cg_raw);
        if (ab > P.dead_band || wb > P.dead_band) {
            const bool correct = (ang_b > 0.0f && cg_raw < 0.0f) || (ang_b < 0.0f && cg_raw > 0.0f);
            const float effect = (ab - aa) + (wb - wa);
            float cgr = acg * (ab + wb) * (correct ? P.k_direction : -P.k_direction) * P.k_cold_gas;
            if (effect > 0.0f) cgr *= fmaf(effect, effect, 1.0f);
            total += cgr;
        } else {
            total -= acg * 0.3f * P.k_cold_gas;
        }
        const bool airborne = y_a > P.ground_y;
        if (airborne && vy_a < 0.0f) {
            // 2. throttle while descending (reward.py:171-186)
            const float af = fmaxf(0.0f, 1.0f - aa * (1.0f / 45.0f));
            total += th_raw * (-vy_a) * af * (1.0f + fminf(1.0f, avy * 0.1f)) * P.k_throttle_descent;
            // 3. free fall (reward.py:190-201)
            if (th_raw < P.throttle_lo_lt) {
                const float prox = 1.0f + __fdividef(8.0f, fmaxf(y_a, 1.0f));
                total -= P.k_free_fall * (-vy_a) * prox * (1.0f + fminf(1.0f, avy * (1.0f / 15.0f)));
            }
        }
        // 4. throttle while tilted (reward.py:205-219)
        if (th_raw > P.throttle_lo_gt && airborne) {
            if (aa > 10.0f) total -= th_raw * (aa * (1.0f / 90.0f)) * P.k_angle_throttle;
            else if (aa > 5.0f) total -= th_raw * ((aa - 5.0f) * 0.2f) * P.k_angle_throttle * 0.5f;
        }
        // 5. climbing (reward.py:222-227)
        if (airborne && vy_a > 0.0f)
            total -= vy_a * 0.5f * fminf(1.0f, y_a * 1e-3f) * (1.0f + fminf(1.0f, vy_a * 0.2f));
        // 6. potential shaping (reward.py:267-270)
        total += fmaf(P.gamma, potential(y_a, vx_a, vy_a, ang_a, om_a), -potential(y_b, vx_b, vy_b, ang_b, om_b));
        // truncation penalties (reward.py:274-279); tip-over never ends the episode
        if (fabsf(x_a) > P.max_x || y_a > P.max_y) {
            total += P.r_oob;
            o.truncated = true;
            o.out_of_bounds = true;
        }
        if (aa > P.tip_angle) {
            total += P.r_tip;
            o.tipped = true;
        }
    }
    if (steps >= P.max_episode_steps) o.truncated = true;      // lander.py:237-241
    o.reward = total;
    o.nonfinite = !(fabsf(total) <= 3.0e38f) || !(fabsf(y_a) <= 3.0e38f);
    r.ep_ret += total;
    clip_obs(P, x_a, y_a, vx_a, vy_a, ax, ay, ang_a, om_a, o.obs);
}

}  // namespace rl
