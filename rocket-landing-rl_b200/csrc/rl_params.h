// Kernel-parameter struct of the environment step and its derivation from RlParams.
// Pure C++ (no CUDA): shared by the library (rl_capi.cu) and by the host-compiled logic check
// of the device math under tests/hostemu/.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <stdint.h>

#include "../../include/rocket_landing_b200.h"

namespace rl {

// Constants in fp32, derived on the host in double (rl_capi.cu: derive_params).  Thresholds
// are rounded in the direction that keeps the comparison exact for fp32 inputs: for a
// double threshold c, `x > c` == `x > down(c)` and `x < c` == `x < up(c)` for every fp32 x.
struct DevParams {
    float dt, inv_dt, dt2, half_dt2;
    float gravity;
    float thrust_power;        // N at throttle 1
    float drag_k;              // 0.5 rho Cd A   (engine.py:88-94); drag force = -drag_k |v| v
    float alpha_k;             // rad2deg(P_cg arm / (0.5 r^2)); alpha = cg alpha_k / m (engine.py:140-153)
    float alpha_min_mass;      // below this mass the inertia gate (engine.py:142) zeroes alpha
    float damp;                // max(0, 1 - angular_damping dt)  (engine.py:206)
    float burn_per_step;       // fuel_consumption_rate dt        (engine.py:242)
    float mass_gate;           // 1e-6 rounded down: m <= gate skips the step (base.py:145)
    float throttle_gate;       // 1e-6 rounded down: thrust only if throttle > gate (engine.py:63)
    float speed2_gate;         // 1e-9 rounded down: drag only if v.v > gate (engine.py:85)
    float ground_y;            // 0.1 rounded down: y <= 0.1 <=> y <= ground_y (reward.py:73)
    float dead_band;           // 0.1 rounded down: |angle| > 0.1 (reward.py:130)
    float throttle_lo_lt;      // 0.1 rounded up:   throttle < 0.1 (reward.py:190)
    float throttle_lo_gt;      // 0.1 rounded down: throttle > 0.1 (reward.py:205)
    float land_lt[3][3];       // [perfect|good|ok][vx|vy|angle], rounded up (strict <, utils.py:19-21)
    float max_x, max_y, tip_angle;            // rounded down (strict >, reward.py:274-278)
    float r_perfect, r_good, r_ok, r_crash, r_oob, r_tip;
    float k_throttle_descent, k_cold_gas, k_free_fall, k_angle_throttle, k_direction, gamma;
    float obs_lo[8], obs_hi[8];
    float reset_lo[10], reset_span[10];
    int max_episode_steps;
};

inline int derive_fail(char* err, size_t err_len, const char* fmt, ...) {
    if (err && err_len) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, err_len, fmt, ap);
        va_end(ap);
    }
    return RL_E_INVALID;
}

// largest fp32 <= c  /  smallest fp32 >= c: keeps `x > c` / `x < c` exact for fp32 x
inline float f32_down(double c) {
    float f = (float)c;
    if ((double)f > c) f = nextafterf(f, -INFINITY);
    return f;
}
inline float f32_up(double c) {
    float f = (float)c;
    if ((double)f < c) f = nextafterf(f, INFINITY);
    return f;
}

inline int derive_params(const RlParams& p, DevParams& d, char* err, size_t err_len) {
    if (!(p.time_step > 0)) return derive_fail(err, err_len, "time_step must be positive");          // engine.py:35
    if (p.thrust_power < 0 || p.cold_gas_thrust_power < 0 || p.fuel_consumption_rate < 0)
        return derive_fail(err, err_len, "thrust powers and fuel rate cannot be negative");          // engine.py:37-44
    if (!(p.radius > 0) || !(p.cold_gas_moment_arm > 0))
        return derive_fail(err, err_len, "radius and moment arm must be positive");                  // engine.py:45-46
    if (p.max_episode_steps < 1 || p.max_episode_steps > (int)RL_META_STEP_MASK)
        return derive_fail(err, err_len, "max_episode_steps out of range");
    const double pi = 3.14159265358979323846;
    const double dt = p.time_step;
    d.dt = (float)dt;
    d.inv_dt = (float)(1.0 / dt);
    d.dt2 = (float)(dt * dt);
    d.half_dt2 = (float)(0.5 * dt * dt);
    d.gravity = (float)p.gravity;
    d.thrust_power = (float)p.thrust_power;
    d.drag_k = (float)(0.5 * p.air_density * p.drag_coefficient * p.reference_area);
    const double half_r2 = 0.5 * p.radius * p.radius;
    d.alpha_k = (float)(p.cold_gas_thrust_power * p.cold_gas_moment_arm / half_r2 * (180.0 / pi));
    // engine.py:137-143: alpha = 0 when m <= 1e-6, radius <= 1e-6 or 0.5 m r^2 < 1e-6
    d.alpha_min_mass = (p.radius <= 1e-6) ? INFINITY : f32_up(1e-6 / half_r2);
    d.damp = (float)fmax(0.0, 1.0 - p.angular_damping * dt);
    d.burn_per_step = (float)(p.fuel_consumption_rate * dt);
    d.mass_gate = f32_down(1e-6);
    d.throttle_gate = f32_down(1e-6);
    d.speed2_gate = f32_down(1e-9);
    d.ground_y = f32_down(0.1);
    d.dead_band = f32_down(0.1);
    d.throttle_lo_lt = f32_up(0.1);
    d.throttle_lo_gt = f32_down(0.1);
    for (int k = 0; k < 3; ++k) {
        d.land_lt[0][k] = f32_up(p.land_perfect[k]);
        d.land_lt[1][k] = f32_up(p.land_good[k]);
        d.land_lt[2][k] = f32_up(p.land_ok[k]);
    }
    d.max_x = f32_down(p.max_horizontal_position);
    d.max_y = f32_down(p.max_altitude);
    d.tip_angle = f32_down(p.tip_over_angle);
    d.r_perfect = (float)p.landing_perfect;
    d.r_good = (float)p.landing_good;
    d.r_ok = (float)p.landing_ok;
    d.r_crash = (float)p.crash_ground;
    d.r_oob = (float)p.out_of_bounds;
    d.r_tip = (float)p.tipped_over;
    d.k_throttle_descent = (float)p.throttle_descent_reward_scale;
    d.k_cold_gas = (float)p.cold_gas_reward_scale;
    d.k_free_fall = (float)p.free_fall_penalty_scale;
    d.k_angle_throttle = (float)p.angle_aware_throttle_scale;
    d.k_direction = (float)p.correct_direction_bonus;
    d.gamma = (float)p.gamma;
    for (int k = 0; k < 8; ++k) {
        d.obs_lo[k] = (float)p.obs_low[k];       // np.array(..., dtype=float32), lander.py:73-101
        d.obs_hi[k] = (float)p.obs_high[k];
        if (!(d.obs_lo[k] <= d.obs_hi[k])) return derive_fail(err, err_len, "obs_low[%d] > obs_high[%d]", k, k);
    }
    for (int k = 0; k < 10; ++k) {
        if (!(p.reset_low[k] <= p.reset_high[k])) return derive_fail(err, err_len, "reset_low[%d] > reset_high[%d]", k, k);
        d.reset_lo[k] = (float)p.reset_low[k];
        d.reset_span[k] = (float)(p.reset_high[k] - p.reset_low[k]);
    }
    d.max_episode_steps = p.max_episode_steps;
    return RL_OK;
}

}  // namespace rl
