// Kernel-parameter struct of the environment step and its derivation from RlParams.
// Pure C++ (no CUDA): shared by the library (rl_capi.cu) and by the host-compiled logic check
// of the device math under tests/hostemu/.
//
// Two providers of the same constants exist, with the same accessor surface:
//   * DevParams      -- runtime values (kernel argument, read through the constant bank);
//   * DefaultParams  -- the reference's config.yaml baked in as constexpr, so the compiler folds
//                       them into immediates (the step kernel is issue-limited as much as
//                       bandwidth-limited; ~50 of its ~450 instructions were constant loads).
// rl_create picks the constexpr instantiation only when EVERY derived constant is bit-identical
// to DefaultParams (matches_default); any other config runs the runtime instantiation of the
// very same templates.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <stdint.h>

#include "../../include/rocket_landing_b200.h"

#if defined(__CUDACC__)
#define RL_HD __host__ __device__
#else
#define RL_HD
#endif

namespace rl {

// Constants in fp32, derived on the host in double (derive_params).  Thresholds are rounded in
// the direction that keeps the comparison exact for fp32 inputs: for a double threshold c,
// `x > c` == `x > down(c)` and `x < c` == `x < up(c)` for every fp32 x.
// Units of the fixed-point state words (DESIGN.md "Long-horizon accuracy"): errors in the
// integrated words persist for the rest of an episode, so they carry 32 significant bits.
constexpr double kAngleUnit = 360.0 / 4294967296.0;      // degrees per unit of the int32 angle: [-180, 180) = int32
constexpr double kAngleInvUnit = 4294967296.0 / 360.0;
constexpr int kDeltaShift = 8;                           // a position delta counts units of pos_unit / 2^kDeltaShift

struct DevParams {
    float dt, inv_dt, dt2, half_dt2;
    // ---- extended-precision words (DESIGN.md "Long-horizon accuracy") ----------------------------
    double dt_d;               // time step in fp64 (Verlet bootstrap of a FRESH rocket)
    double damp_d;             // max(0, 1 - angular_damping dt) in fp64 (angle chain)
    double alpha_dt2_d;        // alpha_k dt^2 in fp64: the angle delta gains cg alpha_dt2_d / m per step
    double fuel_unit_d;        // kg per unit of the fixed-point fuel word (a power of two)
    double alpha_min_mass_d;   // fp64 copy of the inertia gate
    float pos_unit;            // metres per unit of the fixed-point positions x, y (a power of two)
    float pos_inv_unit;        // 1 / pos_unit (reset: metres -> units)
    float delta_unit;          // metres per unit of the position deltas = pos_unit / 2^kDeltaShift
    float dt2_units;           // dt^2 / delta_unit: acceleration -> fixed-point delta increment
    float half_dt2_units;      // dt^2 / 2 / delta_unit: Verlet bootstrap correction of a FRESH rocket
    float dt_units;            // dt / delta_unit: velocity -> fixed-point delta (reset)
    float burn_units;          // fuel_consumption_rate dt / fuel_unit: throttle -> fixed-point burn
    float fuel_inv_unit;       // 1 / fuel_unit (reset: kg -> units)
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
    float max_x, max_y, tip_angle;            // rounded down (strict >, reward.py:274-278)
    float r_perfect, r_good, r_ok, r_crash, r_oob, r_tip;
    float k_throttle_descent, k_cold_gas, k_free_fall, k_angle_throttle, k_direction, gamma;
    float land_lt_[3][3];      // [perfect|good|ok][vx|vy|angle], rounded up (strict <, utils.py:19-21)
    float obs_lo_[8], obs_hi_[8];
    float reset_lo_[10], reset_span_[10];
    int max_episode_steps;

    RL_HD float land_lt(int level, int k) const { return land_lt_[level][k]; }
    RL_HD float obs_lo(int k) const { return obs_lo_[k]; }
    RL_HD float obs_hi(int k) const { return obs_hi_[k]; }
    RL_HD float reset_lo(int k) const { return reset_lo_[k]; }
    RL_HD float reset_span(int k) const { return reset_span_[k]; }
};

// The reference's config.yaml (config.yaml:16-91), derived exactly as derive_params does.
struct DefaultParams {
    static constexpr double kDt = 0.1, kPi = 3.14159265358979323846;
    static constexpr float dt = (float)kDt;
    static constexpr double dt_d = kDt;
    static constexpr double damp_d = 1.0 - 0.05 * kDt;
    static constexpr double alpha_dt2_d = 7.0e4 * 1.85 / (0.5 * 1.85 * 1.85) * (180.0 / kPi) * (kDt * kDt);
    static constexpr double fuel_unit_d = 1.0 / 4096.0;               // 410 000 kg < 2^19 -> 2^(19-31)
    static constexpr double alpha_min_mass_d = 1e-6 / (0.5 * 1.85 * 1.85);
    static constexpr float pos_unit = (float)(1.0 / 32768.0);          // 1.3 x 50 000 m < 2^16 -> 2^(16-31)
    static constexpr float pos_inv_unit = 32768.0f;
    static constexpr float delta_unit = (float)(1.0 / 8388608.0);      // 2^-23 m
    static constexpr float dt2_units = (float)(kDt * kDt * 8388608.0);
    static constexpr float half_dt2_units = (float)(0.5 * kDt * kDt * 8388608.0);
    static constexpr float dt_units = (float)(kDt * 8388608.0);
    static constexpr float burn_units = (float)(1700.0 * kDt * 4096.0);
    static constexpr float fuel_inv_unit = 4096.0f;
    static constexpr float inv_dt = (float)(1.0 / kDt);
    static constexpr float dt2 = (float)(kDt * kDt);
    static constexpr float half_dt2 = (float)(0.5 * kDt * kDt);
    static constexpr float gravity = (float)-9.81;
    static constexpr float thrust_power = 1.0e7f;
    static constexpr float drag_k = (float)(0.5 * 1.225 * 0.8 * 10.8);
    static constexpr float alpha_k = (float)(7.0e4 * 1.85 / (0.5 * 1.85 * 1.85) * (180.0 / kPi));
    static constexpr float alpha_min_mass = 0x1.39baf4p-21f;          // up(1e-6 / (0.5 r^2))
    static constexpr float damp = (float)(1.0 - 0.05 * kDt);
    static constexpr float burn_per_step = (float)(1700.0 * kDt);
    static constexpr float mass_gate = 0x1.0c6f7ap-20f;               // down(1e-6)
    static constexpr float throttle_gate = 0x1.0c6f7ap-20f;           // down(1e-6)
    static constexpr float speed2_gate = 0x1.12e0bep-30f;             // down(1e-9)
    static constexpr float ground_y = 0x1.999998p-4f;                 // down(0.1)
    static constexpr float dead_band = 0x1.999998p-4f;                // down(0.1)
    static constexpr float throttle_lo_lt = 0x1.99999ap-4f;           // up(0.1)
    static constexpr float throttle_lo_gt = 0x1.999998p-4f;           // down(0.1)
    static constexpr float max_x = 50000.0f, max_y = 50000.0f, tip_angle = 90.0f;
    static constexpr float r_perfect = 4000.0f, r_good = 3000.0f, r_ok = 2000.0f, r_crash = -800.0f;
    static constexpr float r_oob = -300.0f, r_tip = -300.0f;
    static constexpr float k_throttle_descent = (float)0.3, k_cold_gas = (float)0.7, k_free_fall = (float)0.2;
    static constexpr float k_angle_throttle = (float)0.3, k_direction = 1.5f, gamma = (float)0.99;
    static constexpr int max_episode_steps = 1000;

    RL_HD static constexpr float land_lt(int level, int k) {
        return level == 0 ? (k == 2 ? 5.0f : 30.0f) : level == 1 ? (k == 2 ? 8.0f : 40.0f) : (k == 2 ? 10.0f : 100.0f);
    }
    RL_HD static constexpr float reset_lo(int k) {
        return k == 0 ? -1500.0f : k == 1 ? 1800.0f : k == 2 ? -25.0f : k == 3 ? -200.0f : k == 4 ? -5.0f
             : k == 5 ? -5.0f : k == 6 ? -15.0f : k == 7 ? -10.0f : k == 8 ? 34000.0f : 370000.0f;
    }
    RL_HD static constexpr float reset_hi(int k) {
        return k == 0 ? 1500.0f : k == 1 ? 2200.0f : k == 2 ? 25.0f : k == 3 ? -150.0f : k == 4 ? 5.0f
             : k == 5 ? 5.0f : k == 6 ? 15.0f : k == 7 ? 10.0f : k == 8 ? 38000.0f : 410000.0f;
    }
    RL_HD static constexpr float reset_span(int k) { return (float)((double)reset_hi(k) - (double)reset_lo(k)); }
    RL_HD static constexpr float obs_lo(int k) { return k == 1 ? 0.0f : reset_lo(k); }     // lander.py:76
    RL_HD static constexpr float obs_hi(int k) { return reset_hi(k); }
};

inline bool matches_default(const DevParams& d) {
    using D = DefaultParams;
    bool ok = d.dt == D::dt && d.dt_d == D::dt_d && d.damp_d == D::damp_d && d.alpha_dt2_d == D::alpha_dt2_d &&
              d.fuel_unit_d == D::fuel_unit_d && d.alpha_min_mass_d == D::alpha_min_mass_d &&
              d.pos_unit == D::pos_unit && d.pos_inv_unit == D::pos_inv_unit && d.delta_unit == D::delta_unit &&
              d.dt2_units == D::dt2_units && d.half_dt2_units == D::half_dt2_units && d.dt_units == D::dt_units && d.burn_units == D::burn_units && d.fuel_inv_unit == D::fuel_inv_unit && d.inv_dt == D::inv_dt && d.dt2 == D::dt2 && d.half_dt2 == D::half_dt2 &&
              d.gravity == D::gravity && d.thrust_power == D::thrust_power && d.drag_k == D::drag_k &&
              d.alpha_k == D::alpha_k && d.alpha_min_mass == D::alpha_min_mass && d.damp == D::damp &&
              d.burn_per_step == D::burn_per_step && d.mass_gate == D::mass_gate &&
              d.throttle_gate == D::throttle_gate && d.speed2_gate == D::speed2_gate && d.ground_y == D::ground_y &&
              d.dead_band == D::dead_band && d.throttle_lo_lt == D::throttle_lo_lt &&
              d.throttle_lo_gt == D::throttle_lo_gt && d.max_x == D::max_x && d.max_y == D::max_y &&
              d.tip_angle == D::tip_angle && d.r_perfect == D::r_perfect && d.r_good == D::r_good &&
              d.r_ok == D::r_ok && d.r_crash == D::r_crash && d.r_oob == D::r_oob && d.r_tip == D::r_tip &&
              d.k_throttle_descent == D::k_throttle_descent && d.k_cold_gas == D::k_cold_gas &&
              d.k_free_fall == D::k_free_fall && d.k_angle_throttle == D::k_angle_throttle &&
              d.k_direction == D::k_direction && d.gamma == D::gamma &&
              d.max_episode_steps == D::max_episode_steps;
    for (int l = 0; l < 3 && ok; ++l)
        for (int k = 0; k < 3; ++k) ok = ok && d.land_lt_[l][k] == D::land_lt(l, k);
    for (int k = 0; k < 8 && ok; ++k) ok = d.obs_lo_[k] == D::obs_lo(k) && d.obs_hi_[k] == D::obs_hi(k);
    for (int k = 0; k < 10 && ok; ++k) ok = d.reset_lo_[k] == D::reset_lo(k) && d.reset_span_[k] == D::reset_span(k);
    return ok;
}

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

// kg per unit of the (signed) fixed-point fuel word: 2^(k - 31) with 2^k the first power of two
// ABOVE the largest fuel mass of the reset distribution (config.yaml:66-67: 410 000 kg -> 2^-12 kg)
inline double fuel_unit_for(double max_fuel) {
    int k = 0;
    if (max_fuel >= 1.0) k = (int)floor(log2(max_fuel)) + 1;
    return ldexp(1.0, k - 31);
}

// metres per unit of the fixed-point positions: 2^(k - 31) with 2^k the first power of two above 1.3 x
// the largest coordinate a rocket can hold before the bounds check ends its episode (reward.py:274-276;
// the margin covers the last step), and never below the reset ranges (config.yaml: 50 000 m -> 2^-15 m)
inline double pos_unit_for(const RlParams& p) {
    double m = fmax(fabs(p.max_horizontal_position), fabs(p.max_altitude));
    for (int k = 0; k < 2; ++k) m = fmax(m, fmax(fabs(p.reset_low[k]), fabs(p.reset_high[k])));
    m *= 1.3;
    int k = 0;
    if (m >= 1.0) k = (int)floor(log2(m)) + 1;
    return ldexp(1.0, k - 31);
}

inline int derive_params(const RlParams& p, DevParams& d, char* err, size_t err_len) {
    if (!(p.time_step > 0)) return derive_fail(err, err_len, "time_step must be positive");          // engine.py:35
    if (p.thrust_power < 0 || p.cold_gas_thrust_power < 0 || p.fuel_consumption_rate < 0)
        return derive_fail(err, err_len, "thrust powers and fuel rate cannot be negative");          // engine.py:37-44
    if (!(p.radius > 0) || !(p.cold_gas_moment_arm > 0))
        return derive_fail(err, err_len, "radius and moment arm must be positive");                  // engine.py:45-46
    if (p.max_episode_steps < 1 || p.max_episode_steps > (int)RL_META_STEP_MASK)
        return derive_fail(err, err_len, "max_episode_steps out of range (1 .. %d)", (int)RL_META_STEP_MASK);
    if (!(p.reset_high[9] < 9.0e18))
        return derive_fail(err, err_len, "reset_high[9] (fuel mass) out of range");
    if (!(fmax(fabs(p.max_horizontal_position), fabs(p.max_altitude)) < 1.0e15))
        return derive_fail(err, err_len, "max_horizontal_position / max_altitude out of range");
    const double pi = 3.14159265358979323846;
    const double dt = p.time_step;
    d.dt = (float)dt;
    d.dt_d = dt;
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
    d.damp_d = fmax(0.0, 1.0 - p.angular_damping * dt);
    d.alpha_dt2_d = p.cold_gas_thrust_power * p.cold_gas_moment_arm / half_r2 * (180.0 / pi) * (dt * dt);
    d.alpha_min_mass_d = (p.radius <= 1e-6) ? INFINITY : 1e-6 / half_r2;
    d.burn_per_step = (float)(p.fuel_consumption_rate * dt);
    // fixed-point fuel: the smallest power-of-two unit whose int32 range covers the largest
    // fuel load the reset distribution can draw (fuel only ever decreases)
    d.fuel_unit_d = fuel_unit_for(p.reset_high[9]);
    d.fuel_inv_unit = (float)(1.0 / d.fuel_unit_d);
    d.burn_units = (float)(p.fuel_consumption_rate * dt / d.fuel_unit_d);
    const double pos_unit = pos_unit_for(p), delta_unit = ldexp(pos_unit, -kDeltaShift);
    d.pos_unit = (float)pos_unit;
    d.pos_inv_unit = (float)(1.0 / pos_unit);
    d.delta_unit = (float)delta_unit;
    d.dt2_units = (float)(dt * dt / delta_unit);
    d.half_dt2_units = (float)(0.5 * dt * dt / delta_unit);
    d.dt_units = (float)(dt / delta_unit);
    d.mass_gate = f32_down(1e-6);
    d.throttle_gate = f32_down(1e-6);
    d.speed2_gate = f32_down(1e-9);
    d.ground_y = f32_down(0.1);
    d.dead_band = f32_down(0.1);
    d.throttle_lo_lt = f32_up(0.1);
    d.throttle_lo_gt = f32_down(0.1);
    for (int k = 0; k < 3; ++k) {
        d.land_lt_[0][k] = f32_up(p.land_perfect[k]);
        d.land_lt_[1][k] = f32_up(p.land_good[k]);
        d.land_lt_[2][k] = f32_up(p.land_ok[k]);
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
        d.obs_lo_[k] = (float)p.obs_low[k];       // np.array(..., dtype=float32), lander.py:73-101
        d.obs_hi_[k] = (float)p.obs_high[k];
        if (!(d.obs_lo_[k] <= d.obs_hi_[k])) return derive_fail(err, err_len, "obs_low[%d] > obs_high[%d]", k, k);
    }
    for (int k = 0; k < 10; ++k) {
        if (!(p.reset_low[k] <= p.reset_high[k])) return derive_fail(err, err_len, "reset_low[%d] > reset_high[%d]", k, k);
        d.reset_lo_[k] = (float)p.reset_low[k];
        d.reset_span_[k] = (float)(p.reset_high[k] - p.reset_low[k]);
    }
    d.max_episode_steps = p.max_episode_steps;
    return RL_OK;
}

}  // namespace rl
