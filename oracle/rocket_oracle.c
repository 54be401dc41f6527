/*
 * CPU oracle in C: scalar fp64 restatement of the reference's environment step.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked or loaded by the product library.  Allowed
 * users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
 * legs.  It exists because the Python oracle (oracle/rocket_oracle.py, pinned
 * bit-for-bit against the live reference) manages ~2e4 steps/s, too slow to check
 * 4096 rockets x 1000 steps on the GPU box or to give a fair CPU baseline.
 *
 * Parity status: pinned.  tests/test_oracle_c.py checks this file against the
 * committed outputs of the reference itself (tests/golden/) and against the Python
 * oracle: flags and landing classes identical, state words within 1e-11 relative
 * (libm sin/cos and the non-BLAS v.v differ from numpy in the last ulp, SURVEY.md
 * section 8c).
 *
 * Each function cites the reference file:line it follows (paths relative to the
 * reference root).  Operation order follows the reference.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define PI_D 3.14159265358979323846

typedef struct {
    double dt, gravity, air_density, thrust_power, cold_gas_thrust_power, fuel_rate;
    double radius, moment_arm, reference_area, drag_coefficient, angular_damping;
    double reset_lo[10], reset_hi[10];       /* draw order: x y vx vy ax ay angle omega dry fuel */
    double land_perfect[3], land_good[3], land_ok[3];   /* |vx| |vy| |angle| */
    double max_horizontal_position, max_altitude, tip_over_angle;
    double r_landing_perfect, r_landing_good, r_landing_ok, r_crash_ground;
    double r_out_of_bounds, r_tipped_over;
    double throttle_descent_reward_scale, cold_gas_reward_scale, free_fall_penalty_scale;
    double angle_aware_throttle_scale, correct_direction_bonus, gamma;
    double obs_low[8], obs_high[8];
    int32_t max_episode_steps;
    int32_t _pad;
} OracleConstants;

/* Rocket.state + the three words of Rocket.previous_state the integrator reads */
typedef struct {
    double x, y, vx, vy, ax, ay, angle, ang_vel, dry_mass, fuel;
    double prev_x, prev_y, prev_angle;
    int32_t current_step;
    int32_t _pad;
} OracleRocket;

static inline double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* engine.py:230-237.  Python's float % is fmod with the sign of the divisor. */
static double normalize_angle_180(double a) {
    double r = fmod(a, 360.0);
    if (r != 0.0 && r < 0.0) r += 360.0;
    if (r >= 180.0) r -= 360.0;
    return r;
}

/* engine.py:100-115 (gravity :48-52, thrust :54-76, drag :78-98) */
static void net_force(const OracleConstants *c, double m, double throttle, double angle_deg,
                      double vx, double vy, double *fx, double *fy) {
    if (m <= 0) { *fx = 0.0; *fy = 0.0; return; }
    double gy = m * c->gravity;
    double th = clampd(throttle, 0.0, 1.0), tx = 0.0, ty = 0.0;
    if (th > 1e-6) {
        double rad = angle_deg * (PI_D / 180.0);            /* np.deg2rad */
        double mag = th * c->thrust_power;
        tx = mag * sin(rad);
        ty = mag * cos(rad);
    }
    double s2 = vx * vx + vy * vy, dx = 0.0, dy = 0.0;
    if (s2 > 1e-9) {
        double sp = sqrt(s2);
        double dm = 0.5 * c->air_density * c->drag_coefficient * c->reference_area * s2;
        dx = dm * (-vx / sp);
        dy = dm * (-vy / sp);
    }
    *fx = (0.0 + tx) + dx;
    *fy = (gy + ty) + dy;
}

/* engine.py:125-155 */
static double angular_acceleration(const OracleConstants *c, double cold_gas, double m) {
    double cg = clampd(cold_gas, -1.0, 1.0);
    if (m <= 1e-6 || c->radius <= 1e-6) return 0.0;
    double inertia = 0.5 * m * (c->radius * c->radius);
    if (inertia < 1e-6) return 0.0;
    double torque = (c->cold_gas_thrust_power * cg) * c->moment_arm;
    return (torque / inertia) * (180.0 / PI_D);             /* np.rad2deg */
}

/* base.py:65-130 */
void oracle_bootstrap(const OracleConstants *c, OracleRocket *r, int64_t n) {
    const double dt = c->dt, dt2 = dt * dt;
    for (int64_t i = 0; i < n; ++i) {
        OracleRocket *s = &r[i];
        double m0 = s->dry_mass + s->fuel, ax0 = 0.0, ay0 = 0.0, al0 = 0.0;
        if (m0 > 1e-6) {
            double fx, fy;
            net_force(c, m0, 0.0, s->angle, s->vx, s->vy, &fx, &fy);
            ax0 = fx / m0;
            ay0 = fy / m0;
            al0 = angular_acceleration(c, 0.0, m0);
        }
        s->prev_x = s->x - (s->vx * dt) + (0.5 * ax0 * dt2);
        s->prev_y = s->y - (s->vy * dt) + (0.5 * ay0 * dt2);
        s->prev_angle = normalize_angle_180(s->angle - (s->ang_vel * dt) + (0.5 * al0 * dt2));
    }
}

/* base.py:132-184 with engine.py:157-228 inlined */
static void apply_action(const OracleConstants *c, OracleRocket *s, double throttle, double cold_gas) {
    const double dt = c->dt, dt2 = dt * dt;
    double th = clampd(throttle, 0.0, 1.0), cg = clampd(cold_gas, -1.0, 1.0);
    double fuel_now = s->fuel;
    if (fuel_now <= 0) { th = 0.0; s->fuel = 0.0; }
    double m = s->dry_mass + fuel_now;
    if (m <= 1e-6) return;
    double fx, fy;
    net_force(c, m, th, s->angle, s->vx, s->vy, &fx, &fy);
    s->ax = fx / m;
    s->ay = fy / m;
    double alpha = angular_acceleration(c, cg, m);
    double nx = 2.0 * s->x - s->prev_x + s->ax * dt2;
    double ny = 2.0 * s->y - s->prev_y + s->ay * dt2;
    double nvx = (nx - s->x) / dt, nvy = (ny - s->y) / dt;
    double damp = fmax(0.0, 1.0 - (c->angular_damping * dt));
    double change = (s->angle - s->prev_angle) * damp;
    double na = s->angle + change + alpha * dt2;
    double nom = (na - s->angle) / dt;
    na = normalize_angle_180(na);
    s->prev_x = s->x; s->prev_y = s->y; s->prev_angle = s->angle;
    s->x = nx; s->y = ny; s->vx = nvx; s->vy = nvy; s->angle = na; s->ang_vel = nom;
    double burned = fmax(0.0, th * c->fuel_rate * dt);
    s->fuel = fmax(0.0, fuel_now - burned);
}

/* utils.py:1-37; codes as backend/protocol.py:31-41: 1 safe 2 good 3 ok 4 unsafe */
static int classify_landing(const OracleConstants *c, double vx, double vy, double angle) {
    double a = fabs(vx), b = fabs(vy), g = fabs(angle);
    if (a < c->land_perfect[0] && b < c->land_perfect[1] && g < c->land_perfect[2]) return 1;
    if (a < c->land_good[0] && b < c->land_good[1] && g < c->land_good[2]) return 2;
    if (a < c->land_ok[0] && b < c->land_ok[1] && g < c->land_ok[2]) return 3;
    return 4;
}

/* reward.py:231-264 */
static double potential(double y, double vx, double vy, double angle, double omega) {
    double yp = fmax(0.0, y), yf = fmin(1.0, yp / 2000.0);
    double w_vy = 0.015 * (2.0 - yf), w_ang = 0.01 * (2.0 - yf), w_stab = 0.05 * (2.0 - yf);
    return -0.005 * yp - w_vy * fabs(vy) - 0.005 * fabs(vx) - w_ang * fabs(angle) - w_stab * fabs(omega);
}

/* reward.py:27-281; throttle / cold_gas are the RAW action (:69-70) */
static double reward_and_flags(const OracleConstants *c, const OracleRocket *b, double throttle,
                               double cold_gas, const OracleRocket *a, int *terminated,
                               int *truncated, int *landing) {
    double total = 0.0;
    *terminated = 0; *truncated = 0; *landing = 0;
    double y_a = a->y, vy_a = a->vy, ang_a = a->angle, om_a = a->ang_vel;
    if (y_a <= 0.1 && b->y > 0.1) {                                   /* :73-102 */
        int cls = classify_landing(c, a->vx, a->vy, a->angle);
        double ab = fmax(0.0, 1.0 - (fabs(ang_a) / 10.0));
        double vb = fmax(0.0, 1.0 - (fabs(vy_a) / 5.0));
        double q = 0.6 * ab + 0.4 * vb;
        if (cls == 1) total += c->r_landing_perfect * (1.0 + 0.5 * q);
        else if (cls == 2) total += c->r_landing_good * (0.8 + 0.2 * q);
        else if (cls == 3) total += c->r_landing_ok;
        else {
            double sev = fmin(1.0, (fabs(vy_a) / 20.0 + fabs(ang_a) / 45.0) / 2.0);
            total += c->r_crash_ground * (0.7 + 0.3 * sev);
        }
        *terminated = 1; *landing = cls;
        return total;
    }
    double ae_b = fabs(b->angle), ae_a = fabs(ang_a), we_b = fabs(b->ang_vel), we_a = fabs(om_a);
    double corr = -(ae_a - ae_b) * 0.5;                               /* :117-124 */
    corr -= (we_a - we_b) * 0.1;
    total += corr;
    if (ae_b > 0.1 || we_b > 0.1) {                                   /* :130-162 */
        int correct = (b->angle > 0 && cold_gas < 0) || (b->angle < 0 && cold_gas > 0);
        double needed = ae_b + we_b;
        double effect = (ae_b - ae_a) + (we_b - we_a);
        double mult = correct ? c->correct_direction_bonus : -c->correct_direction_bonus;
        double cgr = fabs(cold_gas) * needed * mult * c->cold_gas_reward_scale;
        if (effect > 0) cgr *= 1.0 + effect * effect;
        total += cgr;
    } else {
        total -= fabs(cold_gas) * 0.3 * c->cold_gas_reward_scale;     /* :166-167 */
    }
    if (y_a > 0.1 && vy_a < 0) {                                      /* :171-186 */
        double af = fmax(0.0, 1.0 - (fabs(ang_a) / 45.0));
        double df = fmin(1.0, fabs(vy_a) / 10.0);
        total += throttle * (-vy_a) * af * (1.0 + df) * c->throttle_descent_reward_scale;
    }
    if (y_a > 0.1 && vy_a < 0 && throttle < 0.1) {                    /* :190-201 */
        double prox = 1.0 + (8.0 / fmax(y_a, 1.0));
        double sf = fmin(1.0, fabs(vy_a) / 15.0);
        total -= c->free_fall_penalty_scale * (-vy_a) * prox * (1.0 + sf);
    }
    if (throttle > 0.1 && y_a > 0.1) {                                /* :205-219 */
        if (fabs(ang_a) > 10.0)
            total -= throttle * (fabs(ang_a) / 90.0) * c->angle_aware_throttle_scale;
        else if (fabs(ang_a) > 5.0 && fabs(ang_a) <= 10.0)
            total -= throttle * ((fabs(ang_a) - 5.0) / 5.0) * c->angle_aware_throttle_scale * 0.5;
    }
    if (y_a > 0.1 && vy_a > 0) {                                      /* :222-227 */
        double alt = fmin(1.0, y_a / 1000.0), asc = fmin(1.0, vy_a / 5.0);
        total -= vy_a * 0.5 * alt * (1.0 + asc);
    }
    double phi_b = potential(b->y, b->vx, b->vy, b->angle, b->ang_vel);   /* :267-270 */
    double phi_a = potential(a->y, a->vx, a->vy, a->angle, a->ang_vel);
    total += c->gamma * phi_a - phi_b;
    if (fabs(a->x) > c->max_horizontal_position || y_a > c->max_altitude) {   /* :274-276 */
        total += c->r_out_of_bounds;
        *truncated = 1;
    }
    if (fabs(ang_a) > c->tip_over_angle) total += c->r_tipped_over;   /* :278-279 */
    return total;
}

/* lander.py:124-150: float32 cast, then clip */
static void clip_observation(const OracleConstants *c, const OracleRocket *s, float *obs) {
    const double raw[8] = {s->x, s->y, s->vx, s->vy, s->ax, s->ay, s->angle, s->ang_vel};
    for (int k = 0; k < 8; ++k) {
        float v = (float)raw[k], lo = (float)c->obs_low[k], hi = (float)c->obs_high[k];
        obs[k] = v < lo ? lo : (v > hi ? hi : v);
    }
}

void oracle_observe(const OracleConstants *c, const OracleRocket *r, int64_t n, float *obs) {
    for (int64_t i = 0; i < n; ++i) clip_observation(c, &r[i], obs + 8 * i);
}

/* lander.py:188-255 for n independent rockets, no reset.  actions float32 [n,2]. */
void oracle_step(const OracleConstants *c, OracleRocket *r, int64_t n, const float *actions,
                 float *obs, double *reward, uint8_t *terminated, uint8_t *truncated,
                 int8_t *landing) {
    for (int64_t i = 0; i < n; ++i) {
        OracleRocket before = r[i];
        double th = (double)actions[2 * i], cg = (double)actions[2 * i + 1];
        apply_action(c, &r[i], th, cg);
        r[i].current_step += 1;
        int term, trunc, cls;
        double rew = reward_and_flags(c, &before, th, cg, &r[i], &term, &trunc, &cls);
        if (r[i].current_step >= c->max_episode_steps) trunc = 1;      /* :237-241 */
        if (obs) clip_observation(c, &r[i], obs + 8 * i);
        if (reward) reward[i] = rew;
        if (terminated) terminated[i] = (uint8_t)term;
        if (truncated) truncated[i] = (uint8_t)trunc;
        if (landing) landing[i] = (int8_t)cls;
    }
}

/* reward.py:27-281 + lander.py:237-241 evaluated on CALLER-SUPPLIED before/after states
 * (teacher-forced check: the CUDA path's own fp32 states widened to fp64).  Only the words
 * the reward reads need to be filled in `before` (y, vx, vy, angle, ang_vel) and `after`
 * (x, y, vx, vy, angle, ang_vel, current_step). */
void oracle_reward_only(const OracleConstants *c, const OracleRocket *before, const OracleRocket *after,
                        int64_t n, const float *actions, double *reward, uint8_t *terminated,
                        uint8_t *truncated, int8_t *landing) {
    for (int64_t i = 0; i < n; ++i) {
        int term, trunc, cls;
        reward[i] = reward_and_flags(c, &before[i], (double)actions[2 * i], (double)actions[2 * i + 1],
                                     &after[i], &term, &trunc, &cls);
        if (after[i].current_step >= c->max_episode_steps) trunc = 1;
        terminated[i] = (uint8_t)term; truncated[i] = (uint8_t)trunc; landing[i] = (int8_t)cls;
    }
}

/* ---- reset sampling + random-action rollout (CPU baseline timing) ------------------ */
static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t *s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

/* simulation/config.py:26-53: ten uniform draws lo + (hi - lo) * u, in order; then
 * the Verlet bootstrap (base.py:186-218) and current_step = 0 (lander.py:180) */
void oracle_reset(const OracleConstants *c, OracleRocket *s, uint64_t *rng) {
    double d[10];
    for (int k = 0; k < 10; ++k) d[k] = c->reset_lo[k] + (c->reset_hi[k] - c->reset_lo[k]) * u01(rng);
    s->x = d[0]; s->y = d[1]; s->vx = d[2]; s->vy = d[3]; s->ax = d[4]; s->ay = d[5];
    s->angle = d[6]; s->ang_vel = d[7]; s->dry_mass = d[8]; s->fuel = d[9];
    oracle_bootstrap(c, s, 1);
    s->current_step = 0;
}

typedef struct {
    const OracleConstants *c;
    int64_t n_envs, steps;
    uint64_t seed;
    double sum_reward;
    int64_t episodes, env_steps;
    int64_t landings[5];
} RolloutJob;

static void *rollout_worker(void *arg) {
    RolloutJob *j = (RolloutJob *)arg;
    uint64_t rng = j->seed;
    OracleRocket *r = (OracleRocket *)calloc((size_t)j->n_envs, sizeof(OracleRocket));
    for (int64_t i = 0; i < j->n_envs; ++i) oracle_reset(j->c, &r[i], &rng);
    float obs[8];
    for (int64_t t = 0; t < j->steps; ++t) {
        for (int64_t i = 0; i < j->n_envs; ++i) {
            float act[2] = {(float)u01(&rng), (float)(2.0 * u01(&rng) - 1.0)};
            double rew; uint8_t term, trunc; int8_t cls;
            oracle_step(j->c, &r[i], 1, act, obs, &rew, &term, &trunc, &cls);
            j->sum_reward += rew + obs[0] * 1e-30;   /* keep obs live */
            j->env_steps += 1;
            if (term || trunc) {                     /* a16: SB3-style auto-reset */
                j->episodes += 1;
                j->landings[cls] += 1;
                oracle_reset(j->c, &r[i], &rng);
                clip_observation(j->c, &r[i], obs);
            }
        }
    }
    free(r);
    return NULL;
}

/* Random-action rollout with auto-reset over n_threads host threads, each owning
 * n_envs_per_thread rockets for `steps` steps.  out[0] = env steps, out[1] = wall
 * seconds, out[2] = episodes, out[3] = sum of rewards, out[4..8] = landing counts. */
int oracle_rollout_random(const OracleConstants *c, int64_t n_envs_per_thread, int64_t steps,
                          int n_threads, uint64_t seed, double *out) {
    if (n_threads < 1 || n_threads > 1024) return -1;
    pthread_t *tid = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    RolloutJob *jobs = (RolloutJob *)calloc((size_t)n_threads, sizeof(RolloutJob));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < n_threads; ++k) {
        jobs[k].c = c; jobs[k].n_envs = n_envs_per_thread; jobs[k].steps = steps;
        jobs[k].seed = seed * 0x9E3779B97F4A7C15ull + (uint64_t)k * 0xD1B54A32D192ED03ull + 1;
        pthread_create(&tid[k], NULL, rollout_worker, &jobs[k]);
    }
    memset(out, 0, 9 * sizeof(double));
    for (int k = 0; k < n_threads; ++k) {
        pthread_join(tid[k], NULL);
        out[0] += (double)jobs[k].env_steps;
        out[2] += (double)jobs[k].episodes;
        out[3] += jobs[k].sum_reward;
        for (int q = 0; q < 5; ++q) out[4 + q] += (double)jobs[k].landings[q];
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    out[1] = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    free(tid); free(jobs);
    return 0;
}

int64_t oracle_sizeof_constants(void) { return (int64_t)sizeof(OracleConstants); }
int64_t oracle_sizeof_rocket(void) { return (int64_t)sizeof(OracleRocket); }
