/*
 * rocket_landing_b200.h -- C ABI of the B200-native batched rocket-landing environment step.
 *
 * This is the drop-in boundary for ONE path of adimail/rocket-landing-rl: the environment
 * step behind RocketLandingEnv.reset()/step() (reference backend/envs/lander.py:168,188),
 * i.e. physics (backend/physics/engine.py, backend/rocket/base.py:132), reward /
 * termination / landing class (backend/rl/reward.py:27, backend/utils.py:1), observation
 * clipping (lander.py:124) and the VecEnv auto-reset the reference trains under
 * (scripts/train.py:82-88).  The reference is pure Python and has no FFI; these entry
 * points are what a ctypes binding inside the reference would call (INTEGRATION.md shows
 * that binding).
 *
 * Conventions
 *  - every function returns 0 on success, a negative RL_E_* code otherwise;
 *    rl_last_error() returns a thread-local, human-readable message for the last failure.
 *  - all buffer arguments of rl_reset / rl_step / rl_set_state / rl_get_state / rl_stats are
 *    DEVICE pointers on the handle's device, owned by the caller (e.g. torch tensors); the
 *    library borrows them for the duration of the enqueue.  rl_step_host / rl_reset_host
 *    take HOST pointers and do the copies themselves.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls only
 *    ENQUEUE work on it and never synchronise (rl_*_host excepted); rl_step allocates
 *    nothing and is CUDA-graph capturable.  A handle is not thread-safe.  The rl_*_host calls
 *    (and rl_get_epoch / rl_set_epoch) run on private streams of the handle; they first wait
 *    for everything the handle's device-API calls have enqueued on the stream of the LAST such
 *    call, and return only when their own work is complete -- so the two families may be mixed
 *    on one handle without explicit synchronisation as long as the device-API calls of a
 *    handle use one stream at a time (work left on an EARLIER stream is the caller's to order;
 *    a stream under CUDA-graph capture cannot be waited for).
 *  - layouts: actions [N,2] f32 row-major (throttle, cold gas); obs [N,8] f32 row-major
 *    (x, y, vx, vy, ax, ay, angle, angular velocity; lander.py:129-141); reward [N] f32;
 *    terminated / truncated [N] u8 (0/1).
 */
#ifndef ROCKET_LANDING_B200_H_
#define ROCKET_LANDING_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RL_ABI_VERSION 2

#define RL_OK 0
#define RL_E_INVALID (-1)   /* bad argument                        */
#define RL_E_CUDA (-2)      /* a CUDA runtime call / launch failed */
#define RL_E_NOMEM (-3)     /* allocation failed                   */

/* Words of the public state format used by rl_set_state / rl_get_state: 32-bit words [10][N]
 * (structure of arrays, word-major) in a uint32 buffer; "f32" words hold IEEE-754 bits.  They are
 * the kernel's own persistent words, so a get -> set round trip is lossless.  Words that INTEGRATE
 * over an episode are 32-bit FIXED POINT: an fp32 rounding error in them would persist for the
 * rest of the episode, and the reference runs 1000-step episodes in fp64 (DESIGN.md "Long-horizon
 * accuracy").  rocket-landing-rl_b200/layout.py converts to and from the reference's Rocket.state
 * words. */
#define RL_STATE_WORDS 10
#define RL_W_X 0        /* i32  units of rl_pos_unit() m (2^-15 m for the reference's +-50 km world)   */
#define RL_W_Y 1        /* i32  units of rl_pos_unit() m                                       */
#define RL_W_ANGLE 2    /* i32  units of 360 / 2^32 deg: int32 range == [-180, 180), so integer
                           wrap-around IS normalize_angle_180 (engine.py:230)                  */
#define RL_W_SX 3       /* i32  carried x delta x(t) - x(t-dt), units of rl_pos_unit() / 256 m (2^-23 m),
                           so vx = delta / dt (engine.py:193-194); while FRESH: vx dt          */
#define RL_W_SY 4       /* i32  carried y delta, same units; while FRESH: vy dt                */
#define RL_W_SANGLE 5   /* f32  carried pre-wrap angle delta in deg (high part; 8 more bits in
                           RL_W_META), so omega = delta / dt; while FRESH: omega dt            */
#define RL_W_DRY 6      /* f32  kg                                                             */
#define RL_W_FUEL 7     /* i32  units of rl_fuel_unit() kg (2^-12 kg for the reference config) */
#define RL_W_META 8     /* u32 bits: [0,16) episode step count (lander.py:229),
                           [16,24) signed low-order extension of RL_W_SANGLE, in units of
                           ulp(RL_W_SANGLE) / 256,
                           [24,27) signed count k of 360-degree wraps applied by the last
                           normalisation (engine.py:226 stores the normalised angle as the
                           next step's "previous" angle, base.py:171),
                           [27,30) landing code and bit 30 "landed" (simulation mode only),
                           bit 31 FRESH = just reset: the first step still has to subtract the
                           Verlet bootstrap term a0 dt^2 / 2 from the deltas (base.py:65-130)  */
#define RL_W_EPRET 9    /* f32  running episode return                                         */

#define RL_META_STEP_MASK 0x0000FFFFu
#define RL_META_DLO_SHIFT 16
#define RL_META_WRAP_SHIFT 24
#define RL_META_SIM_MASK 0x78000000u   /* bits [27,31): simulation-mode landing code + landed flag */
#define RL_META_FRESH 0x80000000u
#define RL_ANGLE_UNITS_PER_DEG (4294967296.0 / 360.0)
#define RL_DELTA_UNITS_PER_POS_UNIT 256

/* Slots of the statistics vector (float64 [16]) returned by rl_stats. */
#define RL_STATS_WORDS 16
#define RL_S_EPISODES 0     /* finished episodes                                   */
#define RL_S_RETURN_SUM 1   /* sum of their returns                                */
#define RL_S_LENGTH_SUM 2   /* sum of their lengths (steps)                        */
#define RL_S_SAFE 3         /* landing class counts (utils.py:19-30)               */
#define RL_S_GOOD 4
#define RL_S_OK 5
#define RL_S_UNSAFE 6
#define RL_S_TIME_LIMIT 7   /* truncated by rl.max_episode_steps (lander.py:237)   */
#define RL_S_OUT_OF_BOUNDS 8 /* truncated by position limits (reward.py:274)       */
#define RL_S_TIP_STEPS 9    /* steps with |angle| > tip_over_angle (reward.py:278) */
#define RL_S_NONFINITE 10   /* steps whose reward or altitude was NaN / Inf        */
#define RL_S_ENV_STEPS 11   /* env steps executed                                  */
#define RL_S_REWARD_SUM 12  /* sum of all step rewards                             */

/* Constants of the environment: one field per config.yaml key the path reads
 * (reference config.yaml; loaded there by backend/config.py:21-37). */
typedef struct RlParams {
    double time_step;                /* simulation.time_step            */
    double gravity;                  /* environment.gravity             */
    double air_density;              /* environment.air_density         */
    double thrust_power;             /* rocket.thrust_power             */
    double cold_gas_thrust_power;    /* rocket.cold_gas_thrust_power    */
    double fuel_consumption_rate;    /* rocket.fuel_consumption_rate    */
    double radius;                   /* rocket.radius                   */
    double cold_gas_moment_arm;      /* rocket.cold_gas_moment_arm      */
    double reference_area;           /* rocket.reference_area           */
    double drag_coefficient;         /* rocket.drag_coefficient         */
    double angular_damping;          /* rocket.angular_damping          */
    /* reset distribution, draw order of simulation/config.py:26-40:
     * x, y, vx, vy, ax, ay, angle, angular velocity, dry mass, fuel mass */
    double reset_low[10];
    double reset_high[10];
    /* landing.thresholds.{perfect,good,ok}.{speed_vx,speed_vy,angle} */
    double land_perfect[3];
    double land_good[3];
    double land_ok[3];
    double max_horizontal_position;  /* rl.max_horizontal_position      */
    double max_altitude;             /* rl.max_altitude                 */
    double tip_over_angle;           /* rl.tip_over_angle               */
    double landing_perfect;          /* rl.rewards.*                    */
    double landing_good;
    double landing_ok;
    double crash_ground;
    double out_of_bounds;
    double tipped_over;
    double throttle_descent_reward_scale;
    double cold_gas_reward_scale;
    double free_fall_penalty_scale;
    double angle_aware_throttle_scale;
    double correct_direction_bonus;
    double gamma;
    double obs_low[8];               /* observation Box bounds (lander.py:73-101) */
    double obs_high[8];
    int32_t max_episode_steps;       /* rl.max_episode_steps            */
    int32_t reserved;
} RlParams;

typedef struct RlEnv RlEnv;          /* opaque handle */

int rl_abi_version(void);
const char *rl_last_error(void);

/* Fill `p` with the reference's config.yaml defaults (config.yaml:16-91). */
void rl_params_default(RlParams *p);

/* Bytes of device memory holding the persistent state of n_envs rockets (40 B per rocket,
 * padded to a whole number of 256-rocket tiles). */
int64_t rl_state_bytes(int64_t n_envs);

/* Create an environment batch of n_envs rockets on `device`.
 *   env_id_offset : global id of rocket 0 of this shard; the Philox reset stream is keyed by
 *                   (seed, global rocket id, launch epoch) so results do not depend on how
 *                   rockets are sharded over GPUs.
 *   state_buf     : caller-owned device buffer of rl_state_bytes(n_envs) bytes, 256-byte
 *                   aligned (so the caller can checkpoint it), or NULL to let the library
 *                   allocate and own it.
 * Replaces: RocketLandingEnv.__init__ (lander.py:48) x n_envs + make_vec_env (train.py:83). */
int rl_create(const RlParams *params, int64_t n_envs, int64_t env_id_offset, uint64_t seed,
              int device, void *state_buf, RlEnv **out);
int rl_destroy(RlEnv *env);

int64_t rl_num_envs(const RlEnv *env);

/* kg per unit of the fixed-point fuel word RL_W_FUEL: 2^(k-31), 2^k being the first power of two
 * above params->reset_high[9] (the largest fuel load a reset can draw; fuel only decreases).
 * rl_fuel_unit_for computes the same from a parameter set without a handle. */
double rl_fuel_unit(const RlEnv *env);
double rl_fuel_unit_for(const RlParams *params);

/* metres per unit of the fixed-point position words RL_W_X / RL_W_Y: 2^(k-31), 2^k being the first
 * power of two above 1.3 x max(max_horizontal_position, max_altitude, |reset range of x, y|) -- the
 * bounds check (reward.py:274-276) ends an episode before a rocket can leave that range; positions
 * saturate at its edge.  The deltas RL_W_SX / RL_W_SY count 1/256 of it. */
double rl_pos_unit(const RlEnv *env);
double rl_pos_unit_for(const RlParams *params);

/* 1 if this handle runs the kernel instantiation with the reference's config.yaml constants
 * folded in at compile time (chosen only when every derived constant is bit-identical to those
 * defaults), 0 if it runs the instantiation that reads its constants at run time.  Both are the
 * same template; setting the environment variable RL_B200_GENERIC_KERNEL=1 before rl_create
 * forces the run-time one. */
int rl_uses_folded_constants(const RlEnv *env);

/* Launch epoch: counts rl_reset / rl_step launches and is part of the Philox counter.  It
 * lives in device memory and is advanced by the kernels themselves, so a CUDA-graph replay
 * of rl_step advances the reset stream like an ordinary call.  Saving it with the state
 * buffer makes a checkpoint resume bit-identically.  Both calls wait for the work this handle has
 * enqueued (see "Conventions"), not for the whole device. */
uint64_t rl_get_epoch(RlEnv *env);
int rl_set_epoch(RlEnv *env, uint64_t epoch);

/* Resample the initial state of every rocket (mask == NULL) or of the rockets with
 * mask[i] != 0, and write their clipped reset observation (rows of unmasked rockets are
 * left untouched).  obs may be NULL.
 * Replaces: RocketLandingEnv.reset (lander.py:168) -> Rocket.reset (base.py:186) ->
 * get_initial_state (simulation/config.py:26). */
int rl_reset(RlEnv *env, const uint8_t *mask, float *obs, void *stream);

/* One environment step for every rocket: `substeps` consecutive physics+reward steps with
 * the same action (rewards summed, integration of a rocket stops at its first done), then
 * -- if auto_reset != 0 -- rockets that are done are resampled and `obs` holds their reset
 * observation (the SB3 VecEnv contract of scripts/train.py:82-88).
 *   terminal_obs   [N,8] f32, nullable: rows of done rockets receive the terminal observation
 *   episode_return [N] f32, nullable:   entries of done rockets receive the episode return
 *   episode_length [N] i32, nullable:   entries of done rockets receive the episode length
 *                                       (other rows / entries are left untouched)
 *   landing        [N] i8,  nullable:   landing class of this step for EVERY rocket: 0 = no
 *                                       touchdown, 1 safe, 2 good, 3 ok, 4 unsafe
 *                                       (evaluate_landing, utils.py:1; codes of protocol.py:31-41)
 *   raw_accel      [N,2] f32, nullable: the UNCLIPPED accelerations (ax, ay) of this step for
 *                                       every rocket -- what info["raw_state"] carries
 *                                       (lander.py:157-166); the observation clips them to +-5
 * Replaces: RocketLandingEnv.step (lander.py:188) x n_envs. */
int rl_step(RlEnv *env, const float *actions, float *obs, float *reward, uint8_t *terminated,
            uint8_t *truncated, float *terminal_obs, float *episode_return,
            int32_t *episode_length, int8_t *landing, float *raw_accel, int substeps, int auto_reset,
            void *stream);

/* Inject / read back the persistent state in the public format (uint32 [10][N], RL_W_*).
 * rl_set_state writes only rockets with mask[i] != 0 (mask == NULL: all).  Used for
 * deterministic replay against the reference and for checkpointing. */
int rl_set_state(RlEnv *env, const uint32_t *state_words, const uint8_t *mask, void *stream);
int rl_get_state(RlEnv *env, uint32_t *state_words, void *stream);

/* Copy the float64[16] statistics vector (RL_S_*) to the DEVICE buffer `out16` (ready for an
 * NCCL all-reduce on the same stream); reset_after != 0 zeroes the accumulators afterwards. */
int rl_stats(RlEnv *env, double *out16, int reset_after, void *stream);

/* Host-buffer variants (the reference-facing call with numpy arrays): copy `actions` to the
 * device, step, copy obs / reward / flags back, synchronise.  Buffers are ordinary host
 * memory (pinned memory makes the copies faster).  Nullable outputs as in rl_step. */
int rl_step_host(RlEnv *env, const float *actions, float *obs, float *reward,
                 uint8_t *terminated, uint8_t *truncated, float *terminal_obs,
                 float *episode_return, int32_t *episode_length, int8_t *landing, float *raw_accel,
                 int substeps, int auto_reset);
int rl_reset_host(RlEnv *env, float *obs);
/* rl_get_state into HOST memory (uint32 [10][N]); synchronous. */
int rl_get_state_host(RlEnv *env, uint32_t *state_words);

/* Simulation mode -- the reference's second caller of the step, the UI loop: one tick of
 * SimulationController.step (backend/simulation/controller.py:223-269) over RocketControls.step
 * (backend/rocket/controls.py:27-113) for every rocket, emitting the BinaryProtocol record of
 * backend/protocol.py (16 float32 = 64 bytes per rocket: x y vx vy ax ay angle angularVelocity
 * angularAcceleration mass fuelMass reward throttle coldGas landing_code is_active).  The action is
 * clipped before physics AND reward, there is no time limit and no auto-reset, a rocket freezes
 * at touchdown and is reported inactive (reward NaN) from the next tick on; rl_reset re-arms it.
 *   rl_sim_step      : DEVICE buffers; telemetry float32 [N][16]; reward [N] / done [N] nullable.
 *   rl_sim_step_host : HOST buffers; `message` receives the complete wire message,
 *                      1 header byte (1 = telemetry) + 64 N bytes. */
int rl_sim_step(RlEnv *env, const float *actions, float *telemetry, float *reward, uint8_t *done,
                void *stream);
int rl_sim_step_host(RlEnv *env, const float *actions, uint8_t *message);

/* Number of kernels this handle has launched so far (step / reset / state / stats). */
int64_t rl_launch_count(const RlEnv *env);

/* ---------------------------------------------------------------------------------------------
 * On-device VecNormalize -- the wrapper the reference puts around its vectorised environment
 * (scripts/train.py:86-88; statistics re-loaded for inference at backend/rl/agent.py:73-80,181).
 * Semantics of stable-baselines3 2.3.2 (running mean / variance of the observations, clip to
 * +-clip_obs; reward divided by the running std of the discounted return, clip to +-clip_reward).
 * All buffers are DEVICE pointers; calls only enqueue on `stream`; outputs may alias inputs. */
typedef struct RlVecNorm RlVecNorm;

const char *rl_vecnorm_last_error(void);
int rl_vecnorm_create(int64_t n_envs, double gamma, double epsilon, double clip_obs, double clip_reward,
                      int device, RlVecNorm **out);
int rl_vecnorm_destroy(RlVecNorm *vn);
/* VecNormalize.reset() on the observations returned by rl_reset. */
int rl_vecnorm_reset(RlVecNorm *vn, const float *obs_in, float *obs_out, int training, int norm_obs, void *stream);
/* VecNormalize.step_wait() on the outputs of rl_step; terminal_obs (nullable) is normalised in
 * place for done rockets (SB3 normalises infos[i]["terminal_observation"]). */
int rl_vecnorm_step(RlVecNorm *vn, const float *obs_in, const float *reward_in, const uint8_t *terminated,
                    const uint8_t *truncated, float *terminal_obs, float *obs_out, float *reward_out,
                    int training, int norm_obs, int norm_reward, void *stream);
/* normalize_obs of `rows` observations with the current statistics (evaluation / inference). */
int rl_vecnorm_normalize_obs(RlVecNorm *vn, const float *obs_in, float *obs_out, int64_t rows, void *stream);
/* Statistics as 20 doubles in HOST memory: obs mean[8], obs var[8], obs count, return mean, var,
 * count (what SB3 pickles as vecnormalize.pkl).  Both calls synchronise the device. */
int rl_vecnorm_get_stats(RlVecNorm *vn, double *out20);
int rl_vecnorm_set_stats(RlVecNorm *vn, const double *in20);
int64_t rl_vecnorm_launch_count(const RlVecNorm *vn);

/* ---------------------------------------------------------------------------------------------
 * Generalised advantage estimation over a device-resident rollout buffer: what SB3's
 * RolloutBuffer.compute_returns_and_advantage does for the reference's PPO (scripts/train.py:
 * 119-134, gamma / gae_lambda of config.yaml:100-104).  DEVICE pointers; rewards, values,
 * advantages, returns are float32 [T][N]; episode_starts uint8 [T][N]; last_values float32 [N],
 * last_dones uint8 [N] describe the state after the last step.  Enqueue only. */
const char *rl_rollout_last_error(void);
int rl_gae(const float *rewards, const float *values, const uint8_t *episode_starts,
           const float *last_values, const uint8_t *last_dones, float *advantages, float *returns, int T,
           int64_t n_envs, double gamma, double gae_lambda, void *stream);

/* What SB3's collect_rollouts (on_policy_algorithm.py, driven by scripts/train.py:136-144) derives on the
 * host from `dones` and `infos` after every env step, in one pass: episode_start_next[i] = terminated[i] |
 * truncated[i] (nullable), time_limit[i] = truncated[i] & !terminated[i] (the rockets whose reward gets the
 * gamma V(terminal_observation) bootstrap, infos[i]["TimeLimit.truncated"]), and *any_time_limit = 1
 * (int32: device memory, or pinned host memory -- the rollout reads it on the host after an event, without a
 * copy) if there is any such rocket.  Flags are uint8 [N] with values 0 / 1.  Enqueue only. */
int rl_episode_flags(const uint8_t *terminated, const uint8_t *truncated, uint8_t *episode_start_next,
                     uint8_t *time_limit, int32_t *any_time_limit, int64_t n_envs, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused inference of one of the reference's actor-critic MLPs (scripts/train.py:117: net_arch =
 * dict(pi=[256, 256], vf=[256, 256]), tanh; assets/model/v3): out = W3 tanh(W2 tanh(W1 obs + b1) +
 * b2) + b3 for n rockets in one launch, activations kept on chip (tcgen05 tensor cores, bf16 weights,
 * fp32 accumulation in tensor memory).  `blob` is the packed net in DEVICE memory (16-byte aligned), RL_MLP_BLOB_BYTES:
 *   bf16 [256][16]  W1[j][0..7] twice (the kernel feeds the observation as bf16 high | low parts)
 *   f32  [256]      b1
 *   bf16 [32][256][8] W2 in the tensor core's K-major tile order: element (n, k) of nn.Linear.weight
 *                   (n = output unit) at [k / 8][n][k % 8] -- the image of the kernel's shared-memory operand
 *   f32  [256]      b2
 *   f32  [2][256]   W3 (second row ignored when out_dim = 1)
 *   f32  [4]        b3 (first out_dim used)
 * obs float32 [n][8] (16-byte aligned), out float32 [n][out_dim], out_dim 1 or 2.  DEVICE pointers,
 * enqueue only.
 *
 * rl_policy_sample: the policy net (out_dim 2) with SB3's DiagGaussian sampling fused into its
 * epilogue (what OnPolicyAlgorithm.collect_rollouts does around policy.forward): actions[n][2] =
 * mean + exp(log_std) eps, eps ~ N(0, 1) (unclipped: what the rollout buffer stores); clipped[n][2]
 * = actions clipped to the Box [low2, high2] (what env.step receives); log_prob[n]; mean[n][2]
 * optional (nullable).  eps is Philox4x32-10 at counter (row_offset + row, draw) + Box-Muller, keyed by
 * `seed` xor a domain tag -- so the stream is independent of the environment's reset stream (key =
 * seed, counter = (global id, epoch)) even for equal seeds: pass a different `draw` on every call and
 * the shard's global id offset (rl_create's env_id_offset) as `row_offset`, so that ranks sharing a
 * seed draw different noise.  log_std is a DEVICE pointer to 2 floats; low2 / high2 are HOST
 * pointers to 2 floats.
 *
 * rl_policy_value_sample: rl_policy_sample(blob_pi, ...) and values[n] = rl_mlp_forward(blob_vf, obs, ..., 1)
 * in ONE launch (what a rollout step asks of the actor-critic, scripts/train.py:117-134: separate pi and vf
 * nets on the same observation): some CTAs run the policy net, the others the value net.  Results are
 * bit-identical to the two separate calls. */
#define RL_MLP_HIDDEN 256
#define RL_MLP_BLOB_BYTES 143376
const char *rl_policy_last_error(void);
int64_t rl_mlp_blob_bytes(void);
int rl_mlp_forward(const void *blob, const float *obs, float *out, int64_t n, int out_dim, void *stream);
int rl_policy_sample(const void *blob, const float *obs, const float *log_std, uint64_t seed, uint64_t draw,
                     int64_t row_offset, const float *low2, const float *high2, float *actions, float *clipped,
                     float *log_prob, float *mean, int64_t n, void *stream);
int rl_policy_value_sample(const void *blob_pi, const void *blob_vf, const float *obs, const float *log_std,
                           uint64_t seed, uint64_t draw, int64_t row_offset, const float *low2, const float *high2,
                           float *actions, float *clipped, float *log_prob, float *mean, float *values, int64_t n,
                           void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ROCKET_LANDING_B200_H_ */
