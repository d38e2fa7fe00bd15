/* sng.h -- C ABI of the B200-native batched smart-nanogrid environment step.
 *
 * This is the drop-in boundary for the reference's hot path.  The reference has no native
 * interface (it is pure Python); each entry point below names the Python method(s) of
 * Dellintel98/smart-nanogrid-gym it replaces for E environments at once
 * (paths relative to the reference tree, smart_nanogrid_gym/...):
 *
 *   sng_create          SmartNanogridEnv.__init__             envs/smart_nanogrid_environment.py:32-120
 *                       (+ CentralManagementSystem.__init__   utils/central_management_system.py:11-43)
 *   sng_reset           SmartNanogridEnv.reset                envs/smart_nanogrid_environment.py:311-351
 *                       (+ ChargingStation.generate_new_initial_values  utils/charging_station.py:152-279)
 *   sng_load_schedule   ChargingStation.load_initial_values   utils/charging_station.py:119-136
 *   sng_step            SmartNanogridEnv.step                 envs/smart_nanogrid_environment.py:140-188
 *                       (+ CentralManagementSystem.manage_nanogrid  utils/central_management_system.py:84-185)
 *   sng_step_host       the same call made with host (numpy) arrays, as a gym caller does
 *   sng_rollout         n consecutive step() calls of a trainer's rollout loop   solvers/RL/ppo_train.py:94-101
 *   sng_sample_plan     the `initial_values.json` dump        utils/charging_station.py:173-186
 *   sng_sample_actions  env.action_space.sample() per env     envs/smart_nanogrid_environment.py:101-118
 *   sng_error_flags     the reference's `raise ValueError` sites
 *                       (central_management_system.py:158-159, penaliser.py:111)
 *   sng_gae             (next row of the path) stable_baselines3 RolloutBuffer.compute_returns_and_advantage,
 *                       the consumer of the rollouts collected by solvers/RL/ppo_train.py:94-101
 *
 * Conventions: plain pointers and sizes only.  All device buffers are owned by the caller
 * (torch) and merely borrowed between sng_bind and sng_destroy.  Every call is asynchronous
 * on the given CUDA stream (a cudaStream_t passed as void*) unless stated otherwise; there are
 * no hidden allocations or synchronisations in sng_step / sng_rollout.  Return value 0 = OK,
 * negative = error; sng_last_error() gives the message (thread local).  No C++ exception
 * crosses the boundary.  A handle may be used by one host thread at a time.
 */
#ifndef SNG_H
#define SNG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNG_ABI_VERSION 3
#define SNG_MAX_VEHICLES 8   /* schedule slots per spot and day */
#define SNG_MAX_SPOTS 255
#define SNG_MAX_TABLE 512    /* entries of the shared PV / price tables (two days; pv_days + 1 days with multi-day PV) */

typedef struct sng_env sng_env;

enum { SNG_OK = 0, SNG_ERR_ARG = -1, SNG_ERR_CUDA = -2, SNG_ERR_STATE = -3, SNG_ERR_UNSUPPORTED = -4 };
enum { SNG_PEN_NONE = 0, SNG_PEN_ON_DEPARTURE = 1, SNG_PEN_SPARSE = 2, SNG_PEN_DENSE = 3 };
enum { SNG_F32 = 32, SNG_F64 = 64 };

/* sticky per-env error bits (the reference raises instead; kernels never trap) */
enum {
    SNG_FLAG_NEG_DEMAND = 1u,  /* total EV power < 0 without V2X: central_management_system.py:158-159 */
    SNG_FLAG_BATT_SOC_GT1 = 2u, /* penaliser.py:111 */
    SNG_FLAG_NAN_ACTION = 4u
};

/* Mirrors the reference constructor arguments and its hard-coded physical constants. */
typedef struct {
    uint32_t struct_size;   /* = sizeof(sng_config) */
    int32_t precision;      /* SNG_F32 (production) or SNG_F64 (validation build, bit-faithful arithmetic) */
    int64_t n_envs;         /* environments owned by this handle (this GPU's slice) */
    int64_t env_gid0;       /* global id of local env 0: RNG streams are keyed by global id */
    int32_t n_spots;        /* number_of_chargers */
    int32_t n_steps;        /* 24 / time_interval */
    int32_t horizon;        /* NUMBER_OF_HOURS_AHEAD = 3 */
    int32_t table_len;      /* entries of each table below (>= n_steps + horizon) */
    int32_t pv;             /* pv_system_available_in_model */
    int32_t batt;           /* battery_system_available_in_model */
    int32_t v2x;            /* vehicle_to_everything */
    int32_t penalty_mode;   /* SNG_PEN_* (vehicle_uncharged_penalty_mode) */
    int32_t diff_cap;       /* enable_different_vehicle_battery_capacities */
    int32_t req_soc;        /* enable_requested_state_of_charge */
    int32_t default_cap;    /* 40 kWh */
    int32_t auto_reset;     /* 1: a finished env is reset inside the same step (VecEnv semantics) */
    int32_t pv_days;        /* 1 (the reference: NUMBER_OF_DAYS_TO_PREDICT = 1, ...environment.py:51, and row 0 of
                             * solar_irradiance_2 only, pv_system_manager.py:81-91).  D > 1: episode k reads the PV tables
                             * at offset (k % D) * n_steps, i.e. day k % D of a (D + 1)-day series; table_len must then
                             * be >= (D + 1) * n_steps */
    int32_t _reserved;
    double dt;              /* hours per step */
    double ev_pmax, ev_eff; /* 22 kW, 0.95 */
    double b_cap, b_pmax, b_eff, b_dod, b_soc0; /* 80 kWh, 44 kW, 0.95, 0.15, 0.5 */
    double sell_coeff, cost_weight, batt_pen_w, margin, dep_norm; /* 0.8, 0.75, 0.8, 0.05, 24 */
    const double *pv_power;   /* host pointers, table_len entries, copied by sng_create */
    const double *irr_norm;
    const double *price;
    const double *price_norm;
} sng_config;

/* Byte sizes of the opaque per-env state arrays for a given config (caller allocates). */
typedef struct {
    uint32_t struct_size;
    int32_t act_dim, obs_dim;
    int32_t real_bytes;     /* 4 or 8: element size of actions / reward / soc / req */
    int32_t plan_rec_bytes; /* one planned-vehicle record of `plan` (12 or 24) */
    int32_t envst_bytes;    /* one per-env scalar block (16 or 32) */
    int32_t plan_slots;     /* SNG_MAX_VEHICLES */
    int32_t diag_count;     /* reals per env in the optional diagnostics row */
    int32_t env_block;      /* 32: the per-spot state array is blocked by 32 envs (see sng_buffers.spot) */
    int32_t spot_planes;    /* 3: planes of the per-spot state (header word, SoC, requested SoC) */
} sng_layout;

/* Device buffers.  `real` = float (SNG_F32) or double (SNG_F64).  Optional pointers may be NULL.
 * The per-spot state `spot` is ONE array of real-sized words, a structure of arrays, plane-major and
 * blocked by env_block = 32 envs: word (env e, spot i, plane f) lives at index
 *     f * (ceil(E / 32) * n_spots * 32) + ((e / 32) * n_spots + i) * 32 + e % 32
 * with plane 0 = header word (arrival | departure << 8 | capacity << 16 | next arrival << 24, zero
 * extended), plane 1 = SoC column the next step starts from, plane 2 = requested SoC (not maintained
 * while every vehicle requests 1.0, i.e. sampled schedules without enable_requested_state_of_charge)
 * (both `real` bit patterns).  A 32-env block's n_spots lines of one plane are contiguous.  The caller
 * allocates 3 * ceil(E / 32) * n_spots * 32 words of real_bytes
 * bytes. */
typedef struct {
    uint32_t struct_size;
    uint32_t _pad;
    const void *actions;    /* in   [E][act_dim] real, row-major, dense */
    float *obs;             /* out  [E][obs_dim] float32 (the reference casts obs to float32) */
    void *reward;           /* out  [E] real */
    uint8_t *done;          /* out  [E] terminated flag (truncated is always 0 in the reference) */
    float *terminal_obs;    /* out, optional [E][obs_dim]: last obs of the finished episode (auto_reset) */
    void *spot;             /* state, blocked words (see above) */
    void *envst;            /* state [E] x envst_bytes: battery SoC, pv_shift, episode return, (episode, t) */
    void *plan;             /* optional [E][n_spots][SNG_MAX_VEHICLES] x plan_rec_bytes: full-day schedule */
    uint32_t *err;          /* optional [E] sticky SNG_FLAG_* bits */
    void *diag;             /* optional [E][diag_count] real per-step diagnostics */
    void *last_return;      /* optional [E] real: return of the most recently finished episode */
    void *spot_power;       /* optional [E][n_spots] real per-step diagnostics: power of every charging spot
                             * (charger_power_values, utils/charging_station.py:283-300; logged by the
                             * reference's step as 'Charger_power_values', envs/smart_nanogrid_environment.py:161) */
} sng_buffers;

enum { /* diagnostics row, subset of central_management_system.py:128-155 */
    SNG_D_TOTAL_CH = 0, SNG_D_TOTAL_DIS, SNG_D_SOLAR, SNG_D_BATT_POWER, SNG_D_GRID_POWER,
    SNG_D_GRID_COST, SNG_D_PEN_VEH, SNG_D_PEN_BATT, SNG_D_COUNT
};

/* Host view of a schedule to replay (compact per-vehicle records). */
typedef struct {
    uint32_t struct_size;
    int32_t n_slots;          /* V <= SNG_MAX_VEHICLES */
    const int32_t *arr;       /* [E][N][V] arrival step */
    const int32_t *dep;       /* [E][N][V] departure step */
    const int32_t *cap;       /* [E][N][V] capacity, integer kWh */
    const double *soc0;       /* [E][N][V] arrival SoC */
    const double *req;        /* [E][N][V] requested SoC */
    const int32_t *n_veh;     /* [E][N] */
    const double *pv_shift;   /* optional [E] */
    const double *soc_b;      /* optional [E] battery SoC to start from */
} sng_schedule_view;

int sng_abi_version(void);
/* sizeof() of the ABI structs as compiled: 0 config, 1 layout, 2 buffers, 3 schedule_view (binding self-check). */
int sng_sizeof(int which);
const char *sng_last_error(void);

int sng_query_layout(const sng_config *cfg, sng_layout *out);
int sng_create(const sng_config *cfg, int device, sng_env **out);
void sng_destroy(sng_env *env);
int sng_bind(sng_env *env, const sng_buffers *buffers);

/* Start new episodes (sampling mode).  mask: optional DEVICE pointer [E] (1 = reset this env).
 * reset_battery != 0 also sets the battery SoC to b_soc0 (the reference never does after
 * construction, quirk Q8).  Writes the reset observation of the selected envs to `obs`.
 * A masked reset leaves the handle-wide settings alone: it needs sampling mode (not a replayed schedule) and
 * the seed of the last full reset, otherwise SNG_ERR_STATE. */
int sng_reset(sng_env *env, uint64_t seed, const uint8_t *mask, int reset_battery, void *stream);

/* Replay mode: upload a schedule (host arrays), rewind to t = 0 and write the reset obs.
 * Synchronous w.r.t. the host arrays. */
int sng_load_schedule(sng_env *env, const sng_schedule_view *view, void *stream);

/* One env.step() for all envs: reads `actions`, updates state, writes obs/reward/done. */
int sng_step(sng_env *env, void *stream);

/* n_steps consecutive steps in one launch.  actions [n_steps][E][act_dim], obs [n_steps][E][obs_dim],
 * reward [n_steps][E], done [n_steps][E] (device, all required). */
int sng_rollout(sng_env *env, const void *actions, float *obs, void *reward, uint8_t *done, int n_steps,
                void *stream);

/* The gym-facing call with HOST buffers (pinned recommended): H2D actions, step, D2H results,
 * then waits for completion.  Sizes as in sng_buffers. */
int sng_step_host(sng_env *env, const void *actions_host, float *obs_host, void *reward_host,
                  uint8_t *done_host, void *stream);

/* Eagerly generate the whole-day schedule of the CURRENT episode of every env into `plan`
 * (same Philox streams the lazy in-step sampler uses), for export / inspection. */
int sng_sample_plan(sng_env *env, void *stream);

/* The random policy (BASELINE config 2; reference: env.action_space.sample() per env and step, bounds of
 * envs/smart_nanogrid_environment.py:101-118): fills actions [n_steps][E][act_dim] (device, the env's real type) with
 * low + u * (high - low), u uniform in [0, 1) with 24 bits, from Philox4x32-10 keyed by `seed` with counter (global env id,
 * step0 + s, column group): the draw of env e at step s does not depend on how the batch is sharded (env_gid0), on n_steps or
 * on the batch size, and is disjoint from the schedule sampler's streams even for the same seed.  step0 + n_steps < 2^32.
 * Feed the slab to sng_rollout (or a slice to sng_step): SURVEY's `sng_rollout(env, NULL, ...)` with the buffer caller-owned. */
int sng_sample_actions(sng_env *env, uint64_t seed, uint64_t step0, int n_steps, void *actions, void *stream);

/* OR of all per-env error flags (synchronises the stream). */
int sng_error_flags(sng_env *env, uint32_t *host_out, void *stream);

/* Next row of the path: advantages / returns of a collected rollout, for the trainer that
 * drives the step (solvers/RL/ppo_train.py:94-101 -> stable_baselines3 PPO.collect_rollouts ->
 * RolloutBuffer.compute_returns_and_advantage).  All arrays on the device; rewards, values, advantages,
 * returns [n_steps][E] float32, episode_starts [n_steps][E] u8 (1 = the env was reset before that step),
 * last_values [E] = V(obs after the last step), last_dones [E].  Asynchronous on `stream`. */
int sng_gae(const float *rewards, const float *values, const uint8_t *episode_starts, const float *last_values,
            const uint8_t *last_dones, float *advantages, float *returns, int n_steps, int64_t n_envs, float gamma,
            float gae_lambda, void *stream);

/* Fused actor-critic forward pass of Stable-Baselines3's default MlpPolicy (two tanh 64-64 networks, linear
 * action head with state-independent log-std, linear value head; what PPO("MlpPolicy", env) at
 * solvers/RL/ppo_train.py:89-92 builds), for rollout collection around the step: one launch computes values,
 * sampled actions (mean + noise * exp(log_std), `noise` ~ N(0,1) supplied by the caller, NULL = deterministic), the
 * actions clipped to the Box (what env.step receives) and their log-probabilities.  All pointers are device
 * pointers to float32, weights row-major [out][in] as torch.nn.Linear stores them.  actions == NULL computes the
 * values only.  Supported shapes: hidden = 64, the observation / action sizes of 4-, 8- and 10-spot stations;
 * otherwise SNG_ERR_UNSUPPORTED (callers fall back to their own framework).  Asynchronous on `stream`. */
typedef struct {
    uint32_t struct_size;
    int32_t obs_dim, hidden, act_dim;
    const float *w_pi0, *b_pi0, *w_pi1, *b_pi1;   /* actor:  [64][obs_dim], [64], [64][64], [64] */
    const float *w_act, *b_act, *log_std;         /* action head [act_dim][64], [act_dim]; log-std [act_dim] */
    const float *w_vf0, *b_vf0, *w_vf1, *b_vf1;   /* critic: [64][obs_dim], [64], [64][64], [64] */
    const float *w_val, *b_val;                   /* value head [1][64], [1] */
} sng_mlp;
int sng_policy_forward(const sng_mlp *mlp, const float *obs, const float *noise, const float *low, const float *high,
                       float *raw_actions, float *actions, float *values, float *log_probs, int64_t n_envs,
                       void *stream);

/* The same forward pass on the 5th-generation tensor cores (tcgen05.mma, activations in tensor memory, FP32 kept
 * by a 3-term tf32 split; sng_policy_tc.cu): the production path of rollout collection.  The weights are split and
 * laid out for the tensor cores once per weight update by sng_policy_pack into a caller-owned device buffer of
 * sng_policy_packed_bytes() bytes (16-byte aligned); sng_policy_forward_packed then takes that image instead of the
 * nn.Linear tensors.  Arguments otherwise as for sng_policy_forward; obs_dim <= 30, act_dim <= 16, hidden = 64. */
size_t sng_policy_packed_bytes(void);
int sng_policy_pack(const sng_mlp *mlp, void *packed, void *stream);
int sng_policy_forward_packed(const void *packed, int obs_dim, int act_dim, const float *obs, const float *noise,
                              const float *low, const float *high, float *raw_actions, float *actions, float *values,
                              float *log_probs, int64_t n_envs, void *stream);

/* The same forward pass with the exploration noise drawn INSIDE the kernel (no noise tensor, no separate RNG launch per
 * rollout step): z ~ N(0, 1) per (env, action) from Philox4x32-10 keyed by (seed, step), counter = (env_gid0 + env,
 * action / 4) and Box-Muller, so a sharded rollout draws the same noise as an unsharded one.  step = *step_counter (a
 * DEVICE word, so that a captured CUDA graph advances between replays: add the rollout length to it after each rollout)
 * + step_offset (the position inside the rollout).  noise_out (optional, [E][act_dim]) receives the z that were used.
 * Reference caller: SB3's DiagGaussianDistribution.sample() inside PPO.collect_rollouts (solvers/RL/ppo_train.py:89-102). */
int sng_policy_forward_sampled(const void *packed, int obs_dim, int act_dim, const float *obs, uint64_t seed,
                               const uint64_t *step_counter, uint64_t step_offset, uint64_t env_gid0, const float *low,
                               const float *high, float *raw_actions, float *actions, float *values, float *log_probs,
                               float *noise_out, int64_t n_envs, void *stream);

/* ONE launch for a whole rollout step: the forward pass of sng_policy_forward_packed / _sampled over `env`'s batch FUSED
 * with sng_step.  The io warps of the tensor-core kernel (thread = row = env, one warp = one 32-env state block: the step
 * kernel's own mapping) run the env step of their tile as soon as its clipped actions are in shared memory, so the
 * actions never travel through HBM to a second kernel and one kernel boundary per rollout step disappears.  Same
 * env-step body as sng_step: observations, rewards, done flags and the handle's state are bit-identical to the two
 * separate launches.  obs [E][obs_dim] is what the policy sees (the current observation), obs_next / reward / done
 * receive the step's results (SB3: collect_rollouts' policy(obs) -> clip -> env.step, solvers/RL/ppo_train.py:89-102).
 * noise = NULL draws the exploration noise in the kernel (seed, step_counter, step_offset as for
 * sng_policy_forward_sampled; the handle's env_gid0 keys it).  Available for the reference's default station (PV with
 * 3 steps ahead, battery, no requested-SoC plane) at 4 and 10 spots, float32, batches that are a multiple of 128 envs
 * and 16-byte aligned buffers; otherwise SNG_ERR_UNSUPPORTED (launch the two kernels separately). */
int sng_policy_step(sng_env *env, const void *packed, const float *obs, const float *noise, uint64_t seed,
                    const uint64_t *step_counter, uint64_t step_offset, const float *low, const float *high,
                    float *raw_actions, float *actions, float *values, float *log_probs, float *noise_out,
                    float *obs_next, void *reward, uint8_t *done, void *stream);

/* Launches an EMPTY kernel on `stream`: lets a caller measure the device's kernel-to-kernel launch latency (the
 * floor under one sng_step per step for batches that live in L2; bench.py reports it beside those numbers). */
int sng_null_launch(void *stream);

/* Test hook: evaluates the step kernels' arrival-gap function (the number of failed Bernoulli(0.4) arrival trials a
 * 32-bit Philox word encodes: charging_station.py:213-214 sampled per vehicle) on n device words. */
int sng_debug_arrival_gap(sng_env *env, const uint32_t *x, uint32_t *gap, int64_t n, void *stream);

/* Measurement hook: ONE launch that performs the step kernel's memory traffic without its arithmetic (same geometry,
 * occupancy, loads and stores; default 10-spot station, whole 32-env blocks).  Its duration is what the access pattern
 * alone costs: the practical ceiling bench.py reports under the step kernel beside the copy-bandwidth roofline.  The
 * handle's state is written back unchanged; obs / reward / done receive meaningless values.  variant 0 = the step
 * kernel's own pattern; what-if patterns with the same byte counts: 1 the state planes interleaved per spot (the layout of
 * rounds 1-2), 2 loads only, 3 stores only (overwrites the state: reset afterwards), 4 observation rows through the copy
 * engine, 5 like 1 with two planes per spot. */
int sng_debug_traffic_skeleton(sng_env *env, int variant, void *stream);

/* Measurement hook: a one-thread kernel that writes the device's %globaltimer (ns) to *slot.  Capturable in a CUDA graph:
 * stamps between the kernels of a captured loop give each kernel's span on one clock (scripts/rollout_timeline.py). */
int sng_debug_stamp(uint64_t *slot, void *stream);

/* Kernels launched by this handle so far (bench.py's gpu_launches claim). */
int64_t sng_launch_count(const sng_env *env);

/* Tuning knobs for experiments and tests: warps (= blocks of 32 envs) per CTA (0 = auto); force the
 * generic runtime-N kernel instead of the specialised one; how action / observation rows are staged
 * through shared memory when the buffers are 16-byte aligned: -1 scalar loads / stores, 0 coalesced
 * 16-byte vector loads / stores, 1 (default) copy-engine (cp.async.bulk) loads + vector stores -- copy
 * engine both ways for stations of more than 32 spots --, 3 copy engine both ways (unaligned buffers
 * and a partial last block always take the scalar path); number of env
 * chunks sng_step_host pipelines over PCIe (0 = auto). */
int sng_set_tuning(sng_env *env, int warps_per_cta, int use_generic_kernel, int use_bulk_copy, int host_chunks);
/* How sng_step launches its kernel.  0 (default): an ordinary launch.  1: programmatic dependent launch -- the kernel may
 * start (block scheduling, parameter and index set-up) while its predecessor in the stream is still running and waits for
 * it (griddepcontrol.wait) before it touches memory; hides the kernel-to-kernel launch latency of small batches.
 * 2: the same, and the per-env state is loaded BEFORE the wait: only valid when the predecessor in the stream does not
 * write this handle's state (e.g. the policy kernel of a rollout loop; NOT another sng_step of the same handle). */
int sng_set_launch_mode(sng_env *env, int mode);
/* The same switch for sng_policy_forward_packed / _sampled (process-wide), a bit mask.  Bit 0: programmatic dependent
 * launch -- CTA set-up and the fetch of the packed weight image run while the predecessor in the stream (the step kernel
 * that writes the observations) is draining; only valid when the predecessor does not write the weight image (i.e. not
 * directly behind sng_policy_pack).  Bit 1: every CTA claims its SM's whole shared memory, so that no CTA of a kernel
 * launched early behind this one can become resident beside it. */
int sng_policy_set_launch_mode(int mode);

/* Tuning knob: which step kernel runs: 0 (default) one 32-env block per warp, with four lanes per env for
 * specialised stations of more than 32 spots; 1 the persistent software-pipelined kernel (measured slower,
 * see DESIGN.md); 2 one block per warp and always one lane per env; 3 like 0 with two lanes per env.
 * Latency-bound batches: with 0, small batches of the reference's default station shape (PV, 3 steps ahead, battery; 4, 8
 * or 10 spots; float32; sampled or replayed schedules; at most 4,096 envs per sng_step launch, 6,144 per sng_rollout launch)
 * run the one-lane-per-SPOT kernel (two envs per warp, station sums by shuffles), and 10-spot batches of whole 32-env blocks
 * up to 16,384 / 65,536 envs (sampled, no requested SoC) run with two lanes per env -- bit-identical results either way; 3 also means two lanes per env for
 * such 10-spot batches of any size; 4 = the one-lane-per-spot kernel at every batch size; 5 = like 0 without either form.
 * ctas_per_sm caps the resident CTAs per SM of the pipelined kernel (0 = as many as fit). */
int sng_set_pipeline(sng_env *env, int kernel_variant, int ctas_per_sm);

#ifdef __cplusplus
}
#endif
#endif /* SNG_H */
