// sng_api.cu -- the C ABI declared in include/sng.h, plus the float32 production engine.
#include "sng_engine.cuh"

namespace sng {
EngineBase *make_engine_f32(const sng_config &cfg, int device, std::string &err)
{
    auto *e = new Engine<float, false>();
    if (e->init(cfg, device) != SNG_OK) {
        err = e->error;
        delete e;
        return nullptr;
    }
    return e;
}
}  // namespace sng

// Generalised advantage estimation over a rollout, one thread per env walking the steps backwards
// (every access of a warp is one coalesced line of the [n_steps][E] slabs).  Restates
// stable_baselines3 RolloutBuffer.compute_returns_and_advantage, the consumer of the rollouts the
// reference's trainer collects (solvers/RL/ppo_train.py:94-101).
__global__ void __launch_bounds__(256) gae_kernel(const float *__restrict__ rewards, const float *__restrict__ values,
                                                 const uint8_t *__restrict__ episode_starts,
                                                 const float *__restrict__ last_values,
                                                 const uint8_t *__restrict__ last_dones, float *__restrict__ advantages,
                                                 float *__restrict__ returns, int n_steps, long long n_envs, float gamma,
                                                 float lam)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_envs) return;
    float next_value = last_values[e];
    float next_non_terminal = last_dones[e] ? 0.0f : 1.0f;
    float gae = 0.0f;
    for (int t = n_steps - 1; t >= 0; --t) {
        const size_t k = (size_t)t * (size_t)n_envs + (size_t)e;
        const float v = values[k];
        const float delta = rewards[k] + gamma * next_value * next_non_terminal - v;
        gae = delta + gamma * lam * next_non_terminal * gae;
        advantages[k] = gae;
        returns[k] = gae + v;
        next_value = v;
        next_non_terminal = episode_starts[k] ? 0.0f : 1.0f;
    }
}

static __global__ void null_kernel() {}
static __global__ void stamp_kernel(unsigned long long *slot)
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *slot = t;
}

struct sng_env {
    sng::EngineBase *eng;
};

static thread_local std::string g_last_error;

static int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

static int done(sng_env *env, int rc)
{
    if (rc != SNG_OK) g_last_error = env->eng->error;
    return rc;
}

extern "C" {

int sng_abi_version(void) { return SNG_ABI_VERSION; }

int sng_sizeof(int which)
{
    switch (which) {
    case 0: return (int)sizeof(sng_config);
    case 1: return (int)sizeof(sng_layout);
    case 2: return (int)sizeof(sng_buffers);
    case 3: return (int)sizeof(sng_schedule_view);
    default: return -1;
    }
}

const char *sng_last_error(void) { return g_last_error.c_str(); }

int sng_query_layout(const sng_config *cfg, sng_layout *out)
{
    if (!cfg || !out || cfg->struct_size != sizeof(sng_config)) return fail(SNG_ERR_ARG, "sng_query_layout: bad struct_size");
    if (cfg->precision != SNG_F32 && cfg->precision != SNG_F64) return fail(SNG_ERR_ARG, "precision must be 32 or 64");
    const bool f64 = cfg->precision == SNG_F64;
    const int pv = cfg->pv != 0, b = cfg->batt != 0;
    out->struct_size = sizeof(sng_layout);
    out->act_dim = cfg->n_spots + b;
    out->obs_dim = (1 + pv) * (1 + cfg->horizon) + 2 * cfg->n_spots + b;
    out->real_bytes = f64 ? 8 : 4;
    out->plan_rec_bytes = f64 ? (int)sizeof(sng::PlanRec<double>) : (int)sizeof(sng::PlanRec<float>);
    out->envst_bytes = f64 ? (int)sizeof(sng::EnvSt<double>) : (int)sizeof(sng::EnvSt<float>);
    out->plan_slots = SNG_MAX_VEHICLES;
    out->diag_count = SNG_D_COUNT;
    out->env_block = sng::kBlock;
    out->spot_planes = sng::kPlanes;
    return SNG_OK;
}

int sng_create(const sng_config *cfg, int device, sng_env **out)
{
    if (!cfg || !out || cfg->struct_size != sizeof(sng_config)) return fail(SNG_ERR_ARG, "sng_create: bad struct_size");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1)
        return fail(SNG_ERR_CUDA, "sng_create: no CUDA device (this library has no CPU fallback)");
    if (device < 0 || device >= n_dev) return fail(SNG_ERR_ARG, "sng_create: bad device index");
    std::string err;
    sng::EngineBase *eng = nullptr;
    if (cfg->precision == SNG_F32) eng = sng::make_engine_f32(*cfg, device, err);
    else if (cfg->precision == SNG_F64) eng = sng::make_engine_f64(*cfg, device, err);
    else return fail(SNG_ERR_ARG, "sng_create: precision must be 32 or 64");
    if (!eng) return fail(SNG_ERR_ARG, err.empty() ? "sng_create failed" : err);
    *out = new sng_env{eng};
    return SNG_OK;
}

void sng_destroy(sng_env *env)
{
    if (!env) return;
    delete env->eng;
    delete env;
}

#define SNG_ENV_CHECK(env) \
    if (!(env) || !(env)->eng) return fail(SNG_ERR_ARG, "null environment handle")

int sng_bind(sng_env *env, const sng_buffers *buffers)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->bind(buffers));
}

int sng_reset(sng_env *env, uint64_t seed, const uint8_t *mask, int reset_battery, void *stream)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->reset(seed, mask, reset_battery, (cudaStream_t)stream));
}

int sng_load_schedule(sng_env *env, const sng_schedule_view *view, void *stream)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->load_schedule(view, (cudaStream_t)stream));
}

int sng_step(sng_env *env, void *stream)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->step((cudaStream_t)stream));
}

int sng_rollout(sng_env *env, const void *actions, float *obs, void *reward, uint8_t *done_, int n_steps, void *stream)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->rollout(actions, obs, reward, done_, n_steps, (cudaStream_t)stream));
}

int sng_step_host(sng_env *env, const void *actions_host, float *obs_host, void *reward_host, uint8_t *done_host,
                  void *stream)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->step_host(actions_host, obs_host, reward_host, done_host, (cudaStream_t)stream));
}

int sng_sample_plan(sng_env *env, void *stream)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->sample_plan((cudaStream_t)stream));
}

int sng_sample_actions(sng_env *env, uint64_t seed, uint64_t step0, int n_steps, void *actions, void *stream)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->sample_actions(seed, step0, n_steps, actions, (cudaStream_t)stream));
}

int sng_error_flags(sng_env *env, uint32_t *host_out, void *stream)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->error_flags(host_out, (cudaStream_t)stream));
}

int sng_null_launch(void *stream)
{
    null_kernel<<<1, 32, 0, (cudaStream_t)stream>>>();
    return cudaGetLastError() == cudaSuccess ? SNG_OK : fail(SNG_ERR_CUDA, "sng_null_launch failed");
}

int sng_debug_arrival_gap(sng_env *env, const uint32_t *x, uint32_t *gap, int64_t n, void *stream)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->probe_arrival_gap(x, gap, (long long)n, (cudaStream_t)stream));
}

int64_t sng_launch_count(const sng_env *env) { return (env && env->eng) ? env->eng->launches : 0; }

int sng_set_tuning(sng_env *env, int warps_per_cta, int use_generic_kernel, int use_bulk_copy, int host_chunks)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->set_tuning(warps_per_cta, use_generic_kernel, use_bulk_copy, host_chunks));
}

int sng_gae(const float *rewards, const float *values, const uint8_t *episode_starts, const float *last_values,
            const uint8_t *last_dones, float *advantages, float *returns, int n_steps, int64_t n_envs, float gamma,
            float gae_lambda, void *stream)
{
    if (!rewards || !values || !episode_starts || !last_values || !last_dones || !advantages || !returns || n_steps < 1 ||
        n_envs < 1)
        return fail(SNG_ERR_ARG, "sng_gae: bad arguments");
    int prev = -1, dev = -1;                       // launch on the device that owns the buffers
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, rewards) == cudaSuccess && at.type == cudaMemoryTypeDevice) dev = at.device;
    const bool switched = dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev && cudaSetDevice(dev) == cudaSuccess;
    gae_kernel<<<(unsigned)((n_envs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        rewards, values, episode_starts, last_values, last_dones, advantages, returns, n_steps, (long long)n_envs, gamma,
        gae_lambda);
    const cudaError_t e = cudaGetLastError();
    if (switched) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(SNG_ERR_CUDA, std::string("sng_gae: ") + cudaGetErrorString(e));
    return SNG_OK;
}

int sng_policy_step(sng_env *env, const void *packed, const float *obs, const float *noise, uint64_t seed,
                    const uint64_t *step_counter, uint64_t step_offset, const float *low, const float *high,
                    float *raw_actions, float *actions, float *values, float *log_probs, float *noise_out,
                    float *obs_next, void *reward, uint8_t *done_flags, void *stream)
{
    SNG_ENV_CHECK(env);
    sng::PolicyStepArgs a;
    a.packed = packed; a.obs = obs; a.noise = noise; a.step_counter = step_counter; a.step_offset = step_offset; a.seed = seed;
    a.low = low; a.high = high; a.raw_actions = raw_actions; a.actions = actions; a.values = values; a.log_probs = log_probs;
    a.noise_out = noise_out; a.obs_next = obs_next; a.reward = (float *)reward; a.done = done_flags;
    return done(env, env->eng->policy_step(a, (cudaStream_t)stream));
}

int sng_debug_stamp(uint64_t *slot, void *stream)
{
    if (!slot) return fail(SNG_ERR_ARG, "sng_debug_stamp: null slot");
    stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long *>(slot));
    return cudaGetLastError() == cudaSuccess ? SNG_OK : fail(SNG_ERR_CUDA, "sng_debug_stamp: launch failed");
}

int sng_debug_traffic_skeleton(sng_env *env, int variant, void *stream)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->traffic_skeleton(variant, (cudaStream_t)stream));
}

int sng_set_launch_mode(sng_env *env, int mode)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->set_launch_mode(mode));
}

int sng_set_pipeline(sng_env *env, int kernel_variant, int ctas_per_sm)
{
    SNG_ENV_CHECK(env);
    return done(env, env->eng->set_pipeline(kernel_variant, ctas_per_sm));
}

}  // extern "C"
