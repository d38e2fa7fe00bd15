// sng_policy.cu -- fused actor-critic forward pass for the rollout-collection row.
//
// The reference trains with Stable-Baselines3 PPO("MlpPolicy") (solvers/RL/ppo_train.py:89-92): two
// separate tanh 64-64 networks (actor, critic), a linear action head with a state-independent log-std
// (DiagGaussian) and a linear value head; collect_rollouts clips the sampled actions to the Box before
// env.step.  With the step kernel at a few microseconds per 65,536-env step, ~20 small library launches per
// policy call dominate rollout collection, so this kernel does the whole forward pass -- both networks, the
// heads, sampling with supplied N(0,1) noise, clipping, log-probabilities -- in one launch, one thread per env:
// weights in shared memory (broadcast LDS.128), the first hidden layer in registers, the second hidden layer
// consumed by the heads as it is produced.  FP32 throughout (CUDA cores: the matrices are 29x64 / 64x64 per
// env; tensor cores would need TF32/BF16 inputs and change the numerics of a trainer expecting FP32).
#include <cmath>
#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <utility>

#include <cuda_runtime.h>

#include "../../include/sng.h"

namespace {

constexpr int H = 64;          // hidden width of SB3's default MlpPolicy
constexpr int kThreads = 128;

// tanh(x) = 1 - 2 / (exp(2x) + 1) on the special-function unit (EX2 + RCP): absolute error ~1e-7, saturates
// correctly for large |x|.  libm's tanhf costs ~30 instructions with branches -- as much as a whole layer-0 row.
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

// Each thread carries NE envs: the kernel is bound by the broadcast LDS.128 reads of the weights (one per four
// FMAs with one env per thread), so every weight vector fetched from shared memory is used for NE envs.
// The multiply-adds are Blackwell's packed FP32 pairs (__ffma2_rn, SASS FFMA2): a pair accumulates the even-k and
// the odd-k half of a dot product, so the weight pairs come straight out of the float4 loads and the activation
// pairs are adjacent registers -- half the FMA instructions, no packing moves.
//
// y[n][j] = tanh(b0[j] + sum_k w0[j][k] x[n][k]) for j < H; w0 rows padded to DP floats in shared memory;
// x and y are held as (even, odd) pairs.
template <int DP, int NE>
__device__ __forceinline__ void layer0(const float *w0, const float *b0, const float2 (&x)[NE][DP / 2], float2 (&h)[NE][H / 2])
{
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float2 acc[NE];
#pragma unroll
        for (int n = 0; n < NE; ++n) acc[n] = make_float2(b0[j], 0.f);
        const float4 *w = reinterpret_cast<const float4 *>(w0 + j * DP);
#pragma unroll
        for (int k = 0; k < DP / 4; ++k) {
            const float4 v = w[k];
#pragma unroll
            for (int n = 0; n < NE; ++n) {
                acc[n] = __ffma2_rn(make_float2(v.x, v.y), x[n][2 * k], acc[n]);
                acc[n] = __ffma2_rn(make_float2(v.z, v.w), x[n][2 * k + 1], acc[n]);
            }
        }
#pragma unroll
        for (int n = 0; n < NE; ++n) {
            const float y = fast_tanh(acc[n].x + acc[n].y);
            if (j & 1) h[n][j / 2].y = y; else h[n][j / 2].x = y;
        }
    }
}

// out[a] = bh[a] + sum_j wh_t[j][a] * tanh(b1[j] + sum_k w1[j][k] h[k]); the second hidden layer is never stored.
// wh_t is the head weight transposed and padded to AP floats per row; out is held as pairs.
template <int AP, int NE>
__device__ __forceinline__ void layer1_and_head(const float *w1, const float *b1, const float *wh_t, const float *bh,
                                                const float2 (&h)[NE][H / 2], float2 (&out)[NE][AP / 2])
{
#pragma unroll
    for (int n = 0; n < NE; ++n)
#pragma unroll
        for (int a = 0; a < AP / 2; ++a) out[n][a] = make_float2(bh[2 * a], bh[2 * a + 1]);
#pragma unroll 2
    for (int j = 0; j < H; ++j) {
        float2 acc0[NE], acc1[NE];                // two pair chains per env: four independent FMA chains
#pragma unroll
        for (int n = 0; n < NE; ++n) { acc0[n] = make_float2(b1[j], 0.f); acc1[n] = make_float2(0.f, 0.f); }
        const float4 *w = reinterpret_cast<const float4 *>(w1 + j * H);
#pragma unroll
        for (int k = 0; k < H / 4; ++k) {
            const float4 v = w[k];
#pragma unroll
            for (int n = 0; n < NE; ++n) {
                acc0[n] = __ffma2_rn(make_float2(v.x, v.y), h[n][2 * k], acc0[n]);
                acc1[n] = __ffma2_rn(make_float2(v.z, v.w), h[n][2 * k + 1], acc1[n]);
            }
        }
        float2 g[NE];
#pragma unroll
        for (int n = 0; n < NE; ++n) {
            const float t = fast_tanh((acc0[n].x + acc0[n].y) + (acc1[n].x + acc1[n].y));
            g[n] = make_float2(t, t);
        }
        const float4 *t4 = reinterpret_cast<const float4 *>(wh_t + j * AP);
#pragma unroll
        for (int a = 0; a < AP / 4; ++a) {
            const float4 v = t4[a];
#pragma unroll
            for (int n = 0; n < NE; ++n) {
                out[n][2 * a] = __ffma2_rn(make_float2(v.x, v.y), g[n], out[n][2 * a]);
                out[n][2 * a + 1] = __ffma2_rn(make_float2(v.z, v.w), g[n], out[n][2 * a + 1]);
            }
        }
    }
}

// DP: obs_dim padded to a multiple of 4 (<= 32); AP: act_dim padded to a multiple of 4 (<= 16).
template <int DP, int AP, int NE>
__global__ void __launch_bounds__(kThreads) policy_forward_kernel(const sng_mlp m, const float *__restrict__ obs,
                                                                 const float *__restrict__ noise,
                                                                 const float *__restrict__ low,
                                                                 const float *__restrict__ high, float *raw_actions,
                                                                 float *actions, float *values, float *log_probs,
                                                                 long long n_envs)
{
    extern __shared__ __align__(16) float sm[];
    const int D = m.obs_dim, A = m.act_dim;
    // ---- weights into shared memory, rows padded: [H][DP] | [H] | [H][H] | [H] | head^T [H][AP or 4] | head bias ----
    float *p_w0 = sm, *p_b0 = p_w0 + H * DP, *p_w1 = p_b0 + H, *p_b1 = p_w1 + H * H, *p_wh = p_b1 + H, *p_bh = p_wh + H * AP;
    float *v_w0 = p_bh + AP, *v_b0 = v_w0 + H * DP, *v_w1 = v_b0 + H, *v_b1 = v_w1 + H * H, *v_wh = v_b1 + H, *v_bh = v_wh + H * 4;
    const int tid = threadIdx.x;
    for (int i = tid; i < H * DP; i += kThreads) {
        const int j = i / DP, k = i - j * DP;
        p_w0[i] = k < D ? m.w_pi0[j * D + k] : 0.f;
        v_w0[i] = k < D ? m.w_vf0[j * D + k] : 0.f;
    }
    for (int i = tid; i < H * H; i += kThreads) { p_w1[i] = m.w_pi1[i]; v_w1[i] = m.w_vf1[i]; }
    for (int i = tid; i < H; i += kThreads) { p_b0[i] = m.b_pi0[i]; p_b1[i] = m.b_pi1[i]; v_b0[i] = m.b_vf0[i]; v_b1[i] = m.b_vf1[i]; }
    for (int i = tid; i < H * AP; i += kThreads) {
        const int j = i / AP, a = i - j * AP;
        p_wh[i] = a < A ? m.w_act[a * H + j] : 0.f;          // transposed: row j holds the A head weights of hidden unit j
    }
    for (int i = tid; i < H * 4; i += kThreads) v_wh[i] = (i & 3) == 0 ? m.w_val[i >> 2] : 0.f;
    if (tid < AP) p_bh[tid] = tid < A ? m.b_act[tid] : 0.f;
    if (tid < 4) v_bh[tid] = tid == 0 ? m.b_val[0] : 0.f;
    __syncthreads();

    // NE envs per thread (env e + n * stride).  Observation rows are read straight from global memory: 29 strided
    // 4-byte loads per thread hit each 128-byte line four times through L1 -- cheap next to the ~17k arithmetic
    // instructions per env, and it keeps shared memory to the weights alone.
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long e = (long long)blockIdx.x * kThreads + tid; e < n_envs; e += NE * stride) {
        float2 x[NE][DP / 2];
#pragma unroll
        for (int n = 0; n < NE; ++n) {
            const long long en = e + n * stride < n_envs ? e + n * stride : e;    // a missing env repeats env e (not stored)
#pragma unroll
            for (int k = 0; k < DP / 2; ++k)
                x[n][k] = make_float2(2 * k < D ? __ldg(obs + en * D + 2 * k) : 0.f, 2 * k + 1 < D ? __ldg(obs + en * D + 2 * k + 1) : 0.f);
        }
        float2 h[NE][H / 2];
        // ---- critic ----
        float2 val[NE][2];
        layer0<DP, NE>(v_w0, v_b0, x, h);
        layer1_and_head<4, NE>(v_w1, v_b1, v_wh, v_bh, h, val);
#pragma unroll
        for (int n = 0; n < NE; ++n)
            if (e + n * stride < n_envs) values[e + n * stride] = val[n][0].x;
        if (actions == nullptr) continue;        // value-only call (bootstrap value of the last observation)
        // ---- actor ----
        float2 mean[NE][AP / 2];
        layer0<DP, NE>(p_w0, p_b0, x, h);
        layer1_and_head<AP, NE>(p_w1, p_b1, p_wh, p_bh, h, mean);
#pragma unroll
        for (int n = 0; n < NE; ++n) {
            const long long en = e + n * stride;
            if (en >= n_envs) continue;
            float lp = 0.f;
#pragma unroll
            for (int a = 0; a < AP; ++a) {
                if (a < A) {
                    const float ls = __ldg(m.log_std + a);
                    const float z = noise ? noise[en * A + a] : 0.f;
                    const float mu = (a & 1) ? mean[n][a / 2].y : mean[n][a / 2].x;
                    const float act = fmaf(z, expf(ls), mu);               // DiagGaussian sample
                    raw_actions[en * A + a] = act;
                    actions[en * A + a] = fminf(fmaxf(act, __ldg(low + a)), __ldg(high + a));   // SB3 clips Box actions before env.step
                    lp += -0.5f * z * z - ls - 0.91893853320467274f;     // log N(act; mean, std), 0.5 * log(2 pi)
                }
            }
            log_probs[en] = lp;
        }
    }
}

// The dynamic shared-memory opt-in is a per-function, per-DEVICE attribute: raise it once per (device, kernel).
int ensure_smem(int device, const void *kern, size_t smem)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> limit;
    std::lock_guard<std::mutex> lock(mu);
    size_t &cur = limit[std::make_pair(device, kern)];
    if (smem <= cur) return SNG_OK;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return SNG_ERR_CUDA;
    cur = smem;
    return SNG_OK;
}

// Makes the device that owns `ptr` current for the lifetime of the guard (a process may drive several GPUs).
struct PointerDeviceGuard {
    int prev = -1, dev = -1;
    bool switched = false;
    explicit PointerDeviceGuard(const void *ptr)
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type == cudaMemoryTypeDevice) dev = at.device;
        if (dev >= 0 && cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~PointerDeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
};

template <int DP, int AP>
int launch(const sng_mlp &m, const float *obs, const float *noise, const float *low, const float *high, float *raw,
           float *act, float *val, float *lp, long long n, cudaStream_t st)
{
    constexpr int NE = 2;      // envs per thread
    auto kern = policy_forward_kernel<DP, AP, NE>;
    const size_t smem = sizeof(float) * ((size_t)2 * (H * DP + H + H * H + H) + H * AP + AP + H * 4 + 4);
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);                     // the caller (sng_policy_forward) made the buffers' device current
    if (ensure_smem(dev, (const void *)kern, smem) != SNG_OK) return SNG_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
    long long grid = (long long)sms * (per_sm > 0 ? per_sm : 1);
    const long long need = (n + (long long)kThreads * NE - 1) / ((long long)kThreads * NE);
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, kThreads, smem, st>>>(m, obs, noise, low, high, raw, act, val, lp, n);
    return cudaGetLastError() == cudaSuccess ? SNG_OK : SNG_ERR_CUDA;
}

}  // namespace

extern "C" int sng_policy_forward(const sng_mlp *mlp, const float *obs, const float *noise, const float *low,
                                  const float *high, float *raw_actions, float *actions, float *values,
                                  float *log_probs, int64_t n_envs, void *stream)
{
    if (!mlp || mlp->struct_size != sizeof(sng_mlp) || !obs || !values || n_envs < 1) return SNG_ERR_ARG;
    if (actions && (!raw_actions || !log_probs || !low || !high)) return SNG_ERR_ARG;
    if (mlp->hidden != H || mlp->obs_dim < 1 || mlp->obs_dim > 32 || mlp->act_dim < 1 || mlp->act_dim > 16) return SNG_ERR_UNSUPPORTED;
    const int dp = (mlp->obs_dim + 3) / 4 * 4, ap = (mlp->act_dim + 3) / 4 * 4;
    cudaStream_t st = (cudaStream_t)stream;
    PointerDeviceGuard guard(obs);
#define SNG_POLICY_CASE(DP, AP) \
    if (dp == DP && ap == AP) return launch<DP, AP>(*mlp, obs, noise, low, high, raw_actions, actions, values, log_probs, n_envs, st);
    SNG_POLICY_CASE(20, 8)    // N = 4:  D = 17, A = 5
    SNG_POLICY_CASE(28, 12)   // N = 8:  D = 25, A = 9
    SNG_POLICY_CASE(32, 12)   // N = 10: D = 29, A = 11
#undef SNG_POLICY_CASE
    return SNG_ERR_UNSUPPORTED;
}
