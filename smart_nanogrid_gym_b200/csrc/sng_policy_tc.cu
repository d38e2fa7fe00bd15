// sng_policy_tc.cu -- the actor-critic forward pass of rollout collection on the 5th-generation
// tensor cores: tcgen05.mma with the activations in tensor memory (TMEM) and the weights in shared memory.
//
// Caller in the reference: PPO("MlpPolicy", env).learn -> collect_rollouts (solvers/RL/ppo_train.py:89-102): two tanh
// 64-64 networks (actor, critic), a linear action head with a state-independent log-std, a linear value head.
//
// Mapping.  A tile is 128 envs = the M of one tcgen05.mma = the 128 lanes of TMEM: thread t of a 128-thread group
// owns env t of the tile and lane t of TMEM.  A CTA (one per SM, persistent) runs two such groups on alternating
// tiles, so one group's tensor-core phase overlaps the other's tanh epilogue; both share ONE copy of the weights in
// shared memory (117 KB, loaded once per CTA with cp.async.bulk).  Per tile and network:
//     layer 0   D[128x64] = X[128x32]  * W0^T        A operand = X in TMEM, B = W0 in shared memory (K-major)
//     epilogue  tanh(D) -> R (TMEM, written back with tcgen05.st: the accumulator layout IS the A-operand layout)
//     layer 1   D[128x64] = R[128x64]  * W1^T + b1
//     epilogue  tanh(D) -> R
//     head      D[128x16] = R[128x64]  * Wh^T + bh   (actor: 11 means; critic: the value in column 0)
// Numerics: FP32 in, FP32 out.  kind::tf32 reads 10 mantissa bits, so every operand is split x = hi + lo
// (hi = tf32(x), lo = tf32(x - hi)) and each product is three MMAs, hi*hi + lo*hi + hi*lo, accumulated in FP32
// in TMEM: ~2^-21 relative per product, the trainer keeps FP32-level values / log-probabilities
// (tests/test_gpu_rollout.py holds the same tolerances against the FP32 torch modules as for the CUDA-core kernel).
// Biases ride in the MMAs: X carries two constant-one columns (k = 30, 31); W0 has b0 in row k = 30, and layer 1 /
// the head add one extra K = 8 step  X[:, 24:32] * [0 .. 0, hi(b), lo(b)]^T.
// The weights are split and laid out in the canonical no-swizzle K-major core-matrix order once per weight update by
// sng_policy_pack (a [K/4][N][4] float image: core matrix = 8 rows x 16 bytes, SBO = 128 B, LBO = N * 16 B).
#include <cstdint>
#include <map>
#include <mutex>
#include <utility>

#include <cuda_runtime.h>

#include "../../include/sng.h"
#include "sng_tma.cuh"

namespace {
using namespace sng;

constexpr int H = 64;        // hidden width of SB3's default MlpPolicy
constexpr int KP = 32;       // observation width padded to the K of layer 0 (obs_dim <= 30; columns 30, 31 = 1.0)
constexpr int NH = 16;       // head width padded to the smallest N of an M = 128 MMA
constexpr int TILE = 128;    // envs per tile = MMA M = TMEM lanes
constexpr int GROUPS = 2;    // tiles in flight per CTA
constexpr int THREADS = GROUPS * TILE;

// ---- packed weight image (floats), per network ----
constexpr int OFF_W0HI = 0;
constexpr int OFF_W0LO = OFF_W0HI + KP * H;
constexpr int OFF_W1HI = OFF_W0LO + KP * H;
constexpr int OFF_W1LO = OFF_W1HI + H * H;
constexpr int OFF_B1 = OFF_W1LO + H * H;       // one K = 8 step: rows 6, 7 = hi(b1), lo(b1)
constexpr int OFF_WHHI = OFF_B1 + 8 * H;
constexpr int OFF_WHLO = OFF_WHHI + H * NH;
constexpr int OFF_BH = OFF_WHLO + H * NH;
constexpr int NET_FLOATS = OFF_BH + 8 * NH;    // 14,976 floats
constexpr int OFF_STD = 2 * NET_FLOATS;        // exp(log_std) [16] | log_std [16]   (read from global memory)
constexpr int IMG_FLOATS = OFF_STD + 32;
constexpr uint32_t IMG_SMEM_BYTES = 2u * NET_FLOATS * sizeof(float);   // the part staged in shared memory
static_assert(IMG_SMEM_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");

// TMEM columns of one group (256 of the SM's 512)
constexpr uint32_t C_XHI = 0, C_XLO = 32, C_P = 64, C_Q = 128, C_S = 192, GROUP_COLS = 256;

__device__ __forceinline__ uint32_t tf32_rna(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// In-kernel split of an activation: hi = x rounded to tf32 (integer add of half an ulp + mask: two instructions, where
// cvt.rna.tf32 expands to five), lo = x - hi exactly; the tensor core ignores lo's low 13 mantissa bits.
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo)
{
    hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// ------------------------------------------------------------------------------------------
// sng_policy_pack: FP32 nn.Linear weights -> the tf32 hi / lo image the forward kernel bulk-copies
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) policy_pack_kernel(const sng_mlp m, float *img)
{
    const int D = m.obs_dim, A = m.act_dim;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < IMG_FLOATS; idx += gridDim.x * blockDim.x) {
        float v = 0.f;
        bool lo = false;
        if (idx >= OFF_STD) {
            const int a = (idx - OFF_STD) & 15;
            const float ls = a < A ? m.log_std[a] : 0.f;
            img[idx] = (idx - OFF_STD) < 16 ? expf(ls) : ls;
            continue;
        }
        const int net = idx / NET_FLOATS, r = idx - net * NET_FLOATS;     // net 0 = critic, 1 = actor
        const float *w0 = net ? m.w_pi0 : m.w_vf0, *b0 = net ? m.b_pi0 : m.b_vf0;
        const float *w1 = net ? m.w_pi1 : m.w_vf1, *b1 = net ? m.b_pi1 : m.b_vf1;
        const float *wh = net ? m.w_act : m.w_val, *bh = net ? m.b_act : m.b_val;
        const int nh = net ? A : 1;
        if (r < OFF_W1HI) {                          // layer 0: [64][KP]; k = 30 carries the bias
            const int q = r < OFF_W0LO ? r : r - OFF_W0LO;
            lo = r >= OFF_W0LO;
            const int kc = q / (H * 4), n = (q / 4) % H, k = kc * 4 + (q & 3);
            v = k < D ? w0[n * D + k] : (k == KP - 2 ? b0[n] : 0.f);
        } else if (r < OFF_B1) {                     // layer 1: [64][64]
            const int q = r < OFF_W1LO ? r - OFF_W1HI : r - OFF_W1LO;
            lo = r >= OFF_W1LO;
            const int kc = q / (H * 4), n = (q / 4) % H, k = kc * 4 + (q & 3);
            v = w1[n * H + k];
        } else if (r < OFF_WHHI) {                   // bias step of layer 1: k = 6 -> hi(b1), k = 7 -> lo(b1)
            const int q = r - OFF_B1;
            const int kc = q / (H * 4), n = (q / 4) % H, k = kc * 4 + (q & 3);
            v = (k >= 6) ? b1[n] : 0.f;
            lo = k == 7;
            if (k < 6) { img[idx] = 0.f; continue; }
        } else if (r < OFF_BH) {                     // head: [16][64]
            const int q = r < OFF_WHLO ? r - OFF_WHHI : r - OFF_WHLO;
            lo = r >= OFF_WHLO;
            const int kc = q / (NH * 4), n = (q / 4) % NH, k = kc * 4 + (q & 3);
            v = n < nh ? wh[n * H + k] : 0.f;
        } else {                                     // bias step of the head
            const int q = r - OFF_BH;
            const int kc = q / (NH * 4), n = (q / 4) % NH, k = kc * 4 + (q & 3);
            v = (k >= 6 && n < nh) ? bh[n] : 0.f;
            lo = k == 7;
        }
        const float hi = __uint_as_float(tf32_rna(v));
        img[idx] = lo ? __uint_as_float(tf32_rna(v - hi)) : hi;
    }
}

// ------------------------------------------------------------------------------------------
// tcgen05 / TMEM primitives (inline PTX; SASS: UTCHMMA / LDTM / STTM / UTCBAR)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#define SNG_R8(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define SNG_W8(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])
// 32 lanes x 32 columns: thread l of the warp gets columns [col, col + 32) of TMEM lane (lane base + l)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : SNG_R8(v, 0), SNG_R8(v, 8), SNG_R8(v, 16), SNG_R8(v, 24)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : SNG_R8(v, 0), SNG_R8(v, 8)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        SNG_W8(v, 0), SNG_W8(v, 8), SNG_W8(v, 16), SNG_W8(v, 24)
        : "memory");
}

// Shared-memory matrix descriptor of a K-major [N][8] tf32 slice in the canonical no-swizzle layout: start address,
// leading-dimension byte offset (between the two 16-byte K chunks) = N * 16, stride byte offset (between 8-row
// groups) = 128, descriptor version 1 (sm_100), no swizzle.  All fields in 16-byte units.
__device__ __forceinline__ uint64_t b_desc(uint32_t smem_addr, uint32_t n_rows)
{
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (n_rows << 16);
    const uint32_t hi = 8u | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}
// Instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n.
__host__ __device__ constexpr uint32_t i_desc(uint32_t n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24); }

// D[tmem] (+)= A[tmem] * B[smem]^T, one K = 8 step; issued by ONE thread for the whole 128-row tile
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One layer: D = (A_hi + A_lo) * (W_hi + W_lo)^T without the lo * lo term [+ the bias step], K = 8 * ksteps.
// w_hi / w_lo / bias: shared-memory byte addresses of the canonical [K/4][N][4] images (bias = 0: none).
__device__ __forceinline__ void issue_layer(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, uint32_t w_lo, int ksteps,
                                            uint32_t n, uint32_t x_ones, uint32_t bias)
{
    const uint32_t idesc = i_desc(n);
    const uint32_t kbytes = 2u * n * 16u;            // one K = 8 step = two 16-byte chunk planes
#pragma unroll 1
    for (int s = 0; s < ksteps; ++s) {
        const uint64_t dh = b_desc(w_hi + s * kbytes, n), dl = b_desc(w_lo + s * kbytes, n);
        mma_tf32_ts(d, a_hi + 8u * s, dh, idesc, s > 0 ? 1u : 0u);
        mma_tf32_ts(d, a_lo + 8u * s, dh, idesc, 1u);
        mma_tf32_ts(d, a_hi + 8u * s, dl, idesc, 1u);
    }
    if (bias) mma_tf32_ts(d, x_ones, b_desc(bias, n), idesc, 1u);
}

// tanh(x) = 1 - 2 / (exp(2x) + 1) on the special-function unit (EX2 + RCP): absolute error ~1e-7, saturates correctly
__device__ __forceinline__ float fast_tanh(float x)
{
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}

__device__ __forceinline__ void group_barrier(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(TILE) : "memory"); }

// hidden-layer epilogue: accumulator columns [src, src + 64) -> tanh -> hi part written back in place (it becomes
// the A operand of the next layer), lo part to [dst_lo, dst_lo + 64)
__device__ __forceinline__ void tanh_epilogue(uint32_t src, uint32_t dst_lo)
{
    __syncwarp();                      // the .sync.aligned TMEM accesses below need the whole warp (lane 0 issued the MMAs)
#pragma unroll
    for (int c = 0; c < H; c += 32) {
        uint32_t v[32], w[32];
        tmem_ld32(src + c, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            split_tf32(fast_tanh(__uint_as_float(v[j])), v[j], w[j]);
        }
        tmem_st32(src + c, v);
        tmem_st32(dst_lo + c, w);
    }
    tmem_wait_st();
}

struct Smem {
    // byte offsets into dynamic shared memory
    uint32_t obs_stage, row_stage, group_bytes, bars;
};
__host__ __device__ inline Smem smem_plan(int D, int A)
{
    Smem s;
    s.obs_stage = align128((uint32_t)(TILE * D * sizeof(float)));
    s.row_stage = align128((uint32_t)(TILE * A * sizeof(float)));
    s.group_bytes = s.obs_stage + 3 * s.row_stage;                      // obs | noise | raw actions | clipped actions
    s.bars = align128(IMG_SMEM_BYTES) + GROUPS * s.group_bytes;
    return s;
}

__global__ void __launch_bounds__(THREADS, 1)
    policy_tc_kernel(const float *__restrict__ img, const float *__restrict__ obs, const float *__restrict__ noise,
                     const float *__restrict__ low, const float *__restrict__ high, float *raw_actions, float *actions,
                     float *values, float *log_probs, long long n_envs, int D, int A, int aligned)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const Smem sp = smem_plan(D, A);
    const int g = threadIdx.x / TILE, t = threadIdx.x % TILE, warp = threadIdx.x >> 5;
    unsigned char *gbase = smem + align128(IMG_SMEM_BYTES) + (size_t)g * sp.group_bytes;
    float *obs_s = reinterpret_cast<float *>(gbase);
    float *noise_s = reinterpret_cast<float *>(gbase + sp.obs_stage);
    float *raw_s = reinterpret_cast<float *>(gbase + sp.obs_stage + sp.row_stage);
    float *act_s = reinterpret_cast<float *>(gbase + sp.obs_stage + 2 * sp.row_stage);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + sp.bars);
    uint64_t *wbar = bars;                                  // the weight image has landed
    uint64_t *obar = bars + 1 + 3 * g, *nbar = obar + 1, *mbar = obar + 2;   // obs rows | noise rows | MMAs of this group
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 1 + 3 * GROUPS);
    const bool value_only = actions == nullptr;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 1 + 3 * GROUPS; ++i) mbar_init(bars + i, 1);
        fence_mbar_init();
        // the weight image: one copy per CTA, shared by both groups
        mbar_expect_tx(wbar, IMG_SMEM_BYTES);
        for (uint32_t off = 0; off < IMG_SMEM_BYTES; off += 16384u) {
            const uint32_t n = IMG_SMEM_BYTES - off < 16384u ? IMG_SMEM_BYTES - off : 16384u;
            bulk_g2s(smem + off, reinterpret_cast<const unsigned char *>(img) + off, n, wbar);
        }
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot + g * GROUP_COLS;                  // this group's columns, lane 0
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // ... as seen by this warp's 32 lanes
    const uint32_t w_base = smem_u32(smem);

    const long long n_tiles = (n_envs + TILE - 1) / TILE;
    const uint32_t obs_bytes = (uint32_t)(TILE * D * sizeof(float)), row_bytes = (uint32_t)(TILE * A * sizeof(float));
    uint32_t o_phase = 0, n_phase = 0, m_phase = 0;
    bool first = true, prefetched = false, stores_pending = false;

#pragma unroll 1
    for (long long tile = (long long)blockIdx.x * GROUPS + g; tile < n_tiles; tile += (long long)gridDim.x * GROUPS) {
        const long long e0 = tile * TILE;
        const int nv = (int)((n_envs - e0) < TILE ? (n_envs - e0) : TILE);
        const bool full = aligned && nv == TILE;             // whole tile, 16-byte aligned rows: copy engine
        const bool sample = !value_only && noise != nullptr;
        // ---- stage this tile's observation rows (and noise rows) in shared memory ----
        if (t == 0 && stores_pending) bulk_wait_read<0>();   // the previous tile's action rows have left their stages
        if (full) {
            if (t == 0) {
                if (!prefetched) {
                    mbar_expect_tx(obar, obs_bytes);
                    bulk_g2s(obs_s, obs + e0 * D, obs_bytes, obar);
                }
                if (sample) {
                    mbar_expect_tx(nbar, row_bytes);
                    bulk_g2s(noise_s, noise + e0 * A, row_bytes, nbar);
                }
            }
            mbar_wait(obar, o_phase);
            o_phase ^= 1u;
        } else {
            for (int k = t; k < nv * D; k += TILE) obs_s[k] = obs[e0 * D + k];
            if (sample)
                for (int k = t; k < nv * A; k += TILE) noise_s[k] = noise[e0 * A + k];
            group_barrier(g);
        }
        // ---- X = [obs | 0 | 1 1] -> tf32 hi / lo -> TMEM (the A operand of layer 0) ----
        {
            uint32_t xh[32], xl[32];
            const float *row = obs_s + t * D;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const float x = k < D ? (t < nv ? row[k] : 0.f) : (k >= KP - 2 ? 1.0f : 0.f);
                split_tf32(x, xh[k], xl[k]);
            }
            __syncwarp();
            tmem_st32(tlane + C_XHI, xh);
            tmem_st32(tlane + C_XLO, xl);
            tmem_wait_st();
        }
        tc_fence_before();
        group_barrier(g);
        // the observation stage is free again: fetch the next tile's rows while this one is computed
        const long long next = tile + (long long)gridDim.x * GROUPS;
        prefetched = aligned && next < n_tiles && (n_envs - next * TILE) >= TILE;
        if (t == 0 && prefetched) {
            mbar_expect_tx(obar, obs_bytes);
            bulk_g2s(obs_s, obs + next * TILE * D, obs_bytes, obar);
        }

#pragma unroll 1
        for (int net = 0; net < (value_only ? 1 : 2); ++net) {           // 0 = critic, 1 = actor
            const uint32_t wn = w_base + (uint32_t)(net * NET_FLOATS * sizeof(float));
            // ---- layer 0 -> P ----
            if (t == 0) {
                if (first) { mbar_wait(wbar, 0); first = false; }
                tc_fence_after();
                issue_layer(tmem + C_P, tmem + C_XHI, tmem + C_XLO, wn + OFF_W0HI * 4, wn + OFF_W0LO * 4, KP / 8, H, 0, 0);
                mma_commit(mbar);
            }
            mbar_wait(mbar, m_phase);
            m_phase ^= 1u;
            tc_fence_after();
            tanh_epilogue(tlane + C_P, tlane + C_Q);                      // R = (P, Q)
            tc_fence_before();
            group_barrier(g);
            // ---- layer 1 -> S ----
            if (t == 0) {
                tc_fence_after();
                issue_layer(tmem + C_S, tmem + C_P, tmem + C_Q, wn + OFF_W1HI * 4, wn + OFF_W1LO * 4, H / 8, H, tmem + C_XHI + (KP - 8),
                            wn + OFF_B1 * 4);
                mma_commit(mbar);
            }
            mbar_wait(mbar, m_phase);
            m_phase ^= 1u;
            tc_fence_after();
            tanh_epilogue(tlane + C_S, tlane + C_Q);                      // R = (S, Q)
            tc_fence_before();
            group_barrier(g);
            // ---- head -> P[0:16] ----
            if (t == 0) {
                tc_fence_after();
                issue_layer(tmem + C_P, tmem + C_S, tmem + C_Q, wn + OFF_WHHI * 4, wn + OFF_WHLO * 4, H / 8, NH, tmem + C_XHI + (KP - 8),
                            wn + OFF_BH * 4);
                mma_commit(mbar);
            }
            mbar_wait(mbar, m_phase);
            m_phase ^= 1u;
            tc_fence_after();
            uint32_t out[16];
            __syncwarp();
            tmem_ld16(tlane + C_P, out);
            tmem_wait_ld();
            if (net == 0) {
                if (t < nv) values[e0 + t] = __uint_as_float(out[0]);
                tc_fence_before();
                group_barrier(g);          // every lane has read the critic's head before the actor's layer 0 overwrites P
                continue;
            }
            // ---- DiagGaussian sample, clip to the Box, log-probability ----
            if (sample && full) {
                mbar_wait(nbar, n_phase);
                n_phase ^= 1u;
            }
            float lp = 0.f;
#pragma unroll
            for (int a = 0; a < NH; ++a) {
                if (a < A) {
                    const float sd = __ldg(img + OFF_STD + a), ls = __ldg(img + OFF_STD + 16 + a);
                    const float z = sample ? noise_s[t * A + a] : 0.f;
                    const float x = fmaf(z, sd, __uint_as_float(out[a]));
                    raw_s[t * A + a] = x;
                    act_s[t * A + a] = fminf(fmaxf(x, __ldg(low + a)), __ldg(high + a));   // SB3 clips Box actions before env.step
                    lp += -0.5f * z * z - ls - 0.91893853320467274f;                        // log N(x; mean, std)
                }
            }
            if (t < nv) log_probs[e0 + t] = lp;
            if (full) {
                fence_proxy_async();
                group_barrier(g);
                if (t == 0) {
                    bulk_s2g(raw_actions + e0 * A, raw_s, row_bytes);
                    bulk_s2g(actions + e0 * A, act_s, row_bytes);
                    bulk_commit();
                }
                stores_pending = true;
            } else {
                group_barrier(g);
                for (int k = t; k < nv * A; k += TILE) {
                    raw_actions[e0 * A + k] = raw_s[k];
                    actions[e0 * A + k] = act_s[k];
                }
            }
        }
        // every lane's tcgen05.ld of this tile has completed (wait::ld above) before the next tile's X store and MMAs:
        // ordered by the fence + group barrier that follows the X store
        tc_fence_before();
    }
    if (t == 0 && stores_pending) bulk_wait_read<0>();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

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

bool aligned16(const void *p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" size_t sng_policy_packed_bytes(void) { return (size_t)IMG_FLOATS * sizeof(float); }

extern "C" int sng_policy_pack(const sng_mlp *mlp, void *packed, void *stream)
{
    if (!mlp || mlp->struct_size != sizeof(sng_mlp) || !packed) return SNG_ERR_ARG;
    if (mlp->hidden != H || mlp->obs_dim < 1 || mlp->obs_dim > KP - 2 || mlp->act_dim < 1 || mlp->act_dim > NH) return SNG_ERR_UNSUPPORTED;
    if (!aligned16(packed)) return SNG_ERR_ARG;
    PointerDeviceGuard guard(packed);
    policy_pack_kernel<<<(IMG_FLOATS + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*mlp, reinterpret_cast<float *>(packed));
    return cudaGetLastError() == cudaSuccess ? SNG_OK : SNG_ERR_CUDA;
}

extern "C" int sng_policy_forward_packed(const void *packed, int obs_dim, int act_dim, const float *obs, const float *noise,
                                         const float *low, const float *high, float *raw_actions, float *actions,
                                         float *values, float *log_probs, int64_t n_envs, void *stream)
{
    if (!packed || !obs || !values || n_envs < 1) return SNG_ERR_ARG;
    if (actions && (!raw_actions || !log_probs || !low || !high)) return SNG_ERR_ARG;
    if (obs_dim < 1 || obs_dim > KP - 2 || act_dim < 1 || act_dim > NH) return SNG_ERR_UNSUPPORTED;
    PointerDeviceGuard guard(obs);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    const Smem sp = smem_plan(obs_dim, act_dim);
    const size_t smem = sp.bars + 128;
    if (ensure_smem(dev, (const void *)policy_tc_kernel, smem) != SNG_OK) return SNG_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long n_tiles = (n_envs + TILE - 1) / TILE;
    long long grid = (n_tiles + GROUPS - 1) / GROUPS;
    if (grid > sms) grid = sms;
    const int aligned = aligned16(obs) && aligned16(noise) && aligned16(raw_actions) && aligned16(actions) && aligned16(packed);
    if (!aligned16(packed)) return SNG_ERR_ARG;
    policy_tc_kernel<<<(unsigned)grid, THREADS, smem, (cudaStream_t)stream>>>(reinterpret_cast<const float *>(packed), obs, noise, low, high,
                                                                             raw_actions, actions, values, log_probs,
                                                                             (long long)n_envs, obs_dim, act_dim, aligned);
    return cudaGetLastError() == cudaSuccess ? SNG_OK : SNG_ERR_CUDA;
}
