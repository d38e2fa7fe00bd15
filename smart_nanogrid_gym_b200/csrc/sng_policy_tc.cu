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
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include <cuda_runtime.h>

#include "../../include/sng.h"
#include "sng_engine.cuh"
#include "sng_tma.cuh"

namespace {
using namespace sng;
long long *g_trace = nullptr;      // debugging aid: sng_policy_debug_trace
int g_policy_pdl = 0;              // sng_policy_set_launch_mode: bit 0 programmatic dependent launch, bit 1 the CTA takes its SM's whole shared memory

constexpr int H = 64;        // hidden width of SB3's default MlpPolicy
constexpr int KP = 32;       // observation width padded to the K of layer 0 (obs_dim <= 30; columns 30, 31 = 1.0)
constexpr int NH = 16;       // head width padded to the smallest N of an M = 128 MMA
constexpr int TILE = 128;    // envs per tile = MMA M = TMEM lanes
constexpr int CB = 4;        // warps per 32 TMEM lanes: each takes 64 / CB of a layer's columns in the epilogues
constexpr int THREADS = CB * TILE;   // compute threads (16 warps); one more warp issues the MMAs
constexpr float kTanhScale = 2.8853900817779268f;   // 2 log2(e), folded into the hidden layers' weights and biases

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

// TMEM columns (all 512 of the SM): two X buffers (tile i and tile i + 1), and per network P | Q | S
constexpr uint32_t C_X0 = 0, C_X1 = 448, C_NET0 = 64, C_NETSTRIDE = 192, C_P = 0, C_Q = 64, C_S = 128;
__device__ __forceinline__ uint32_t x_cols(int buf) { return buf ? C_X1 : C_X0; }      // hi at +0, lo at +32

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
            v = kTanhScale * (k < D ? w0[n * D + k] : (k == KP - 2 ? b0[n] : 0.f));
        } else if (r < OFF_B1) {                     // layer 1: [64][64]
            const int q = r < OFF_W1LO ? r - OFF_W1HI : r - OFF_W1LO;
            lo = r >= OFF_W1LO;
            const int kc = q / (H * 4), n = (q / 4) % H, k = kc * 4 + (q & 3);
            v = kTanhScale * w1[n * H + k];
        } else if (r < OFF_WHHI) {                   // bias step of layer 1: k = 6 -> hi(b1), k = 7 -> lo(b1)
            const int q = r - OFF_B1;
            const int kc = q / (H * 4), n = (q / 4) % H, k = kc * 4 + (q & 3);
            v = (k >= 6) ? kTanhScale * b1[n] : 0.f;
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : SNG_R8(v, 0), SNG_R8(v, 8)
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        SNG_W8(v, 0), SNG_W8(v, 8)
        : "memory");
}

__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr)
{
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return v;
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

// D[tmem] (+)= A[tmem] * B[smem]^T, one K = 8 step.  Executed by a WHOLE warp with warp-uniform operands; the elected
// lane (`elected` != 0 in exactly one lane) issues it for the 128-row tile.  (Issued from divergent code -- if (t == 0) --
// the compiler wraps every UTCHMMA in an election loop: ~45 cycles per MMA, more than an N = 64 MMA takes to execute.)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                            uint32_t elected)
{
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(elected)
        : "memory");
}
__device__ __forceinline__ uint32_t elect_one()
{
    uint32_t el;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(el));
    return el;
}
// arrives on the mbarrier when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint32_t bar_smem_addr, uint32_t elected)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar_smem_addr), "r"(elected) : "memory");
}

// One layer: D = (A_hi + A_lo) * (W_hi + W_lo)^T without the lo * lo term [+ the bias step], K = 8 * KSTEPS, N columns.
// w_hi / w_lo / bias: shared-memory byte addresses of the canonical [K/4][N][4] images (bias = 0: none).
template <int KSTEPS, int N>
__device__ __forceinline__ void issue_layer(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, uint32_t w_lo, uint32_t x_ones,
                                            uint32_t bias, uint32_t elected)
{
    constexpr uint32_t idesc = i_desc(N);
    constexpr uint32_t kstep16 = 2u * N;             // one K = 8 step = two 16-byte chunk planes of N rows, in 16-byte units
    const uint64_t dh0 = b_desc(w_hi, N), dl0 = b_desc(w_lo, N);
#pragma unroll
    for (int s = 0; s < KSTEPS; ++s) {
        const uint64_t dh = dh0 + (uint64_t)(s * kstep16), dl = dl0 + (uint64_t)(s * kstep16);
        mma_tf32_ts(d, a_hi + 8u * s, dh, idesc, s > 0 ? 1u : 0u, elected);
        mma_tf32_ts(d, a_lo + 8u * s, dh, idesc, 1u, elected);
        mma_tf32_ts(d, a_hi + 8u * s, dl, idesc, 1u, elected);
    }
    if (bias) mma_tf32_ts(d, x_ones, b_desc(bias, N), idesc, 1u, elected);
}

// tanh of a PAIR of pre-activations and the tf32 split of the results.  The pack kernel scales the hidden layers' weights and
// biases by 2 log2(e), so the accumulator holds u = 2 log2(e) x and tanh(x) = 1 - 2 / (2^u + 1):
//   one MUFU.EX2 per element; the reciprocal is computed on the FMA pipe (magic-constant seed, 5 % error, one cubic and one
//   quadratic Newton step: 8e-8) with Blackwell's packed FP32 pairs (FFMA2) -- with EX2 + RCP the epilogue is bound by the
//   16-lane special-function unit (2 x 8 cycles per warp and element).  |error| < 3e-7 absolute.
__device__ __forceinline__ void tanh2_split(uint32_t &a, uint32_t &b, uint32_t &lo_a, uint32_t &lo_b)
{
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(__uint_as_float(a), 26.0f)));   // 2^26 + 1: tanh = 1 - 3e-8
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(__uint_as_float(b), 26.0f)));
    const float2 one = make_float2(1.0f, 1.0f);
    const float2 nd = __ffma2_rn(make_float2(e0, e1), make_float2(-1.0f, -1.0f), make_float2(-1.0f, -1.0f));   // -(e + 1)
    // 1 / d seed: bits(r0) = 0x7EF311C7 - bits(d); with the sign bit of -d folded into the constant
    float2 r = make_float2(__uint_as_float(0xFEF311C7u - __float_as_uint(nd.x)), __uint_as_float(0xFEF311C7u - __float_as_uint(nd.y)));
    float2 err = __ffma2_rn(nd, r, one);                 // 1 - d r
    r = __ffma2_rn(r, __ffma2_rn(err, err, err), r);     // cubic step: r (1 + err + err^2)
    err = __ffma2_rn(nd, r, one);
    r = __ffma2_rn(r, err, r);                           // quadratic step
    const float2 y = __ffma2_rn(r, make_float2(-2.0f, -2.0f), one);
    const uint32_t ha = __float_as_uint(y.x) & 0xFFFFE000u, hb = __float_as_uint(y.y) & 0xFFFFE000u;   // tf32 by truncation: lo < 2^-10 |y|
    const float2 l = __fadd2_rn(y, make_float2(-__uint_as_float(ha), -__uint_as_float(hb)));
    a = ha; b = hb;
    lo_a = __float_as_uint(l.x); lo_b = __float_as_uint(l.y);
}

// ---- in-kernel exploration noise: Philox4x32-10 (Salmon et al., SC'11) keyed by (seed, step), counter = (global env, column
//      block); one block yields the four standard normals of a thread's four action columns (Box-Muller) ----
// u in (0, 1): 24 random bits, centred; z0 = sqrt(-2 ln u1) cos(2 pi u2), z1 = ... sin(2 pi u2)
__device__ __forceinline__ void box_muller(uint32_t x1, uint32_t x2, float &z0, float &z1)
{
    const float u1 = ((float)(x1 >> 8) + 0.5f) * 5.9604644775390625e-08f;
    const float u2 = ((float)(x2 >> 8) + 0.5f) * 5.9604644775390625e-08f;
    const float r = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(6.283185307179586f * u2, &sn, &cs);
    z0 = r * cs;
    z1 = r * sn;
}


// hidden-layer epilogue of one warp: 16 accumulator columns [src, src + 16) of its 32 lanes -> tanh -> hi part written back
// in place (it becomes the A operand of the next layer), lo part to [dst_lo, dst_lo + 16)
__device__ __forceinline__ void tanh_epilogue(uint32_t src, uint32_t dst_lo)
{
    __syncwarp();                      // the .sync.aligned TMEM accesses below need the whole warp
    uint32_t v[16], w[16];
    tmem_ld16(src, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; j += 2) tanh2_split(v[j], v[j + 1], w[j], w[j + 1]);
    tmem_st16(src, v);
    tmem_st16(dst_lo, w);
    tmem_wait_st();
}

struct Smem {
    // byte offsets into dynamic shared memory
    uint32_t obs_stage, row_stage, stages, rows, out_rows, consts, bars;
};
// `sets` io-warp sets (1, or 2 with the fused env step); each set owns one observation stage's parity, three row stages
// (noise | raw actions | clipped actions) and -- fused env step -- one stage of next-observation rows
__host__ __device__ inline Smem smem_plan(int D, int A, int sets, bool env_step = true)
{
    Smem s;
    s.obs_stage = align128((uint32_t)(TILE * D * sizeof(float)));
    s.row_stage = align128((uint32_t)(TILE * A * sizeof(float)));
    s.stages = align128(IMG_SMEM_BYTES);                                // obs x 2 | per set: noise | raw actions | clipped actions
    s.rows = s.stages + 2 * s.obs_stage;
    s.out_rows = s.rows + (uint32_t)sets * 3 * s.row_stage;            // per set (fused env step only): next-observation rows
    s.consts = s.out_rows + ((sets > 1 && env_step) ? (uint32_t)sets * s.obs_stage : 0u);   // std | log_std | low | high, 16 floats each
    s.bars = s.consts + 4 * NH * (uint32_t)sizeof(float);
    return s;
}

// One arrival per WARP (512 per-thread arrivals on one shared-memory word serialise: ~0.25 us per barrier round).  Every
// lane has completed its TMEM stores (tcgen05.wait::st) and fenced them; __syncwarp orders the lanes before lane 0 arrives.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t *bar)
{
    __syncwarp();
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Layer `phase` (0: layer 0, X -> P;  1: layer 1, (P, Q) -> S;  2: head, (S, Q) -> `head_out`) of the network whose TMEM
// columns start at `tn` and whose weight image starts at shared-memory address `wn`; executed by the whole issuer warp
// with warp-uniform arguments.  `x` = columns of the tile's X buffer.
__device__ __forceinline__ void issue_phase(int phase, uint32_t tn, uint32_t x, uint32_t head_out, uint32_t wn, uint32_t done_bar)
{
    tc_fence_after();
    const uint32_t elected = elect_one();     // taken right here: ptxas then knows the predicate holds in ONE lane
    if (phase == 0)
        issue_layer<KP / 8, H>(tn + C_P, x, x + 32, wn + OFF_W0HI * 4, wn + OFF_W0LO * 4, 0, 0, elected);
    else if (phase == 1)
        issue_layer<H / 8, H>(tn + C_S, tn + C_P, tn + C_Q, wn + OFF_W1HI * 4, wn + OFF_W1LO * 4, x + (KP - 8), wn + OFF_B1 * 4, elected);
    else
        issue_layer<H / 8, NH>(head_out, tn + C_S, tn + C_Q, wn + OFF_WHHI * 4, wn + OFF_WHLO * 4, x + (KP - 8), wn + OFF_BH * 4, elected);
    mma_commit(done_bar, elected);
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// The forward kernel.  One CTA per SM, persistent over tiles of 128 envs, three warp roles:
//   warps 0..15  compute: the tanh epilogues.  Warp w owns TMEM lanes 32 (w % 4) .. + 31 (the envs of those rows) and column
//                block w / 4 (16 of a layer's 64 columns) -- four warps per scheduler partition, which is what the epilogue
//                needs to hide its dependent FMA chains;
//   warp 16      issues every tcgen05.mma (one elected lane), in the order the operands become ready;
//   warps 17..20 io (thread = row of the tile): convert the NEXT tile's observation rows into the tf32 hi / lo A operand
//                in TMEM, and finish the PREVIOUS heads: value, DiagGaussian sample (noise drawn here), clip, log-prob,
//                coalesced stores.  Both are latency-bound, short phases; on the compute warps they were 35 % of a tile.
// Critic and actor are two independent chains of  MMA -> tanh epilogue -> MMA -> ...; the compute warps alternate
// between them, so while they run the epilogue of one network the tensor pipe runs the layer of the other:
//     tensor pipe   c0 a0 | c1      | a1 c0' a0' | ch      | ah      | c1'     | ...
//     compute             | E(c0)   | E(a0)      | E(c1)   | E(a1)   | E(c0')  | E(a0') ...
//     io            X'                                     |  value, actions (this tile)  X'' ...
// (c0 = critic layer 0, ch = critic head, ' = next tile).  Synchronisation, all mbarriers: io -> issuer `xbar[buf]` (X of a
// tile is in TMEM), compute -> issuer `rbar[net]` (tanh of a layer is in TMEM), issuer -> compute `mbar[net][layer]`
// (tcgen05.commit of a hidden layer), issuer -> io, compute `hbar[net][buf]` (commit of a head).  Every barrier that can be
// signalled again before a slow waiter has looked is doubled by tile parity: a parity wait cannot be two completions
// behind (r2: a single issuer -> compute barrier per chain stalled forever when another kernel shared the SM).
// ------------------------------------------------------------------------------------------
constexpr int IO_THREADS = TILE;
constexpr int ALL_THREADS = THREADS + 32 + IO_THREADS;
// named barrier of io set `set` (ids 2, 3)
__device__ __forceinline__ void io_barrier(int set) { asm volatile("bar.sync %0, %1;" ::"r"(2 + set), "n"(IO_THREADS) : "memory"); }

// The env step fused behind the forward pass (NCT > 0): what the step writes besides the handle's own state.
struct StepOut {
    float *obs_next;      // [E][D] the observation after the step (the rollout buffer's next slab)
    float *reward;        // [E]
    uint8_t *done;        // [E]
};

// NCT = 0: the forward pass alone.  NCT > 0: FUSED POLICY + ENV STEP for a default station of NCT spots (battery on, PV with
// three steps ahead, no requested-SoC plane: obs_dim = 8 + 2 NCT + 1, act_dim = NCT + 1).  The io warps -- thread = row = env,
// one warp = one 32-env state block, exactly the step kernel's mapping -- carry on with the env step of their tile as soon
// as its clipped actions are in shared memory: same env_step() body, same warp-cooperative admission of arrivals, so
// the results are bit-identical to sng_policy_forward_* followed by sng_step.  The env step is ~1,100 dependent-ish
// instructions per warp, longer than a tile's tensor-core work, so TWO io sets alternate over the tiles (set = parity of
// the CTA's tile number = its X buffer): a set has two tiles' time for staging X, the heads' tail and the env step.
template <int NCT, int IO_SETS>
__global__ void __launch_bounds__(ALL_THREADS + (IO_SETS - 1) * IO_THREADS, 1)
    policy_tc_kernel(const float *__restrict__ img, const float *__restrict__ obs, const float *__restrict__ noise,
                     const float *__restrict__ low, const float *__restrict__ high, float *raw_actions, float *actions,
                     float *values, float *log_probs, long long n_envs, int D, int A, int aligned, long long *trace,
                     const unsigned long long *rng_step, unsigned long long rng_offset, unsigned long long rng_seed,
                     unsigned long long rng_gid0, float *noise_out, int early, const Params<float> envp, const StepOut so)
{
    static_assert(IO_SETS == 1 || IO_SETS == 2, "one or two io sets");
    static_assert(NCT == 0 || IO_SETS == 2, "the fused env step needs two io sets");
    int tr = 0;
#define TRACE() do { if (trace && blockIdx.x == 0 && threadIdx.x == 0 && tr < 64) trace[tr++] = clock64(); } while (0)
    // the same for the first thread of io set 0 (slots 64..127) and of io set 1 (slots 128..191)
#define TRACE_IO() do { if (trace && blockIdx.x == 0 && tio == 0 && tr < 64) trace[64 * (1 + set) + tr++] = clock64(); } while (0)
    if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[255] = clock64();
    extern __shared__ __align__(128) unsigned char smem[];
    const Smem sp = smem_plan(D, A, IO_SETS, NCT > 0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool issuer = warp == THREADS / 32, io = warp > THREADS / 32;
    const int set = io ? (warp - THREADS / 32 - 1) / (IO_THREADS / 32) : 0;      // io set of this warp
    const int cb = warp >> 2;                               // column block of a compute warp in the epilogues
    const int t = (warp & 3) * 32 + lane;                   // row of the tile = TMEM lane = env (compute and io warps)
    float *obs_s0 = reinterpret_cast<float *>(smem + sp.stages);
    float *noise_s = reinterpret_cast<float *>(smem + sp.rows + (uint32_t)set * 3 * sp.row_stage);
    float *raw_s = reinterpret_cast<float *>(smem + sp.rows + (uint32_t)set * 3 * sp.row_stage + sp.row_stage);
    float *act_s = reinterpret_cast<float *>(smem + sp.rows + (uint32_t)set * 3 * sp.row_stage + 2 * sp.row_stage);
    float *oout_s = reinterpret_cast<float *>(smem + sp.out_rows + (uint32_t)set * sp.obs_stage);     // NCT > 0
    float *const_s = reinterpret_cast<float *>(smem + sp.consts);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + sp.bars);
    uint64_t *wbar = bars;                                  // [4] weight image: critic layer 0 | actor layer 0 | rest of the critic | rest of the actor
    uint64_t *obar = bars + 4;                              // [2] observation rows of the even / odd tiles have landed
    uint64_t *nbar = bars + 6;                              // noise rows have landed (io set 1: bars + 19)
    uint64_t *xbar = bars + 7;                              // [2] io -> issuer (4 warp arrivals): X of an even / odd tile is in TMEM
    uint64_t *rbar = bars + 9;                              // [2] compute -> issuer (16 warp arrivals): tanh(critic / actor layer) is in TMEM
    uint64_t *mbar = bars + 11;                             // [2][2] issuer -> compute: layer 0 / layer 1 of the critic / the actor is complete
    uint64_t *hbar = bars + 15;                             // [2][2] issuer -> io, compute: the critic's / the actor's head of an even / odd tile
    constexpr int N_BARS = 20;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + N_BARS);
    const bool value_only = actions == nullptr;
    const int n_nets = value_only ? 1 : 2;
    const long long n_tiles = (n_envs + TILE - 1) / TILE;
    const long long stride = gridDim.x;
    const long long first_tile = blockIdx.x;
    const int my_tiles = first_tile < n_tiles ? (int)((n_tiles - 1 - first_tile) / stride + 1) : 0;
    const uint32_t obs_bytes = (uint32_t)(TILE * D * sizeof(float));
    auto tile_full = [&](long long tile) { return aligned && tile < n_tiles && (n_envs - tile * TILE) >= TILE; };
    // copy-engine fetch of a whole tile's observation rows into stage `buf` (one thread)
    auto fetch_obs = [&](long long tile, int buf) {
        mbar_expect_tx(obar + buf, obs_bytes);
        bulk_g2s(reinterpret_cast<unsigned char *>(obs_s0) + (size_t)buf * sp.obs_stage, obs + tile * TILE * D, obs_bytes, obar + buf);
    };

    // the weight image, one copy per CTA, in four pieces in the order they are needed
    constexpr uint32_t L0_BYTES = (uint32_t)(OFF_W1HI * sizeof(float)), NET_BYTES = (uint32_t)(NET_FLOATS * sizeof(float));
    auto fetch_weights = [&](int piece) {       // 0: critic layer 0, 1: actor layer 0, 2: rest of the critic, 3: rest of the actor
        const uint32_t lo = (piece & 1) * NET_BYTES + (piece >= 2 ? L0_BYTES : 0u), hi = (piece & 1) * NET_BYTES + (piece >= 2 ? NET_BYTES : L0_BYTES);
        mbar_expect_tx(wbar + piece, hi - lo);
        for (uint32_t off = lo; off < hi; off += 16384u) {
            const uint32_t n = hi - off < 16384u ? hi - off : 16384u;
            bulk_g2s(smem + off, reinterpret_cast<const unsigned char *>(img) + off, n, wbar + piece);
        }
    };
    if (threadIdx.x == 0) {
        for (int i = 0; i < N_BARS; ++i)
            mbar_init(bars + i, (i == 7 || i == 8) ? IO_THREADS / 32 : ((i == 9 || i == 10) ? THREADS / 32 : 1));
        fence_mbar_init();
        // `early` (programmatic dependent launch): the weight image was complete before the predecessor started, so the
        // whole of it is requested while the predecessor is still running
        if (early) {
            fetch_weights(0);
            if (!value_only) fetch_weights(1);
            fetch_weights(2);
            if (!value_only) fetch_weights(3);
        }
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    // programmatic dependent launch: everything above ran while the predecessor (the step kernel that writes `obs`) was
    // still running; nothing below may start before it has completed.  No-ops for an ordinary launch.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        // head of the dependency chain first: the first tile's observation rows and the critic's layer 0 (the rest of the
        // image is requested after the CTA barrier by a thread that would otherwise wait: issuing a bulk copy costs the
        // issuing thread ~100 cycles, and everybody waits for this thread at the barrier)
        if (tile_full(first_tile)) fetch_obs(first_tile, 0);
        if (!early) fetch_weights(0);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {                     // compute warp 0, about to wait for the first layer anyway
        if (!early) {
            if (!value_only) fetch_weights(1);
            fetch_weights(2);
            if (!value_only) fetch_weights(3);
        }
        if (tile_full(first_tile + stride)) fetch_obs(first_tile + stride, 1);
    }
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // TMEM as seen by this warp's 32 lanes

    if (issuer) {
        // ---- the MMA issuer ----
        const uint32_t w_base = smem_u32(smem);
        const uint32_t w_net1 = w_base + (uint32_t)(NET_FLOATS * sizeof(float));
        const uint32_t mb = smem_u32(mbar), hb = smem_u32(hbar);        // mbar[net][layer] at mb + 8 (2 net + layer), hbar[net][buf] likewise
        const uint32_t tn0 = tmem + C_NET0, tn1 = tn0 + C_NETSTRIDE;
        uint32_t cp = 0u, ap = 0u;
        // layer 0 of both networks for this CTA's tile number `it`: needs only X of that tile
        auto layer0 = [&](int it) {
            const int buf = it & 1;
            const uint32_t x = tmem + x_cols(buf);
            mbar_wait(xbar + buf, (uint32_t)((it >> 1) & 1));
            if (it == 0) mbar_wait(wbar, 0);
            issue_phase(0, tn0, x, 0, w_base, mb);
            if (n_nets == 2) {
                if (it == 0) mbar_wait(wbar + 1, 0);
                issue_phase(0, tn1, x, 0, w_net1, mb + 16u);
            }
        };
        if (my_tiles > 0) layer0(0);
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = it & 1;
            const uint32_t x = tmem + x_cols(buf);
            // the heads write into the (by then unused) lo half of the tile's own X buffer
            const uint32_t head0 = x + 32u, head1 = x + 48u;
            mbar_wait(rbar, cp); cp ^= 1u;                             // tanh(critic layer 0)
            if (it == 0) mbar_wait(wbar + 2, 0);
            issue_phase(1, tn0, x, 0, w_base, mb + 8u);
            if (n_nets == 2) {
                mbar_wait(rbar + 1, ap); ap ^= 1u;                     // tanh(actor layer 0)
                if (it == 0) mbar_wait(wbar + 3, 0);
                issue_phase(1, tn1, x, 0, w_net1, mb + 24u);
            }
            // the NEXT tile's layer 0 goes in ahead of this tile's heads: the compute warps then find it complete when
            // they finish this tile's last epilogue (the heads -- 50 small MMAs -- would otherwise sit in front of it)
            if (it + 1 < my_tiles) layer0(it + 1);
            mbar_wait(rbar, cp); cp ^= 1u;                             // tanh(critic layer 1)
            issue_phase(2, tn0, x, head0, w_base, hb + 8u * (uint32_t)buf);
            if (n_nets == 2) {
                mbar_wait(rbar + 1, ap); ap ^= 1u;                     // tanh(actor layer 1)
                issue_phase(2, tn1, x, head1, w_net1, hb + 8u * (uint32_t)(2 + buf));
            }
        }
    } else if (!io) {
        // ---- compute warps: the tanh epilogues; the epilogue of one network overlaps the MMAs of the other ----
        const uint32_t tc = tlane + C_NET0, ta = tc + C_NETSTRIDE;           // critic / actor regions
#pragma unroll 1
        for (int it = 0; it < my_tiles; ++it) {
            const uint32_t ph = (uint32_t)(it & 1);                      // every mbar[net][layer] completes once per tile
            TRACE();
#pragma unroll
            for (int layer = 0; layer < 2; ++layer) {
                const uint32_t src = layer == 0 ? C_P : C_S;
                mbar_wait(mbar + layer, ph);
                // layer 0 of this tile was issued ahead of the previous tile's heads, and its epilogue overwrites Q, which
                // those heads read: wait for them (normally long complete)
                if (layer == 0 && it > 0) mbar_wait(hbar + ((it - 1) & 1), (uint32_t)(((it - 1) >> 1) & 1));
                tc_fence_after();
                TRACE();
                tanh_epilogue(tc + src + 16 * cb, tc + C_Q + 16 * cb);
                tc_fence_before();
                mbar_arrive_warp(rbar);
                TRACE();
                if (n_nets == 2) {
                    mbar_wait(mbar + 2 + layer, ph);
                    if (layer == 0 && it > 0) mbar_wait(hbar + 2 + ((it - 1) & 1), (uint32_t)(((it - 1) >> 1) & 1));
                    tc_fence_after();
                    TRACE();
                    tanh_epilogue(ta + src + 16 * cb, ta + C_Q + 16 * cb);
                    tc_fence_before();
                    mbar_arrive_warp(rbar + 1);
                    TRACE();
                }
            }
        }
    } else if (my_tiles > 0) {
        // ---- io warps ----
        const int tio = (int)threadIdx.x - (THREADS + 32) - set * IO_THREADS;   // 0 .. 127 within the set
        uint64_t *nbar_s = set ? bars + 19 : nbar;
        const uint32_t row_bytes = (uint32_t)(TILE * A * sizeof(float));
        const bool rng = !value_only && noise == nullptr && rng_step != nullptr;      // in-kernel Gaussian noise
        const bool sample = !value_only && noise != nullptr;                         // caller-supplied noise rows
        const unsigned long long step = rng ? *rng_step + rng_offset : 0ull;
        uint32_t n_phase = 0;
        if (!value_only && tio < 4 * NH) {          // per-action constants of the sampling epilogue (first read behind the io barriers below)
            const int q = tio / NH, a = tio % NH;
            float v = 0.f;
            if (a < A) v = q == 0 ? img[OFF_STD + a] : (q == 1 ? img[OFF_STD + 16 + a] : (q == 2 ? low[a] : high[a]));
            const_s[q * NH + a] = v;
        }
        // X = [obs | 0 | 1 1] of this CTA's tile number `it` -> tf32 hi / lo -> X buffer `it & 1` in TMEM (the A operand of
        // layer 0), then tell the issuer; the observation stage is refilled with the tile two further on.
        // (The X buffer is free: its previous tile's heads -- which live in its lo half -- were consumed by this warp's own
        //  tail two calls ago, after the commit that covers all of that tile's MMAs.)
        auto stage_x = [&](long long tile, int it) {
            const int buf = it & 1;
            float *obs_s = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(obs_s0) + (size_t)buf * sp.obs_stage);
            const long long e0 = tile * TILE;
            const int nv = (int)((n_envs - e0) < TILE ? (n_envs - e0) : TILE);
            const bool full = tile_full(tile);
            if (full) {
                mbar_wait(obar + buf, (uint32_t)((it >> 1) & 1));
            } else {                                                     // ragged last tile / unaligned rows: plain loads
                for (int k = tio; k < nv * D; k += IO_THREADS) obs_s[k] = obs[e0 * D + k];
                io_barrier(set);
            }
            const float *row = obs_s + t * D;
            __syncwarp();
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t xh[16], xl[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int k = 16 * half + j;
                    const float x = k < D ? (t < nv ? row[k] : 0.f) : (k >= KP - 2 ? 1.0f : 0.f);
                    split_tf32(x, xh[j], xl[j]);
                }
                tmem_st16(tlane + x_cols(buf) + 16 * half, xh);
                tmem_st16(tlane + x_cols(buf) + 32 + 16 * half, xl);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive_warp(xbar + buf);
            io_barrier(set);                                                // every io thread is done with this observation stage
            if (tio == 0 && tile_full(tile + 2 * stride)) fetch_obs(tile + 2 * stride, buf);
        };

        // set `set` owns the CTA's tiles number set, set + IO_SETS, ... (one set: all of them)
        for (int k = set; k < 2 && k < my_tiles; k += IO_SETS) stage_x(first_tile + (long long)k * stride, k);
        if (NCT) publish_dep_table_warp(envp);          // the env step's departure / arrival-gap table (per warp, no CTA barrier)

#pragma unroll 1
        for (int it = set; it < my_tiles; it += IO_SETS) {
            const long long tile = first_tile + (long long)it * stride;
            const long long e0 = tile * TILE;
            const int nv = (int)((n_envs - e0) < TILE ? (n_envs - e0) : TILE);
            const bool full = tile_full(tile);
            const int buf = it & 1;
            const uint32_t hph = (uint32_t)((it >> 1) & 1);
            if (sample) {                                                // noise_s: the previous tail ended with a barrier
                if (full) {
                    if (tio == 0) {
                        mbar_expect_tx(nbar_s, row_bytes);
                        bulk_g2s(noise_s, noise + e0 * A, row_bytes, nbar_s);
                    }
                } else {
                    for (int k = tio; k < nv * A; k += IO_THREADS) noise_s[k] = noise[e0 * A + k];
                    io_barrier(set);
                }
            }
            if (rng) {
                // In-kernel exploration noise, drawn BEFORE the heads are waited for (it does not depend on them: this is idle
                // time of the io warps): one Philox block per four action columns, keyed by (seed, step), counter = (global
                // env, column block), Box-Muller; each thread parks its own row in the noise stage
                const unsigned long long gid = rng_gid0 + (unsigned long long)(e0 + t);
#pragma unroll
                for (int q = 0; q < NH / 4; ++q) {
                    if (4 * q < A) {                                     // warp-uniform
                        uint32_t x[4];
                        float zz[4];
                        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)q, (uint32_t)step, (uint32_t)rng_seed,
                                      (uint32_t)(rng_seed >> 32) ^ (uint32_t)(step >> 32), x);
                        box_muller(x[0], x[1], zz[0], zz[1]);
                        box_muller(x[2], x[3], zz[2], zz[3]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int a = 4 * q + j;
                            if (a < A) {
                                noise_s[t * A + a] = zz[j];
                                if (noise_out && t < nv) noise_out[(e0 + t) * A + a] = zz[j];
                            }
                        }
                    }
                }
            }
            // ---- this tile's heads: value (column 32 of its X buffer), action means (columns 48..63), into registers ----
            const uint32_t xb = tlane + x_cols(buf);
            uint32_t out[16];
            mbar_wait(hbar + buf, hph);
            tc_fence_after();
            __syncwarp();
            const uint32_t value_bits = tmem_ld1(xb + 32);
            if (n_nets == 2) {
                mbar_wait(hbar + 2 + buf, hph);
                tc_fence_after();
                __syncwarp();
                tmem_ld16(xb + 48, out);
            }
            tmem_wait_ld();
            TRACE_IO();
            // The X buffer is free now (the heads' commits cover every MMA that read it, and their outputs are in
            // registers): stage the tile that uses it next -- two tiles on -- BEFORE the sampling work, so that the
            // issuer finds X of the next tile ready a whole tile early and can run its layer 0 ahead of the heads.
            if (it + 2 < my_tiles) stage_x(tile + 2 * stride, it + 2);
            TRACE_IO();
            if (t < nv) values[e0 + t] = __uint_as_float(value_bits);
            if (n_nets == 2) {
                if (sample && full) { mbar_wait(nbar_s, n_phase); n_phase ^= 1u; }
                // DiagGaussian sample, clip to the Box, log-probability (z: this thread's row of the noise stage)
                float lp = 0.f;
#pragma unroll
                for (int q = 0; q < NH / 4; ++q) {
                    if (4 * q < A) {                                     // warp-uniform
                        float part = 0.f;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int a = 4 * q + j;
                            if (a < A) {
                                const float sd = const_s[a], ls = const_s[NH + a];
                                const float z = (sample || rng) ? noise_s[t * A + a] : 0.f;
                                const float x = fmaf(z, sd, __uint_as_float(out[a]));
                                raw_s[t * A + a] = x;
                                act_s[t * A + a] = fminf(fmaxf(x, const_s[2 * NH + a]), const_s[3 * NH + a]);   // SB3 clips Box actions before env.step
                                part += -0.5f * z * z - ls - 0.91893853320467274f;                            // log N(x; mean, std)
                            }
                        }
                        lp += part;
                    }
                }
                if (t < nv) log_probs[e0 + t] = lp;
                io_barrier(set);                                            // all rows of the two slabs are in shared memory
                if (full) {                                              // coalesced 16-byte stores
                    const int nvec = (int)(row_bytes / 16);
                    const float4 *rs = reinterpret_cast<const float4 *>(raw_s), *as = reinterpret_cast<const float4 *>(act_s);
                    float4 *rg = reinterpret_cast<float4 *>(raw_actions + e0 * A), *ag = reinterpret_cast<float4 *>(actions + e0 * A);
                    for (int k = tio; k < 2 * nvec; k += IO_THREADS) {
                        if (k < nvec) rg[k] = rs[k];
                        else ag[k - nvec] = as[k - nvec];
                    }
                } else {
                    for (int k = tio; k < nv * A; k += IO_THREADS) {
                        raw_actions[e0 * A + k] = raw_s[k];
                        actions[e0 * A + k] = act_s[k];
                    }
                }
                io_barrier(set);                                            // the slabs may be overwritten by the next tile
            }
            TRACE_IO();
            if constexpr (NCT > 0) {
                // ---- the env step of this tile's 128 envs: warp = one 32-env state block, thread = env; the clipped action
                //      rows are in act_s, the next observation is assembled in oout_s and leaves as one coalesced slab ----
                typedef WordOf<float>::type word;
                const long long e = e0 + t;
                word *spot = envp.spot + (size_t)(e / kBlock) * (size_t)(NCT * kBlock) + (size_t)(e % kBlock);
                StateRegs<float, NCT> st;
                load_state<float, NCT, 1, true>(envp, e, spot, st);
                const RowIO<float, 1> rio = {act_s + t * A, act_s + t * A, oout_s + t * D, 8, 8 + NCT};
                const Arrivals arr = env_step<float, NCT, 8, false, true, true, 1, true>(envp, e, spot, st, rio, so.reward, so.done);
                TRACE_IO();
                // the admission queue reuses this warp's own action rows (every lane is done with its row: the shuffles at
                // the top of the admission are the warp's barrier)
                admit_arrivals_warp<float, NCT, true>(envp, e - lane, lane, spot - lane, arr, reinterpret_cast<uint16_t *>(act_s + (t - lane) * A));
                TRACE_IO();
                io_barrier(set);                                         // all 128 rows of the next observation are in shared memory
                {
                    const int nvec = (int)(TILE * D * sizeof(float) / 16);
                    const float4 *os = reinterpret_cast<const float4 *>(oout_s);
                    float4 *og = reinterpret_cast<float4 *>(so.obs_next + e0 * D);
                    for (int k = tio; k < nvec; k += IO_THREADS) og[k] = os[k];
                }
                io_barrier(set);                                         // the stage may be overwritten by the set's next tile
                TRACE_IO();
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(*tmem_slot, 512);
    if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[254] = clock64();
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

extern "C" void sng_policy_debug_trace(long long *device_buf) { g_trace = device_buf; }

extern "C" int sng_policy_set_launch_mode(int mode)
{
    if (mode < 0 || mode > 3) return SNG_ERR_ARG;
    g_policy_pdl = mode;
    return SNG_OK;
}

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

namespace {
// nct = 0: the forward pass alone; 4 / 10: fused with the env step of that default station (envp, so)
int launch_policy_tc(const void *packed, int obs_dim, int act_dim, const float *obs, const float *noise, const float *low,
                     const float *high, float *raw_actions, float *actions, float *values, float *log_probs, int64_t n_envs,
                     const unsigned long long *rng_step, unsigned long long rng_offset, unsigned long long rng_seed,
                     unsigned long long rng_gid0, float *noise_out, void *stream, int nct = 0, const Params<float> *envp = nullptr,
                     const StepOut *so = nullptr)
{
    if (!packed || !obs || !values || n_envs < 1) return SNG_ERR_ARG;
    if (actions && (!raw_actions || !log_probs || !low || !high)) return SNG_ERR_ARG;
    if (obs_dim < 1 || obs_dim > KP - 2 || act_dim < 1 || act_dim > NH) return SNG_ERR_UNSUPPORTED;
    PointerDeviceGuard guard(obs);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    // two io sets whenever the io warps carry more than the plain tail: the env step, or drawing the exploration noise
    // (Philox + Box-Muller: ~500 instructions per thread and tile; measured on C3: 29.9 -> 28.1 us per rollout step)
    const int sets = (nct || (actions && rng_step)) ? 2 : 1;
    const Smem sp = smem_plan(obs_dim, act_dim, sets, nct > 0);
    size_t smem = sp.bars + 176;   // 20 mbarriers + the TMEM address slot
    const void *kern = nct == 10 ? (const void *)policy_tc_kernel<10, 2> : (nct == 4 ? (const void *)policy_tc_kernel<4, 2> :
                       (sets == 2 ? (const void *)policy_tc_kernel<0, 2> : (const void *)policy_tc_kernel<0, 1>));
    if (g_policy_pdl & 2) {
        // the CTA claims its SM's whole shared memory: no CTA of a kernel launched early behind this one (programmatic
        // dependent launch) can become resident next to it and take issue slots from the compute warps
        int optin = 0;
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if ((size_t)optin > smem) smem = (size_t)optin;
    }
    if (ensure_smem(dev, kern, smem) != SNG_OK) return SNG_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long n_tiles = (n_envs + TILE - 1) / TILE;
    long long grid = n_tiles;
    if (grid > sms) grid = sms;
    int aligned = aligned16(obs) && aligned16(noise) && aligned16(raw_actions) && aligned16(actions) && aligned16(packed);
    if (!aligned16(packed)) return SNG_ERR_ARG;
    Params<float> ep;
    StepOut eo = {nullptr, nullptr, nullptr};
    memset(&ep, 0, sizeof(ep));
    if (nct) {
        // the fused kernel handles whole, aligned tiles only (the caller falls back to two launches otherwise)
        if (!envp || !so || !actions || !aligned || !aligned16(so->obs_next) || n_envs % TILE != 0) return SNG_ERR_UNSUPPORTED;
        ep = *envp;
        eo = *so;
    }
    const int pdl = g_policy_pdl & 1;
    long long nenv = (long long)n_envs;
    int early = pdl;
    const float *img = reinterpret_cast<const float *>(packed);
    void *args[] = {&img, &obs, &noise, &low, &high, &raw_actions, &actions, &values, &log_probs, &nenv, &obs_dim, &act_dim, &aligned,
                    &g_trace, &rng_step, &rng_offset, &rng_seed, &rng_gid0, &noise_out, &early, &ep, &eo};
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3(ALL_THREADS + (sets - 1) * IO_THREADS); lc.dynamicSmemBytes = smem; lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelExC(&lc, kern, args) == cudaSuccess ? SNG_OK : SNG_ERR_CUDA;
}
}  // namespace

// Fused policy forward + env step (called by Engine<float>::policy_step through sng_policy_step).
namespace sng {
int launch_policy_step(const Params<float> &p, const PolicyStepArgs &a, cudaStream_t st)
{
    const StepOut so = {a.obs_next, a.reward, a.done};
    return launch_policy_tc(a.packed, p.D, p.A, a.obs, a.noise, a.low, a.high, a.raw_actions, a.actions, a.values, a.log_probs, p.n_envs,
                            reinterpret_cast<const unsigned long long *>(a.step_counter), a.step_offset, a.seed, p.gid0, a.noise_out, (void *)st,
                            p.N, &p, &so);
}
}  // namespace sng

extern "C" int sng_policy_forward_packed(const void *packed, int obs_dim, int act_dim, const float *obs, const float *noise,
                                         const float *low, const float *high, float *raw_actions, float *actions,
                                         float *values, float *log_probs, int64_t n_envs, void *stream)
{
    return launch_policy_tc(packed, obs_dim, act_dim, obs, noise, low, high, raw_actions, actions, values, log_probs, n_envs,
                            nullptr, 0ull, 0ull, 0ull, nullptr, stream);
}

extern "C" int sng_policy_forward_sampled(const void *packed, int obs_dim, int act_dim, const float *obs, uint64_t seed,
                                          const uint64_t *step_counter, uint64_t step_offset, uint64_t env_gid0,
                                          const float *low, const float *high, float *raw_actions, float *actions,
                                          float *values, float *log_probs, float *noise_out, int64_t n_envs, void *stream)
{
    if (!step_counter || !actions) return SNG_ERR_ARG;
    return launch_policy_tc(packed, obs_dim, act_dim, obs, nullptr, low, high, raw_actions, actions, values, log_probs, n_envs,
                            reinterpret_cast<const unsigned long long *>(step_counter), step_offset, seed, env_gid0, noise_out, stream);
}
