// sng_engine.cuh -- kernels and host-side launch logic, templated on the arithmetic type.
// Instantiated as Engine<float,false> (sng_api.cu) and Engine<double,true> (sng_f64.cu, built
// with -fmad=false so that no multiply-add is contracted).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/sng.h"
#include "sng_device.cuh"
#include "sng_tma.cuh"

namespace sng {

// sng_policy_step: the actor-critic forward pass of one rollout step fused with the env step (sng_policy_tc.cu)
struct PolicyStepArgs {
    const void *packed;                 // sng_policy_pack image
    const float *obs;                   // [E][D] what the policy sees (rollout slab s)
    const float *noise;                 // [E][A] standard normals, or null: drawn in the kernel (seed, step_counter, step_offset)
    const uint64_t *step_counter;
    uint64_t step_offset, seed;
    const float *low, *high;            // [A]
    float *raw_actions, *actions, *values, *log_probs, *noise_out;
    float *obs_next, *reward;           // [E][D] (rollout slab s + 1), [E]
    uint8_t *done;                      // [E]
};
int launch_policy_step(const Params<float> &p, const PolicyStepArgs &a, cudaStream_t st);

struct EngineBase {
    virtual ~EngineBase() {}
    virtual int policy_step(const PolicyStepArgs &a, cudaStream_t st) = 0;
    virtual int bind(const sng_buffers *b) = 0;
    virtual int reset(uint64_t seed, const uint8_t *mask, int reset_battery, cudaStream_t st) = 0;
    virtual int load_schedule(const sng_schedule_view *v, cudaStream_t st) = 0;
    virtual int step(cudaStream_t st) = 0;
    virtual int rollout(const void *actions, float *obs, void *reward, uint8_t *done, int n_steps, cudaStream_t st) = 0;
    virtual int step_host(const void *a, float *obs, void *rew, uint8_t *done, cudaStream_t st) = 0;
    virtual int sample_plan(cudaStream_t st) = 0;
    virtual int sample_actions(uint64_t seed, uint64_t step0, int n_steps, void *actions, cudaStream_t st) = 0;
    virtual int error_flags(uint32_t *out, cudaStream_t st) = 0;
    virtual int probe_arrival_gap(const uint32_t *x, uint32_t *g, long long n, cudaStream_t st) = 0;
    virtual int traffic_skeleton(int variant, cudaStream_t st) = 0;
    virtual int set_tuning(int warps_per_cta, int use_generic, int use_bulk, int host_chunks) = 0;
    virtual int set_pipeline(int kernel_variant, int ctas_per_sm) = 0;
    virtual int set_launch_mode(int mode) = 0;
    int64_t launches = 0;
    std::string error;
};

EngineBase *make_engine_f32(const sng_config &cfg, int device, std::string &err);
EngineBase *make_engine_f64(const sng_config &cfg, int device, std::string &err);

// Makes `dev` the current device for the lifetime of the guard and restores the caller's afterwards
// (a process may drive several GPUs, one handle each).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
};

#define SNG_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (call);                                                                \
        if (_e != cudaSuccess) {                                                                \
            error = std::string(#call) + ": " + cudaGetErrorString(_e);                         \
            return SNG_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

// ------------------------------------------------------------------------------------------
// Step kernels.  One warp = one block of 32 consecutive envs, one thread per env; warps are
// fully independent (the step kernel has no CTA-wide barrier: every warp publishes the 1 KB departure table itself).  Action rows arrive and
// observation rows leave through shared memory, moved by the copy engine (cp.async.bulk, mbarrier
// complete_tx / bulk_group) or by 16-byte coalesced vector accesses (STAGE_* below); the blocked state is
// read and written with coalesced 128-byte warp accesses; env_step() is the same body everywhere.
//
//   step_simple_kernel     THE production kernel: one block per warp, all loads issued up front, latency
//                          hidden by 32 resident warps per SM.  Also the rollout (n_steps > 1: the same
//                          warp advances its envs n_steps times, one action / obs / reward / done slab
//                          per step), the partial last block, and every row-staging mode.
//   step_pipelined_kernel  an alternative kept for comparison: persistent warps walk the blocks
//                          grid-stride and software-pipeline them (state words of block k+1 loading
//                          into registers, its action rows into a second shared-memory stage, while
//                          block k is computed).  Measured slower: the prefetch registers
//                          cost a third of the resident warps, and the step is bound by per-warp
//                          instruction latency rather than by DRAM latency.
// ------------------------------------------------------------------------------------------
template <typename real> __device__ __forceinline__ void publish_dep_table(const Params<real> &p)
{
    float4 *tab = reinterpret_cast<float4 *>(dep_table_smem());
#pragma unroll 1
    for (int k = threadIdx.x; k < kSmemTab / 4; k += blockDim.x) tab[k] = __ldg(reinterpret_cast<const float4 *>(p.dep_norm) + k);
    __syncthreads();
}

// The same without a CTA barrier: every warp writes the WHOLE table itself (identical values, so the warps of a CTA may
// overwrite each other freely) and relies only on its own stores, ordered by __syncwarp.  The warps of a CTA then never
// wait for one another.
template <typename real> __device__ __forceinline__ void publish_dep_table_warp(const Params<real> &p)
{
    float4 *tab = reinterpret_cast<float4 *>(dep_table_smem());
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < kSmemTab / 4 / 32; ++k) tab[lane + 32 * k] = __ldg(reinterpret_cast<const float4 *>(p.dep_norm) + lane + 32 * k);
    __syncwarp();
}
static_assert(kSmemTab % 128 == 0, "publish_dep_table_warp copies whole float4 x 32 rounds");

#ifndef SNG_PIPE_THREADS
#define SNG_PIPE_THREADS 64
#endif
#ifndef SNG_PIPE_MINB
#define SNG_PIPE_MINB 10
#endif
template <typename real, int NCT, int ND, bool EXACT>
__global__ void __launch_bounds__(SNG_PIPE_THREADS, SNG_PIPE_MINB)
    step_pipelined_kernel(const Params<real> p, long long n_blocks)
{
    extern __shared__ __align__(128) unsigned char smem[];
    typedef typename WordOf<real>::type word;
    publish_dep_table(p);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int N = NCT ? NCT : p.N;
    const int A = p.A, D = p.D;
    const uint32_t act_bytes = (uint32_t)(kBlock * A * sizeof(real)), obs_bytes = (uint32_t)(kBlock * D * sizeof(float));
    const uint32_t act_stage = align128(act_bytes);
    const uint32_t per_warp = 2 * act_stage + align128(obs_bytes);
    unsigned char *wbase = smem + (size_t)warp * per_warp;
    float *obs_s = reinterpret_cast<float *>(wbase + 2 * act_stage);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (size_t)wpb * per_warp) + 2 * warp;   // one per action stage

    const long long stride = (long long)gridDim.x * wpb;
    long long blk = (long long)blockIdx.x * wpb + warp;
    if (blk >= n_blocks) return;
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        fence_mbar_init();
    }
    __syncwarp();
    const size_t blk_words = (size_t)N * kBlock;
    StateRegs<real, NCT> cur, nxt;
    load_state<real, NCT>(p, blk * kBlock + lane, p.spot + (size_t)blk * blk_words + lane, cur);
    if (lane == 0) {
        mbar_expect_tx(bar, act_bytes);
        bulk_g2s(wbase, p.actions + (size_t)blk * kBlock * A, act_bytes, bar);
    }
    for (int it = 0; blk < n_blocks; ++it, blk += stride) {
        const long long nblk = blk + stride;
        const int stage = it & 1;
        if (nblk < n_blocks) {   // warp-uniform: put block k+1 in flight
            load_state<real, NCT>(p, nblk * kBlock + lane, p.spot + (size_t)nblk * blk_words + lane, nxt);
            if (lane == 0) {
                mbar_expect_tx(bar + (stage ^ 1), act_bytes);
                bulk_g2s(wbase + (stage ^ 1) * act_stage, p.actions + (size_t)nblk * kBlock * A, act_bytes, bar + (stage ^ 1));
            }
        }
        mbar_wait(bar + stage, (uint32_t)((it >> 1) & 1));
        const long long e = blk * kBlock + lane;
        const real *act_row = reinterpret_cast<const real *>(wbase + stage * act_stage) + lane * A;
        const RowIO<real> io = {act_row, act_row, obs_s + lane * D, Offsets<NCT, ND>::soc(p), Offsets<NCT, ND>::dep(p)};
        env_step<real, NCT, ND, EXACT, true>(p, e, p.spot + (size_t)blk * blk_words + lane, cur, io, p.reward, p.done);
        __syncwarp();            // obs rows complete; everyone is done reading this action stage
        {   // 32 rows = 8 * D float4 (16-byte aligned on both sides): coalesced 512-byte warp stores.
            // (Plain stores, not cp.async.bulk: waiting for a bulk group drains the scoreboard the
            //  prefetched state loads sit on, which would serialise the pipeline.)
            const float4 *src = reinterpret_cast<const float4 *>(obs_s);
            float4 *dst = reinterpret_cast<float4 *>(p.obs + (size_t)blk * kBlock * D);
#pragma unroll 4
            for (int k = lane; k < 8 * D; k += 32) dst[k] = src[k];
        }
        __syncwarp();            // obs_s may be overwritten by the next block
        cur = nxt;
    }
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor in the stream is still running; griddepcontrol.wait blocks until the predecessor has completed
// and its writes are visible (a no-op for ordinary launches), griddepcontrol.launch_dependents lets the successor start.
// A kernel releases its dependents only AFTER its own wait, so completion is transitive along the stream.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifndef SNG_STEP_MAXT
#define SNG_STEP_MAXT 128
#endif
#ifndef SNG_STEP_MINB
#define SNG_STEP_MINB 8
#endif
#ifndef SNG_STEP_MINB_LANES
#define SNG_STEP_MINB_LANES 5     // several lanes per env, <= 16 spots per lane: 96 registers without spills, 20 warps per SM (measured: 6 -> 80 registers + spills 0.106 ms, 5 -> 0.093 ms, 4 -> 0.098 ms, 8 -> 0.143 ms per C5 step)
#endif
// Row staging modes (kernel argument `mode`): how the 32 action rows reach shared memory and the 32
// observation rows leave it.
enum : int {
    STAGE_SCALAR = 0,     // plain 4-byte loads / stores: unaligned buffers; always used for a partial last block
    STAGE_TMA_LOAD = 1,   // actions: one cp.async.bulk per warp (mbarrier complete_tx)
    STAGE_TMA_STORE = 2,  // observations: one cp.async.bulk per warp (bulk_group)
    STAGE_ALIGNED = 4,    // buffers are 16-byte aligned: 16-byte coalesced vector loads / stores where no TMA bit is set
    STAGE_PDL_EARLY = 8   // programmatic dependent launch whose predecessor does not touch this handle's state: the state
                          // loads are issued BEFORE waiting for the predecessor (sng_set_launch_mode(env, 2))
};

// MULTI: n_steps > 1 (rollout); the single-step instantiation carries no slab arithmetic.
// L: lanes per env (1; 4 or 2 for large specialised stations: a warp then covers 8 or 16 envs, needs a quarter / half
// the shared memory per warp and of the work per thread, so more warps stay resident.  L > 1 handles whole
// 32-env state blocks only -- the host sends a ragged last block to the L = 1 instantiation, whose sums
// associate identically).
// FIXED: battery on and no requested-SoC plane, known at compile time (the reference's default station): the row widths
// A and D and every shared-memory offset are then constants.
template <typename real, int NCT, int ND, bool EXACT, bool MULTI, int L, bool FIXED>
__global__ void __launch_bounds__(SNG_STEP_MAXT, (EXACT || NCT / L > 32) ? 2 : (NCT / L > 16 ? 4 : ((L > 1 && NCT > 16) ? SNG_STEP_MINB_LANES : SNG_STEP_MINB)))   // large rows: shared memory bounds occupancy, not registers
    step_simple_kernel(const Params<real> p, const real *actions, float *obs_out, real *reward, uint8_t *done, int n_steps,
                       int mode)
{
    constexpr int EPW = kBlock / L;   // envs per warp
    extern __shared__ __align__(128) unsigned char smem[];
    typedef typename WordOf<real>::type word;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int n_envs = (int)p.n_envs;                            // a handle owns < 2^31 envs (sng_create)
    const int blk = blockIdx.x * wpb + warp;                     // this warp's group of EPW envs (L = 1: a state block)
    const int e0 = blk * EPW;
    const bool active = e0 < n_envs;                             // warp-uniform; idle warps leave at once
    const int N = NCT ? NCT : p.N;
    const int A = FIXED ? NCT + 1 : p.A, D = FIXED ? ND + 2 * NCT + 1 : p.D;
    const uint32_t act_bytes = (uint32_t)(EPW * A * sizeof(real)), obs_bytes = (uint32_t)(EPW * D * sizeof(float));
    // the rollout keeps TWO action stages: the copy engine fetches the next step's rows while this step is computed
    constexpr uint32_t NSTAGE = MULTI ? 2u : 1u;
    const uint32_t act_stage = align128(act_bytes);
    const uint32_t per_warp = NSTAGE * act_stage + align128(obs_bytes);
    // [one mbarrier per warp (128-byte header) | per warp: action rows (x NSTAGE), observation rows]
    unsigned char *wbase = smem + 128 + (size_t)warp * per_warp;
    real *act_s0 = reinterpret_cast<real *>(wbase);
    float *obs_s = reinterpret_cast<float *>(wbase + NSTAGE * act_stage);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem) + warp;
    static_assert(SNG_STEP_MAXT / 32 * 8 <= 128, "the mbarrier header holds one word per warp");
    constexpr bool COOP = !EXACT && NCT > 0 && NCT <= 16 && L == 1;   // warp-cooperative admission of arriving vehicles
    // (its queue reuses the step's action stage: every lane is done with its action row when the admission starts --
    //  the warp-wide shuffles at its top are the barrier -- and 32 * N * 2 B <= 32 * A * sizeof(real))

    const int n_valid = active ? min(EPW, n_envs - e0) : 0;   // envs of this warp (L = 2: always EPW, see above)
    const bool full = n_valid == EPW;
    const bool tma_load = full && (mode & STAGE_TMA_LOAD), tma_store = full && (mode & STAGE_TMA_STORE);
    const bool vec = full && (mode & STAGE_ALIGNED);
    const int el = lane % EPW, sub = lane / EPW;                 // env within the warp's group, lane within the env
    const bool valid = el < n_valid;
    const int e = e0 + el;
    // (this env, the lane's first spot, plane 0) in the plane-major, 32-env blocked state array
    word *spot = p.spot + (size_t)(e / kBlock) * (size_t)(N * kBlock) + (e % kBlock) + sub * kBlock;
    // float4 per lane of a group's action rows held in registers by the vector path (specialised kernels)
    constexpr int AV = (NCT && sizeof(real) == 4) ? ((NCT + 1) * EPW / 4 + 31) / 32 : 1;
    const int act_vec = (int)(act_bytes / 16);                    // float4 per block of action rows

#pragma unroll 1
    for (int s = 0; s < (MULTI ? n_steps : 1); ++s) {
        const size_t slab = MULTI ? (size_t)s * (size_t)n_envs : 0;   // row offset of this step's slab
        const real *act_g = actions + (slab + (size_t)e0) * A;
        float *obs_g = obs_out + (slab + (size_t)e0) * D;
        StateRegs<real, NCT / L> st;
        float4 areg[AV];
        // this step's action rows: with the copy engine the rollout alternates between the two stages
        real *act_s = (MULTI && tma_load && (s & 1)) ? reinterpret_cast<real *>(wbase + act_stage) : act_s0;
        uint16_t *queue = reinterpret_cast<uint16_t *>(act_s);
        // ---- put everything this block needs in flight first ----
        if (MULTI && s > 0) {
            // the stages are reused: the previous step's bulk store must have read obs_s before any lane writes it
            // again, and every lane's generic-proxy writes to its action stage (the admission queue) are ordered
            // before the copy engine refills that stage (requested below, behind a __syncwarp)
            if (tma_load) fence_proxy_async();
            if (tma_store && lane == 0) bulk_wait_read<0>();
            __syncwarp();
        }
        // the state loads go first: they have the longest way (DRAM), and every instruction ahead of them is time
        // the warp holds its slot with nothing in flight
        const bool early = !MULTI && (mode & STAGE_PDL_EARLY);
        if (!MULTI && !early) { pdl_wait(); pdl_launch_dependents(); }     // ordinary launch: no-ops
        if (valid) load_state<real, NCT / L, L, FIXED>(p, e, spot, st);
        if (early) { pdl_wait(); pdl_launch_dependents(); }                // actions (and everything written below) after the predecessor
        if (tma_load) {
            if (lane == 0 && s == 0) {         // (the rows of the later steps of a rollout were requested one step ahead, below)
                mbar_init(bar, 1);
                fence_mbar_init();
                mbar_expect_tx(bar, act_bytes);
                bulk_g2s(act_s, act_g, act_bytes, bar);
            }
        } else if (vec && NCT && sizeof(real) == 4) {
#pragma unroll
            for (int j = 0; j < AV; ++j) {
                const int k = lane + 32 * j;
                if (k < act_vec) areg[j] = reinterpret_cast<const float4 *>(act_g)[k];
            }
        }
        if (!active) return;
        if (s == 0) publish_dep_table_warp(p);     // per warp, no CTA barrier; its __syncwarp also orders the mbarrier init before the other lanes' waits
        // ---- action rows into shared memory ----
        if (tma_load) {
            if (MULTI) __syncwarp();
            mbar_wait(bar, (uint32_t)(s & 1));
            if (MULTI && s + 1 < n_steps && lane == 0) {
                // the next step's rows, into the other stage: its last users (the action reads and the admission queue
                // of step s - 1) are complete and fenced (top of this iteration)
                mbar_expect_tx(bar, act_bytes);
                bulk_g2s(reinterpret_cast<unsigned char *>(act_s0) + ((s + 1) & 1) * act_stage,
                         actions + ((size_t)(s + 1) * (size_t)n_envs + (size_t)e0) * A, act_bytes, bar);
            }
        } else if (vec) {
            if (NCT && sizeof(real) == 4) {
#pragma unroll
                for (int j = 0; j < AV; ++j) {
                    const int k = lane + 32 * j;
                    if (k < act_vec) reinterpret_cast<float4 *>(act_s)[k] = areg[j];
                }
            } else {
#pragma unroll 1
                for (int k = lane; k < act_vec; k += 32) reinterpret_cast<float4 *>(act_s)[k] = reinterpret_cast<const float4 *>(act_g)[k];
            }
            __syncwarp();
        } else {
#pragma unroll 1
            for (int k = lane; k < n_valid * A; k += 32) act_s[k] = act_g[k];
            __syncwarp();
        }
        Arrivals arrivals = {0u, 0u, 0};
        if (valid) {
            const RowIO<real, L> io = {act_s + el * A + sub, act_s + el * A, obs_s + el * D, Offsets<NCT, ND>::soc(p) + sub,
                                       Offsets<NCT, ND>::dep(p) + sub};
            arrivals = env_step<real, NCT, ND, EXACT, true, COOP, L, FIXED>(p, e, spot, st, io, reward + slab, done + slab, sub);
        }
        if (COOP) admit_arrivals_warp<real, NCT, FIXED>(p, e0, lane, spot - lane, arrivals, queue);
        // ---- observation rows out of shared memory ----
        if (tma_store) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                bulk_s2g(obs_g, obs_s, obs_bytes);
                bulk_commit();
            }
        } else if (vec) {
            // EPW rows = EPW / 4 * D float4 (16-byte aligned on both sides): coalesced 512-byte warp stores
            __syncwarp();
            const float4 *src = reinterpret_cast<const float4 *>(obs_s);
            float4 *dst = reinterpret_cast<float4 *>(obs_g);
            if (NCT && ND) {
                // D = ND + 2 NCT (+ 1 with a battery): the trip count is known up to the tail, so the loop is
                // straight-line code with one predicated store at the end (a runtime bound costs a peeled
                // remainder loop that only some lanes take)
                constexpr int DV0 = (EPW / 4) * (ND + 2 * NCT), FULL = DV0 / 32;
#pragma unroll
                for (int j = 0; j < FULL; ++j) dst[lane + 32 * j] = src[lane + 32 * j];
#pragma unroll
                for (int j = FULL; j < FULL + 2; ++j) {
                    const int k = lane + 32 * j;
                    if (k < (EPW / 4) * D) dst[k] = src[k];
                }
            } else {
#pragma unroll 4
                for (int k = lane; k < (EPW / 4) * D; k += 32) dst[k] = src[k];
            }
            if (MULTI) __syncwarp();
        } else {
            __syncwarp();
#pragma unroll 1
            for (int k = lane; k < n_valid * D; k += 32) obs_g[k] = obs_s[k];
            __syncwarp();
        }
    }
    if (tma_store && lane == 0) bulk_wait_read<0>();   // shared memory must stay valid until the store has read it
}

}  // namespace sng
#include "sng_lanes.cuh"   // step_lanes_kernel: one lane per charging spot, for latency-bound batches
namespace sng {

// Measurement hook (sng_debug_traffic_skeleton): the step kernel's memory traffic WITHOUT its arithmetic -- same launch
// geometry, shared-memory footprint and register cap (so the same 32 resident warps per SM), the same loads (per-spot
// header and SoC words, env scalars, the block's action rows through the copy engine) and the same stores (SoC words,
// env scalars, reward, done flag, the block's observation rows as coalesced 16-byte stores).  Its duration is what the
// access pattern alone costs on this machine: the practical ceiling under the step kernel, next to the copy-bandwidth
// roofline.  State is written back unchanged (variant 3 overwrites it: reset the handle afterwards); obs / reward / done
// receive meaningless values.  Default station, whole 32-env blocks only.
// `variant` (what-if patterns, same byte counts): 0 the step kernel's own (plane-major state: a block's ten header lines
// and its ten SoC lines are each 1.25 KB contiguous); 1 the state planes interleaved per spot ([block][spot][3][32]: the
// layout of rounds 1-2, two lines read out of every three); 2 loads only; 3 stores only; 4 observation rows through the
// copy engine; 5 like 1 with two planes per spot.
template <typename real, int NCT>
__global__ void __launch_bounds__(SNG_STEP_MAXT, SNG_STEP_MINB) traffic_skeleton_kernel(const Params<real> p, int variant)
{
    extern __shared__ __align__(128) unsigned char smem[];
    typedef typename WordOf<real>::type word;
    constexpr int A = NCT + 1, D = 8 + 2 * NCT + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int blk = blockIdx.x * wpb + warp;
    const int e0 = blk * kBlock;
    if (e0 + kBlock > (int)p.n_envs) return;
    constexpr uint32_t act_bytes = (uint32_t)(kBlock * A * sizeof(real)), obs_bytes = (uint32_t)(kBlock * D * sizeof(float));
    constexpr uint32_t per_warp = align128(act_bytes) + align128(obs_bytes);
    unsigned char *wbase = smem + 128 + (size_t)warp * per_warp;
    real *act_s = reinterpret_cast<real *>(wbase);
    float *obs_s = reinterpret_cast<float *>(wbase + align128(act_bytes));
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem) + warp;
    const int e = e0 + lane;
    word *spot = p.spot + (size_t)blk * (size_t)(NCT * kBlock) + lane;
    int sp = kBlock;                                                               // words between a block's spots
    size_t so = PL_SOC * (size_t)p.plane;                                          // from a spot's header word to its SoC word
    if (variant == 1 || variant == 5) {                                            // planes interleaved per spot: [block][spot][plane][32]
        const int np = variant == 1 ? kPlanes : 2;
        spot = p.spot + (size_t)blk * (size_t)(NCT * np * kBlock) + lane;
        sp = np * kBlock;
        so = kBlock;
    }
    const bool do_ld = variant != 3, do_st = variant != 2;
    word h[NCT], sw[NCT];
    EnvSt<real> es;
    memset(&es, 0, sizeof(es));
    if (do_ld) {
#pragma unroll
        for (int j = 0; j < NCT; ++j) {
            h[j] = spot[(size_t)j * sp];
            sw[j] = spot[(size_t)j * sp + so];
        }
        es = p.envst[e];
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
            mbar_expect_tx(bar, act_bytes);
            bulk_g2s(act_s, p.actions + (size_t)e0 * A, act_bytes, bar);
        }
        __syncwarp();
        mbar_wait(bar, 0u);
    } else {
#pragma unroll
        for (int j = 0; j < NCT; ++j) { h[j] = (word)j; sw[j] = (word)lane; }
    }
    word acc = 0;
    real asum = 0;
#pragma unroll
    for (int j = 0; j < NCT; ++j) acc |= h[j] & sw[j];
    if (do_ld) {
#pragma unroll
        for (int k = 0; k < A; ++k) asum += act_s[lane * A + k];
    }
    if (!do_st) {
        if (acc == 0x12345u && asum == (real)7) p.done[e] = 1;    // keeps the loads alive; never true in practice
        return;
    }
#pragma unroll
    for (int j = 0; j < NCT; ++j) spot[(size_t)j * sp + so] = sw[j];
    p.envst[e] = es;
    p.reward[e] = asum;
    p.done[e] = (uint8_t)(acc & 1u);
#pragma unroll
    for (int k = 0; k < D; ++k) obs_s[lane * D + k] = (float)asum;
    if (variant == 4) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(p.obs + (size_t)e0 * D, obs_s, obs_bytes);
            bulk_commit();
            bulk_wait_read<0>();
        }
        return;
    }
    __syncwarp();
    const float4 *src = reinterpret_cast<const float4 *>(obs_s);
    float4 *dst = reinterpret_cast<float4 *>(p.obs + (size_t)e0 * D);
#pragma unroll
    for (int j = 0; j < (8 * D + 31) / 32; ++j) {
        const int k = lane + 32 * j;
        if (k < 8 * D) dst[k] = src[k];
    }
}

// Test hook (sng_debug_arrival_gap): the table-based geometric gap of the step kernels, evaluated on given words with
// this handle's own threshold table.
template <typename real>
__global__ void __launch_bounds__(128) gap_probe_kernel(const Params<real> p, const uint32_t *x, uint32_t *g, long long n)
{
    publish_dep_table(p);
    const uint32_t base = dep_table_base();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        g[i] = geometric_gap_tab(x[i], base);
}

template <typename real>
__global__ void __launch_bounds__(256) reset_kernel(const Params<real> p, const uint8_t *mask, int init,
                                                   int new_episode, int reset_battery)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.n_envs) return;
    if (mask && !mask[e]) return;
    EnvSt<real> es = p.envst[e];
    uint32_t episode = es.t_ep >> 8;
    real soc_b = es.soc_b, shift = es.pv_shift;
    if (init) {
        episode = 0;
        shift = 1;
    } else if (new_episode) {
        episode = (episode + 1u) & 0xFFFFFFu;
    }
    if (init || reset_battery) soc_b = p.batt ? p.b_soc0 : (real)0;
    if (p.mode == MODE_SAMPLE) shift = sample_pv_shift(p, p.N, p.gid0 + (unsigned long long)e, episode);
    typename WordOf<real>::type *spot = p.spot + (size_t)(e / kBlock) * p.N * kBlock + (size_t)(e % kBlock);
    begin_episode<real, 0, 0, false>(p, p.N, e, spot, episode, shift, soc_b, p.obs + (size_t)e * p.D);
    es.soc_b = soc_b;
    es.pv_shift = shift;
    es.ep_ret = 0;
    es.t_ep = episode << 8;
    p.envst[e] = es;
    if (init && p.err) p.err[e] = 0;
}

// Whole-day schedule of the current episode, one thread per (env, spot): walks the same chain of
// Philox blocks the in-step sampler follows (first arrival, then vehicle -> next arrival).
template <typename real>
__global__ void __launch_bounds__(256) sample_plan_kernel(const Params<real> p, PlanRec<real> *plan)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid >= p.n_envs * p.N) return;
    const long long e = gtid / p.N;
    const int i = (int)(gtid % p.N);
    const uint32_t episode = p.envst[e].t_ep >> 8;
    PlanRec<real> *pl = plan + (size_t)gtid * kMaxVehicles;
    int nv = 0;
    uint32_t next = first_arrival(p, p.N, e, i, episode);
    while (next != kNoVehicle && nv < kMaxVehicles) {
        const Vehicle<real> v = fetch_vehicle(p, p.N, e, i, episode, (int)next);
        PlanRec<real> r;
        memset(&r, 0, sizeof(r));
        r.hdr = v.hdr; r.soc0 = v.soc0; r.req = v.req;
        pl[nv++] = r;
        next = v.hdr >> 24;
    }
    for (int v = nv; v < kMaxVehicles; ++v) {
        PlanRec<real> z;
        memset(&z, 0, sizeof(z));
        z.hdr = make_hdr(kNoVehicle, 0, 0, kNoVehicle);
        pl[v] = z;
    }
}

// Uniform random actions from the action box (the random policy of BASELINE config 2; reference: env.action_space.sample(),
// envs/smart_nanogrid_environment.py:101-118 for the bounds): thread = (step, env, group of four action columns), one
// Philox4x32-10 block each, key = the caller's seed, counter = (global env id, step, kCtrActions | group) -- disjoint from the
// schedule sampler's counters, and a function of the GLOBAL env id, so a sharded batch draws the same actions.
constexpr uint32_t kCtrActions = 0xAC710000u;
template <typename real>
__global__ void __launch_bounds__(256) sample_actions_kernel(const Params<real> p, real *actions, unsigned long long seed,
                                                             unsigned long long step0, int n_steps)
{
    const int groups = (p.A + 3) / 4;
    const long long total = (long long)n_steps * p.n_envs * groups;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % groups);
        const long long row = i / groups, e = row % p.n_envs, s = row / p.n_envs;
        const unsigned long long gid = p.gid0 + (unsigned long long)e, step = step0 + (unsigned long long)s;
        uint32_t x[4];
        philox4x32_10((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)step, kCtrActions | (uint32_t)g, (uint32_t)seed, (uint32_t)(seed >> 32), x);
        real *a = actions + (size_t)row * p.A;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = 4 * g + j;
            if (k < p.A) {
                const float u = __fmul_rn((float)(x[j] >> 8), 5.9604644775390625e-08f);            // [0, 1), 24 bits
                const float low = (k < p.N && !p.v2x) ? 0.0f : -1.0f;                               // chargers: [0 | -1, 1); battery: [-1, 1)
                a[k] = (real)__fmaf_rn(1.0f - low, u, low);
            }
        }
    }
}

static __global__ void or_reduce_kernel(const uint32_t *err, long long n, uint32_t *out)
{
    uint32_t v = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x)
        v |= err[k];
    v = __reduce_or_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicOr(out, v);
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
template <typename real, bool EXACT> class Engine : public EngineBase {
public:
    sng_config cfg;
    Params<real> p;
    sng_buffers buf;
    bool bound = false, started = false;
    int device = 0;
    int warps_per_cta = 0;   // 0 = auto
    int use_generic = 0, use_bulk = 1, host_chunks = 0, ctas_per_sm = 0;
    int launch_mode = 0;      // 0 ordinary launches, 1 programmatic dependent launch, 2 the same with the state loads ahead of the wait
    int kernel_variant = 0;   // 0 / 2 one block per warp, 1 persistent pipelined
    int lanes_kernel = 0;     // the one-lane-per-spot kernel (sng_lanes.cuh): 0 auto (small batches, below), 1 whenever
                              // the station has an instantiation, -1 never
    // measured on a B200 (scripts/lanes_sweep.py, us per step, this kernel vs one block per warp): per-step launches
    // 4,096 envs 3.0 vs 3.6, 8,192 envs 4.2 vs 3.7; 24 steps per launch 4,096 envs 1.37 vs 3.34, 6,144 envs 2.06 vs 3.41,
    // 8,192 envs 2.31 vs 3.35 (two lanes per env: 2.14), 16,384 envs 4.6 vs 3.6
    static constexpr long long kLanesMaxEnvsStep = 4096, kLanesMaxEnvsRollout = 6144;
    static constexpr long long kTwoLanesMaxEnvsStep = 16384, kTwoLanesMaxEnvsRollout = 65536;   // 10 spots, two lanes per env (launch_step_n)
    int lanes_per_env = 0;    // 0 auto (4 for specialised stations of more than 32 spots), 1 / 2: that many lanes per env
    int num_sms = 148;
    size_t smem_optin = 0;
    void *d_tables = nullptr;
    uint32_t *d_flag = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;      // sng_step_host pipeline
    std::vector<cudaEvent_t> ev_in, ev_step;

    int init(const sng_config &c, int dev)
    {
        cfg = c;
        device = dev;
        if (c.n_spots < 1 || c.n_spots > SNG_MAX_SPOTS || c.n_steps < 1 || c.n_envs < 1 || c.n_envs > (1ll << 30) ||
            c.n_steps + (int)(4.0 / c.dt) > 250 ||
            c.table_len < c.n_steps + c.horizon || c.table_len > SNG_MAX_TABLE || !c.price || !c.price_norm ||
            (c.pv && (!c.pv_power || !c.irr_norm)) || c.penalty_mode < 0 || c.penalty_mode > 3 || c.horizon < 0) {
            error = "sng_create: invalid configuration";
            return SNG_ERR_ARG;
        }
        if (c.pv_days < 0 || c.pv_days > 255 ||
            (c.pv_days > 1 && c.table_len < (c.pv_days - 1) * c.n_steps + c.n_steps + c.horizon)) {
            error = "sng_create: pv_days needs PV tables of at least (pv_days - 1) * n_steps + n_steps + horizon entries";
            return SNG_ERR_ARG;
        }
        if (EXACT && c.n_spots > 128) {
            error = "sng_create: the float64 validation build supports at most 128 spots";
            return SNG_ERR_UNSUPPORTED;
        }
        DeviceGuard guard(dev);
        {
            int v = 0;
            SNG_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
            num_sms = v;
            SNG_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
            smem_optin = (size_t)v;
        }
        memset(&p, 0, sizeof(p));
        p.n_envs = c.n_envs;
        p.plane = (c.n_envs + kBlock - 1) / kBlock * (long long)c.n_spots * kBlock;
        p.gid0 = (unsigned long long)c.env_gid0;
        p.N = c.n_spots; p.T = c.n_steps; p.H = c.horizon;
        p.pv = c.pv != 0; p.batt = c.batt != 0; p.v2x = c.v2x != 0;
        p.A = p.N + p.batt;
        const int nd = (1 + p.pv) * (1 + p.H);
        p.off_soc = nd; p.off_dep = nd + p.N; p.off_batt = nd + 2 * p.N;
        p.D = nd + 2 * p.N + p.batt;
        p.pen_mode = c.penalty_mode; p.diff_cap = c.diff_cap != 0; p.req_soc = c.req_soc != 0;
        p.max_togo = c.penalty_mode == PEN_DENSE ? (1 << 20) : (c.penalty_mode == PEN_SPARSE ? 3 : (c.penalty_mode == PEN_ON_DEPARTURE ? 1 : 0));
        p.default_cap = c.default_cap; p.auto_reset = c.auto_reset != 0; p.mode = MODE_SAMPLE;
        p.pv_days = (c.pv && c.pv_days > 1) ? c.pv_days : 1;
        p.i4 = (int)(4.0 / c.dt); p.i10 = (int)(10.0 / c.dt); p.i1 = (int)(1.0 / c.dt);
        p.dt = (real)c.dt; p.ev_pmax = (real)c.ev_pmax; p.ev_eff = (real)c.ev_eff;
        p.b_cap = (real)c.b_cap; p.b_pmax = (real)c.b_pmax; p.b_eff = (real)c.b_eff; p.b_dod = (real)c.b_dod;
        p.b_soc0 = (real)c.b_soc0; p.sell = (real)c.sell_coeff; p.cost_w = (real)c.cost_weight;
        p.batt_w = (real)c.batt_pen_w; p.margin = (real)c.margin;
        p.dt_cap = c.b_cap > 0 ? (real)(c.dt / c.b_cap) : (real)0; p.cap_dt = (real)(c.b_cap / c.dt);
        if (c.default_cap < 1 || c.default_cap > 255) { error = "sng_create: default_cap must be in 1..255"; return SNG_ERR_ARG; }
        // shared tables: 4 x table_len reals + the departure-normalisation table
        const int n = c.table_len;
        std::vector<real> h(4 * (size_t)n, (real)0);
        for (int k = 0; k < n; ++k) {
            h[k] = c.pv ? (real)c.pv_power[k] : (real)0;
            h[n + k] = c.pv ? (real)c.irr_norm[k] : (real)0;
            h[2 * n + k] = (real)c.price[k];
            h[3 * n + k] = (real)c.price_norm[k];
        }
        std::vector<float> dn(kSmemTab, 0.0f);
        for (int k = 0; k < kDepTab; ++k) dn[k] = (float)((double)k / c.dep_norm);  // ...environment.py:208,228
        {   // thresholds of the geometric arrival gap (geometric_gap's recurrence), th_0 unused
            uint32_t th = 0x99999999u;
            uint32_t bits = 0xFFFFFFFFu;
            memcpy(&dn[kDepTab], &bits, 4);
            for (int k = 1; k < kGapTab; ++k) {
                memcpy(&dn[kDepTab + k], &th, 4);
                th = (uint32_t)(((unsigned long long)th * 0x9999999Aull) >> 32);
            }
        }
        const size_t tb = h.size() * sizeof(real), db = dn.size() * sizeof(float);
        SNG_CUDA(cudaMalloc(&d_tables, tb + db));
        SNG_CUDA(cudaMemcpy(d_tables, h.data(), tb, cudaMemcpyHostToDevice));
        SNG_CUDA(cudaMemcpy((char *)d_tables + tb, dn.data(), db, cudaMemcpyHostToDevice));
        const real *t0 = (const real *)d_tables;
        p.pv_power = t0; p.irr_norm = t0 + n; p.price = t0 + 2 * n; p.price_norm = t0 + 3 * n;
        p.dep_norm = (const float *)((char *)d_tables + tb);
        SNG_CUDA(cudaMalloc((void **)&d_flag, sizeof(uint32_t)));
        return SNG_OK;
    }

    ~Engine() override
    {
        DeviceGuard guard(device);
        if (d_tables) cudaFree(d_tables);
        if (d_flag) cudaFree(d_flag);
        for (auto e : ev_in) cudaEventDestroy(e);
        for (auto e : ev_step) cudaEventDestroy(e);
        if (copy_in) cudaStreamDestroy(copy_in);
        if (copy_out) cudaStreamDestroy(copy_out);
    }

    int bind(const sng_buffers *b) override
    {
        if (!b || b->struct_size != sizeof(sng_buffers)) { error = "sng_bind: bad struct_size"; return SNG_ERR_ARG; }
        if (!b->actions || !b->obs || !b->reward || !b->done || !b->spot || !b->envst) {
            error = "sng_bind: actions, obs, reward, done, spot and envst are required";
            return SNG_ERR_ARG;
        }
        buf = *b;
        p.actions = (const real *)b->actions; p.obs = b->obs; p.reward = (real *)b->reward; p.done = b->done;
        p.tobs = b->terminal_obs; p.spot = (typename WordOf<real>::type *)b->spot;
        p.envst = (EnvSt<real> *)b->envst; p.plan = (const PlanRec<real> *)b->plan; p.err = b->err;
        p.diag = (real *)b->diag; p.last_ret = (real *)b->last_return; p.spot_power = (real *)b->spot_power;
        bound = true;
        return SNG_OK;
    }

    static unsigned grid_for(long long threads) { return (unsigned)((threads + 255) / 256); }

    int check_ready(bool need_started)
    {
        if (!bound) { error = "buffers not bound (call sng_bind)"; return SNG_ERR_STATE; }
        if (need_started && !started) { error = "environment not reset (call sng_reset or sng_load_schedule)"; return SNG_ERR_STATE; }
        return SNG_OK;
    }

    int launch_reset(const uint8_t *mask, int init, int new_episode, int reset_battery, cudaStream_t st)
    {
        reset_kernel<real><<<grid_for(p.n_envs), 256, 0, st>>>(p, mask, init, new_episode, reset_battery);
        ++launches;
        SNG_CUDA(cudaGetLastError());
        return SNG_OK;
    }

    int reset(uint64_t seed, const uint8_t *mask, int reset_battery, cudaStream_t st) override
    {
        int rc = check_ready(false);
        if (rc) return rc;
        DeviceGuard guard(device);
        const bool first = !started;
        if (first && mask) { error = "sng_reset: the first reset must cover all envs (mask = NULL)"; return SNG_ERR_STATE; }
        if (mask) {
            // seed, sampling / replay mode and the requested-SoC plane rule belong to the whole handle: a masked
            // reset must not change them under the envs it leaves alone
            if (p.mode != MODE_SAMPLE) { error = "sng_reset: a masked reset needs sampling mode (the handle replays a loaded schedule)"; return SNG_ERR_STATE; }
            if (p.seed_lo != (uint32_t)seed || p.seed_hi != (uint32_t)(seed >> 32)) {
                error = "sng_reset: a masked reset cannot change the seed";
                return SNG_ERR_STATE;
            }
        } else {
            p.seed_lo = (uint32_t)seed;
            p.seed_hi = (uint32_t)(seed >> 32);
            p.mode = MODE_SAMPLE;
            p.has_req = p.req_soc;
        }
        rc = launch_reset(mask, first ? 1 : 0, 1, reset_battery, st);
        if (rc == SNG_OK) started = true;
        return rc;
    }

    int load_schedule(const sng_schedule_view *v, cudaStream_t st) override
    {
        int rc = check_ready(false);
        if (rc) return rc;
        if (!v || v->struct_size != sizeof(sng_schedule_view) || !v->arr || !v->dep || !v->cap || !v->soc0 ||
            !v->req || !v->n_veh || v->n_slots < 1 || v->n_slots > SNG_MAX_VEHICLES) {
            error = "sng_load_schedule: bad view";
            return SNG_ERR_ARG;
        }
        if (!buf.plan) { error = "sng_load_schedule: no `plan` buffer bound"; return SNG_ERR_STATE; }
        DeviceGuard guard(device);
        const long long E = p.n_envs;
        const int N = p.N, V = v->n_slots;
        std::vector<PlanRec<real>> plan((size_t)E * N * kMaxVehicles);
        for (long long e = 0; e < E; ++e)
            for (int i = 0; i < N; ++i) {
                const size_t sp = (size_t)e * N + i;
                const int nv = v->n_veh[sp];
                if (nv < 0 || nv > V) { error = "sng_load_schedule: n_veh out of range"; return SNG_ERR_ARG; }
                int prev_dep = -1;
                for (int k = 0; k < kMaxVehicles; ++k) {
                    PlanRec<real> r;
                    memset(&r, 0, sizeof(r));
                    r.hdr = make_hdr(kNoVehicle, 0, 0, kNoVehicle);
                    if (k < nv) {
                        const size_t q = sp * V + k;
                        const int a = v->arr[q], d = v->dep[q], c = v->cap[q];
                        // invariants of generated schedules the step kernel relies on (schedule.py validate())
                        if (a < 0 || a >= p.T || d <= a || d > 250 || d - a >= kDepTab || c < 1 || c > 255 || a <= prev_dep) {   // dep - t indexes the departure table
                            error = "sng_load_schedule: invalid vehicle record (arrival/departure/capacity)";
                            return SNG_ERR_ARG;
                        }
                        prev_dep = d;
                        const uint32_t nxt = (k + 1 < nv) ? (uint32_t)v->arr[q + 1] : kNoVehicle;
                        r.hdr = make_hdr((uint32_t)a, (uint32_t)d, (uint32_t)c, nxt);
                        r.soc0 = (real)v->soc0[q];
                        r.req = (real)v->req[q];
                    }
                    plan[sp * kMaxVehicles + k] = r;
                }
            }
        SNG_CUDA(cudaMemcpyAsync(buf.plan, plan.data(), plan.size() * sizeof(PlanRec<real>), cudaMemcpyHostToDevice, st));
        if (!started || v->pv_shift || v->soc_b) {
            std::vector<EnvSt<real>> es((size_t)E);
            if (started) {
                SNG_CUDA(cudaMemcpyAsync(es.data(), buf.envst, es.size() * sizeof(EnvSt<real>), cudaMemcpyDeviceToHost, st));
                SNG_CUDA(cudaStreamSynchronize(st));
            } else {
                memset(es.data(), 0, es.size() * sizeof(EnvSt<real>));
                for (auto &x : es) { x.soc_b = p.batt ? p.b_soc0 : (real)0; x.pv_shift = 1; }
            }
            for (long long e = 0; e < E; ++e) {
                if (v->pv_shift) es[e].pv_shift = (real)v->pv_shift[e];
                if (v->soc_b) es[e].soc_b = (real)v->soc_b[e];
            }
            SNG_CUDA(cudaMemcpyAsync(buf.envst, es.data(), es.size() * sizeof(EnvSt<real>), cudaMemcpyHostToDevice, st));
            SNG_CUDA(cudaStreamSynchronize(st));
        }
        if (!started && buf.err) SNG_CUDA(cudaMemsetAsync(buf.err, 0, sizeof(uint32_t) * E, st));
        p.mode = MODE_REPLAY;
        p.has_req = 1;
        rc = launch_reset(nullptr, 0, 0, 0, st);
        SNG_CUDA(cudaStreamSynchronize(st));  // `plan` staging vector goes out of scope
        if (rc == SNG_OK) started = true;
        return rc;
    }

    static bool aligned16(const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

    // The dynamic shared-memory limit is a per-function, per-device attribute shared by every handle:
    // raise it monotonically, once per (device, kernel), instead of setting it on every launch.
    int ensure_smem(const void *kern, size_t smem)
    {
        static std::mutex mu;
        static std::map<std::pair<int, const void *>, size_t> limit;
        std::lock_guard<std::mutex> lock(mu);
        size_t &cur = limit[std::make_pair(device, kern)];
        if (smem <= cur) return SNG_OK;
        SNG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
        return SNG_OK;
    }

    size_t row_bytes(int stages) const
    {
        return (size_t)stages * align128((uint32_t)(kBlock * p.A * sizeof(real))) + align128((uint32_t)(kBlock * p.D * sizeof(float)));
    }
    static constexpr size_t kStaticSmem = kSmemTab * sizeof(float);

    // L lanes per env (see step_simple_kernel); L = 2 requires q.n_envs to be a multiple of 32.
    template <int NCT, int ND, int L = 1, bool FIXED = false>
    int launch_simple(const Params<real> &q, const real *actions, float *obs, real *reward, uint8_t *done, int n_steps,
                      int bulk, cudaStream_t st)
    {
        constexpr int EPW = kBlock / L;
        const size_t per_warp = (n_steps > 1 ? 2 : 1) * align128((uint32_t)(EPW * p.A * sizeof(real))) + align128((uint32_t)(EPW * p.D * sizeof(float)));
        int wpb = warps_per_cta > 0 ? warps_per_cta : 2;
        if (wpb * 32 > SNG_STEP_MAXT) wpb = SNG_STEP_MAXT / 32;
        while (wpb > 1 && kStaticSmem + 128 + (size_t)wpb * per_warp > smem_optin) wpb >>= 1;
        const size_t smem = 128 + (size_t)wpb * per_warp;   // mbarrier header + the warps' row stages
        if (kStaticSmem + smem > smem_optin) { error = "step kernel: one warp's action/observation rows do not fit in shared memory"; return SNG_ERR_UNSUPPORTED; }
        auto kern = n_steps > 1 ? step_simple_kernel<real, NCT, ND, EXACT, true, L, FIXED> : step_simple_kernel<real, NCT, ND, EXACT, false, L, FIXED>;
        int rc = ensure_smem((const void *)kern, smem);
        if (rc) return rc;
        const long long groups = (q.n_envs + EPW - 1) / EPW;       // one warp each
        const unsigned grid = (unsigned)((groups + wpb - 1) / wpb);
        if (launch_mode > 0 && n_steps == 1) {
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(grid); lc.blockDim = dim3(wpb * 32); lc.dynamicSmemBytes = smem; lc.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            lc.attrs = at; lc.numAttrs = 1;
            const int mode = bulk | (launch_mode == 2 ? STAGE_PDL_EARLY : 0);
            SNG_CUDA(cudaLaunchKernelEx(&lc, kern, q, actions, obs, reward, done, n_steps, mode));
        } else {
            kern<<<grid, wpb * 32, smem, st>>>(q, actions, obs, reward, done, n_steps, bulk);
        }
        ++launches;
        SNG_CUDA(cudaGetLastError());
        return SNG_OK;
    }

    // One lane per charging spot, two envs per warp (sng_lanes.cuh): float32, the reference's default station shape.
    template <int NCT>
    int launch_lanes(const Params<real> &q, const real *actions, float *obs, real *reward, uint8_t *done, int n_steps, cudaStream_t st)
    {
        if constexpr (!EXACT && NCT <= kLaneGroup) {
            auto kern = q.has_req ? (n_steps > 1 ? step_lanes_kernel<NCT, true, true> : step_lanes_kernel<NCT, false, true>)
                                  : (n_steps > 1 ? step_lanes_kernel<NCT, true, false> : step_lanes_kernel<NCT, false, false>);
            const int wpb = (warps_per_cta > 0 && warps_per_cta < 4) ? warps_per_cta : 4;
            const long long warps = (q.n_envs + 1) / 2;
            const unsigned grid = (unsigned)((warps + wpb - 1) / wpb);
            if (launch_mode > 0 && n_steps == 1) {
                cudaLaunchConfig_t lc = {};
                lc.gridDim = dim3(grid); lc.blockDim = dim3(wpb * 32); lc.dynamicSmemBytes = 0; lc.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                lc.attrs = at; lc.numAttrs = 1;
                SNG_CUDA(cudaLaunchKernelEx(&lc, kern, q, actions, obs, reward, done, n_steps));
            } else {
                kern<<<grid, wpb * 32, 0, st>>>(q, actions, obs, reward, done, n_steps);
            }
            ++launches;
            SNG_CUDA(cudaGetLastError());
            return SNG_OK;
        } else {
            return SNG_ERR_UNSUPPORTED;
        }
    }

    // Persistent pipelined kernel over the full 32-env blocks of q; returns SNG_ERR_UNSUPPORTED (without
    // launching) when a warp's stages do not fit in shared memory.
    template <int NCT, int ND> int launch_pipelined(const Params<real> &q, long long n_blocks, cudaStream_t st)
    {
        const size_t per_warp = row_bytes(2) + 16;
        int wpb = warps_per_cta > 0 ? warps_per_cta : 2;
        if (wpb * 32 > SNG_PIPE_THREADS) wpb = SNG_PIPE_THREADS / 32;
        const size_t smem = (size_t)wpb * per_warp;
        if (kStaticSmem + smem > smem_optin) return SNG_ERR_UNSUPPORTED;
        auto kern = step_pipelined_kernel<real, NCT, ND, EXACT>;
        int rc0 = ensure_smem((const void *)kern, smem);
        if (rc0) return rc0;
        int per_sm = 0;
        SNG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpb * 32, smem));
        if (per_sm < 1) return SNG_ERR_UNSUPPORTED;
        if (ctas_per_sm > 0 && per_sm > ctas_per_sm) per_sm = ctas_per_sm;
        long long grid = (long long)num_sms * per_sm;
        const long long need = (n_blocks + wpb - 1) / wpb;
        if (grid > need) grid = need;
        kern<<<(unsigned)grid, wpb * 32, smem, st>>>(q, n_blocks);
        ++launches;
        SNG_CUDA(cudaGetLastError());
        return SNG_OK;
    }

    template <int NCT, int ND, bool FIXED = false>
    int launch_step_n(const Params<real> &q, const real *actions, float *obs, real *reward, uint8_t *done, int n_steps,
                      int bulk, cudaStream_t st)
    {
        if ((bulk & STAGE_ALIGNED) && kernel_variant == 1 && n_steps == 1 && q.n_envs >= kBlock && actions == q.actions && obs == q.obs &&
            reward == q.reward && done == q.done) {
            const long long n_blocks = q.n_envs / kBlock;
            const int rc = launch_pipelined<NCT, ND>(q, n_blocks, st);
            if (rc == SNG_OK) {
                const long long e0 = n_blocks * kBlock;
                if (e0 == q.n_envs) return SNG_OK;
                const Params<real> tail = slice_of(q, e0, q.n_envs - e0);   // ragged last block
                return launch_simple<NCT, ND, 1, FIXED>(tail, tail.actions, tail.obs, tail.reward, tail.done, 1, 0, st);
            }
            if (rc != SNG_ERR_UNSUPPORTED) return rc;
        }
        if constexpr (!EXACT && NCT > 32 && NCT % 2 == 0) {
            // large stations (64 spots): four (or two) lanes per env over the whole 32-env blocks, the ragged last block one lane
            // per env (a 32-spot station already reaches 85 % of the HBM roofline with one lane per env)
            if (lanes_per_env != 1 && q.n_envs >= kBlock) {
                // rows of 65 / 137 floats: the observation rows leave through the copy engine as well (measured on C5:
                // 0.0828 vs 0.0877 ms per step; at 10 spots the two are equal)
                if (use_bulk == 1 && (bulk & STAGE_ALIGNED)) bulk |= STAGE_TMA_STORE;
                const long long full = q.n_envs / kBlock * kBlock;
                const bool four = lanes_per_env != 2 && NCT % 4 == 0;      // default: four lanes per env (a warp covers 8 envs)
                if (full == q.n_envs)
                    return four ? launch_simple<NCT, ND, 4, FIXED>(q, actions, obs, reward, done, n_steps, bulk, st)
                                : launch_simple<NCT, ND, 2, FIXED>(q, actions, obs, reward, done, n_steps, bulk, st);
                if (n_steps == 1 && actions == q.actions && obs == q.obs && reward == q.reward && done == q.done) {
                    const Params<real> head = slice_of(q, 0, full), tail = slice_of(q, full, q.n_envs - full);
                    const int rc = four ? launch_simple<NCT, ND, 4, FIXED>(head, head.actions, head.obs, head.reward, head.done, 1, bulk, st)
                                        : launch_simple<NCT, ND, 2, FIXED>(head, head.actions, head.obs, head.reward, head.done, 1, bulk, st);
                    if (rc) return rc;
                    return launch_simple<NCT, ND, 1, FIXED>(tail, tail.actions, tail.obs, tail.reward, tail.done, 1, STAGE_SCALAR, st);
                }
            }
        }
        if constexpr (!EXACT && FIXED && NCT == 10) {
            // two lanes per env at 10 spots: half the per-warp chain for batches that are a single wave of warps anyway
            // (whole 32-env blocks; bit-identical).  Measured (scripts/lanes_sweep.py --two, us per step, two lanes vs one):
            // 24 steps per launch 8,192 envs 2.14 vs 3.37, 16,384 envs 2.39 vs 3.58, 32,768 envs 3.25 vs 4.47, 65,536 envs 5.13
            // vs 5.31, 131,072 envs 9.3 vs 7.6; per-step launches 8,192 envs 3.34 vs 3.59, 16,384 envs 3.61 vs 3.75, 32,768 equal
            const bool mid = lanes_per_env == 0 && kernel_variant == 0 && lanes_kernel == 0 &&
                             q.n_envs <= (n_steps > 1 ? kTwoLanesMaxEnvsRollout : kTwoLanesMaxEnvsStep);
            if ((lanes_per_env == 2 || mid) && q.n_envs % kBlock == 0)
                return launch_simple<NCT, ND, 2, FIXED>(q, actions, obs, reward, done, n_steps, bulk, st);
        }
        return launch_simple<NCT, ND, 1, FIXED>(q, actions, obs, reward, done, n_steps, bulk, st);
    }

    // q: parameters (possibly of a slice of envs starting at a multiple of 32)
    int launch_step(const Params<real> &q, const real *actions, float *obs, real *reward, uint8_t *done, int n_steps,
                    cudaStream_t st)
    {
        // the copy engine and the 16-byte vector path need aligned row slabs: bases aligned and, for rollouts,
        // slab strides too; otherwise every block is staged with scalar accesses
        bool aligned = aligned16(actions) && aligned16(obs);
        if (n_steps > 1 && (((size_t)q.n_envs * q.A * sizeof(real)) % 16 != 0 || ((size_t)q.n_envs * q.D * sizeof(float)) % 16 != 0))
            aligned = false;
        const int bulk = (aligned && use_bulk >= 0) ? (STAGE_ALIGNED | (use_bulk & 3)) : STAGE_SCALAR;
        if constexpr (EXACT) {
            return launch_step_n<0, 0>(q, actions, obs, reward, done, n_steps, bulk, st);
        } else {
            // specialised kernels: the station sizes BASELINE.json names, with the reference's observation
            // shape (PV on, 3 steps ahead: 8 disturbance entries); everything else runs the generic kernel
            // (keyed on the flags themselves: PV off with 7 steps ahead also has 8 disturbance entries, but a
            // different layout -- eight prices)
            // latency-bound batches of that shape with a battery (sampled or replayed, with or without requested SoCs): one lane
            // per spot (bit-identical; 4,096 envs are only 128 warps of the kernels below)
            if (!use_generic && q.pv && q.H == 3 && q.pv_days == 1 && q.batt && kernel_variant == 0 &&
                (lanes_kernel > 0 || (lanes_kernel == 0 && q.n_envs <= (n_steps > 1 ? kLanesMaxEnvsRollout : kLanesMaxEnvsStep)))) {
                switch (q.N) {
                case 4: return launch_lanes<4>(q, actions, obs, reward, done, n_steps, st);
                case 8: return launch_lanes<8>(q, actions, obs, reward, done, n_steps, st);
                case 10: return launch_lanes<10>(q, actions, obs, reward, done, n_steps, st);
                default: break;
                }
            }
            if (!use_generic && q.pv && q.H == 3 && q.pv_days == 1 && q.batt && !q.has_req) {
                // ... and the rest of the reference's default station (battery, no requested-SoC plane) at compile time too
                switch (q.N) {
                case 4: return launch_step_n<4, 8, true>(q, actions, obs, reward, done, n_steps, bulk, st);
                case 8: return launch_step_n<8, 8, true>(q, actions, obs, reward, done, n_steps, bulk, st);
                case 10: return launch_step_n<10, 8, true>(q, actions, obs, reward, done, n_steps, bulk, st);
                case 32: return launch_step_n<32, 8, true>(q, actions, obs, reward, done, n_steps, bulk, st);
                case 64: return launch_step_n<64, 8, true>(q, actions, obs, reward, done, n_steps, bulk, st);
                default: break;
                }
            }
            if (!use_generic && q.pv && q.H == 3 && q.pv_days == 1) {
                switch (q.N) {
                case 4: return launch_step_n<4, 8>(q, actions, obs, reward, done, n_steps, bulk, st);
                case 8: return launch_step_n<8, 8>(q, actions, obs, reward, done, n_steps, bulk, st);
                case 10: return launch_step_n<10, 8>(q, actions, obs, reward, done, n_steps, bulk, st);
                case 32: return launch_step_n<32, 8>(q, actions, obs, reward, done, n_steps, bulk, st);
                case 64: return launch_step_n<64, 8>(q, actions, obs, reward, done, n_steps, bulk, st);
                default: break;
                }
            }
            // the same station sizes with any other observation shape (forecast horizon != 3, no PV, multi-day PV):
            // compile-time spot count, observation offsets read from the parameters
            if (!use_generic) {
                switch (q.N) {
                case 4: return launch_step_n<4, 0>(q, actions, obs, reward, done, n_steps, bulk, st);
                case 8: return launch_step_n<8, 0>(q, actions, obs, reward, done, n_steps, bulk, st);
                case 10: return launch_step_n<10, 0>(q, actions, obs, reward, done, n_steps, bulk, st);
                case 32: return launch_step_n<32, 0>(q, actions, obs, reward, done, n_steps, bulk, st);
                case 64: return launch_step_n<64, 0>(q, actions, obs, reward, done, n_steps, bulk, st);
                default: break;
                }
            }
            return launch_step_n<0, 0>(q, actions, obs, reward, done, n_steps, bulk, st);
        }
    }

    int step(cudaStream_t st) override
    {
        int rc = check_ready(true);
        if (rc) return rc;
        DeviceGuard guard(device);
        return launch_step(p, p.actions, p.obs, p.reward, p.done, 1, st);
    }

    int rollout(const void *actions, float *obs, void *reward, uint8_t *done, int n_steps, cudaStream_t st) override
    {
        int rc = check_ready(true);
        if (rc) return rc;
        if (!actions || !obs || !reward || !done || n_steps < 1) { error = "sng_rollout: bad arguments"; return SNG_ERR_ARG; }
        DeviceGuard guard(device);
        return launch_step(p, (const real *)actions, obs, (real *)reward, done, n_steps, st);
    }

    // Parameters with every per-env pointer advanced by e0 envs (e0 a multiple of 32).
    static Params<real> slice_of(const Params<real> &src, long long e0, long long n)
    {
        Params<real> q = src;
        q.n_envs = n;
        q.gid0 = src.gid0 + (unsigned long long)e0;
        q.actions += (size_t)e0 * src.A; q.obs += (size_t)e0 * src.D; q.reward += e0; q.done += e0;
        if (q.tobs) q.tobs += (size_t)e0 * src.D;
        q.spot += (size_t)e0 * src.N; q.envst += e0;        // plane-major: the slice starts e0 / 32 blocks into every plane
        if (q.plan) q.plan += (size_t)e0 * src.N * kMaxVehicles;
        if (q.err) q.err += e0;
        if (q.diag) q.diag += (size_t)e0 * D_COUNT;
        if (q.last_ret) q.last_ret += e0;
        if (q.spot_power) q.spot_power += (size_t)e0 * src.N;
        return q;
    }
    Params<real> slice_params(long long e0, long long n) const { return slice_of(p, e0, n); }

    // The gym-facing call with host buffers.  The batch is cut into chunks of envs that flow through
    // a three-stage pipeline on three streams (H2D actions | step kernel | D2H obs/reward/done), so the
    // two PCIe directions and the kernel overlap; with one chunk it degenerates to copy-step-copy.
    int step_host(const void *a, float *obs, void *rew, uint8_t *done, cudaStream_t st) override
    {
        int rc = check_ready(true);
        if (rc) return rc;
        if (!a || !obs || !rew || !done) { error = "sng_step_host: null host buffer"; return SNG_ERR_ARG; }
        DeviceGuard guard(device);
        struct ModeGuard {      // the chunk kernels follow each other and copies: ordinary launches
            int &m; int saved;
            explicit ModeGuard(int &r) : m(r), saved(r) { m = 0; }
            ~ModeGuard() { m = saved; }
        } mode_guard(launch_mode);
        const long long E = p.n_envs;
        int chunks = host_chunks > 0 ? host_chunks : (E >= (1 << 16) ? 8 : 1);
        long long per = ((E + chunks - 1) / chunks + 1023) / 1024 * 1024;      // multiple of 32 (and of the TMA slab alignment)
        if (per >= E) { chunks = 1; per = E; }
        chunks = (int)((E + per - 1) / per);
        if (chunks == 1) {
            SNG_CUDA(cudaMemcpyAsync((void *)buf.actions, a, (size_t)E * p.A * sizeof(real), cudaMemcpyHostToDevice, st));
            rc = step(st);
            if (rc) return rc;
            SNG_CUDA(cudaMemcpyAsync(obs, buf.obs, (size_t)E * p.D * sizeof(float), cudaMemcpyDeviceToHost, st));
            SNG_CUDA(cudaMemcpyAsync(rew, buf.reward, (size_t)E * sizeof(real), cudaMemcpyDeviceToHost, st));
            SNG_CUDA(cudaMemcpyAsync(done, buf.done, (size_t)E, cudaMemcpyDeviceToHost, st));
            SNG_CUDA(cudaStreamSynchronize(st));
            return SNG_OK;
        }
        if (!copy_in) {
            SNG_CUDA(cudaStreamCreateWithFlags(&copy_in, cudaStreamNonBlocking));
            SNG_CUDA(cudaStreamCreateWithFlags(&copy_out, cudaStreamNonBlocking));
        }
        while ((int)ev_in.size() < chunks) {
            cudaEvent_t e1, e2;
            SNG_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            SNG_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            ev_in.push_back(e1);
            ev_step.push_back(e2);
        }
        // the copy streams start after whatever the caller queued on `st`
        cudaEvent_t ev0 = ev_in[0];
        SNG_CUDA(cudaEventRecord(ev0, st));
        SNG_CUDA(cudaStreamWaitEvent(copy_in, ev0, 0));
        SNG_CUDA(cudaStreamWaitEvent(copy_out, ev0, 0));
        // chunk boundaries: equal chunks, except that the first one is cut into 1/4 + 1/4 + 1/2 -- the D2H engine (the
        // bottleneck) idles until the first chunk's actions have arrived and its step has run, so that chunk is small
        std::vector<long long> cut;
        for (long long e0 = 0; e0 < E; e0 += per) {
            if (e0 == 0 && per % 4096 == 0) { cut.push_back(0); cut.push_back(per / 4); cut.push_back(per / 2); }
            else cut.push_back(e0);
        }
        cut.push_back(E);
        chunks = (int)cut.size() - 1;
        while ((int)ev_in.size() < chunks) {
            cudaEvent_t e1, e2;
            SNG_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            SNG_CUDA(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            ev_in.push_back(e1);
            ev_step.push_back(e2);
        }
        for (int c = 0; c < chunks; ++c) {
            const long long e0 = cut[c], n = cut[c + 1] - cut[c];
            SNG_CUDA(cudaMemcpyAsync((char *)buf.actions + (size_t)e0 * p.A * sizeof(real), (const char *)a + (size_t)e0 * p.A * sizeof(real),
                                     (size_t)n * p.A * sizeof(real), cudaMemcpyHostToDevice, copy_in));
            SNG_CUDA(cudaEventRecord(ev_in[c], copy_in));
            SNG_CUDA(cudaStreamWaitEvent(st, ev_in[c], 0));
            const Params<real> q = slice_params(e0, n);
            rc = launch_step(q, q.actions, q.obs, q.reward, q.done, 1, st);
            if (rc) return rc;
            SNG_CUDA(cudaEventRecord(ev_step[c], st));
            SNG_CUDA(cudaStreamWaitEvent(copy_out, ev_step[c], 0));
            SNG_CUDA(cudaMemcpyAsync(obs + (size_t)e0 * p.D, buf.obs + (size_t)e0 * p.D, (size_t)n * p.D * sizeof(float), cudaMemcpyDeviceToHost, copy_out));
        }
        // rewards and done flags are 4 % of the bytes: one copy each behind the last chunk's rows instead of two small
        // copies per chunk (every copy costs ~10 us of fixed time on the D2H engine, which is the bottleneck)
        SNG_CUDA(cudaMemcpyAsync(rew, buf.reward, (size_t)E * sizeof(real), cudaMemcpyDeviceToHost, copy_out));
        SNG_CUDA(cudaMemcpyAsync(done, buf.done, (size_t)E, cudaMemcpyDeviceToHost, copy_out));
        SNG_CUDA(cudaStreamSynchronize(copy_out));
        SNG_CUDA(cudaStreamSynchronize(st));
        return SNG_OK;
    }

    int sample_plan(cudaStream_t st) override
    {
        int rc = check_ready(true);
        if (rc) return rc;
        if (!buf.plan) { error = "sng_sample_plan: no `plan` buffer bound"; return SNG_ERR_STATE; }
        if (p.mode != MODE_SAMPLE) { error = "sng_sample_plan: handle is in replay mode"; return SNG_ERR_STATE; }
        DeviceGuard guard(device);
        sample_plan_kernel<real><<<grid_for(p.n_envs * p.N), 256, 0, st>>>(p, (PlanRec<real> *)buf.plan);
        ++launches;
        SNG_CUDA(cudaGetLastError());
        return SNG_OK;
    }

    int sample_actions(uint64_t seed, uint64_t step0, int n_steps, void *actions, cudaStream_t st) override
    {
        if (!actions || n_steps < 1) { error = "sng_sample_actions: bad arguments"; return SNG_ERR_ARG; }
        if (step0 + (uint64_t)n_steps > 0xFFFFFFFFull) { error = "sng_sample_actions: step0 + n_steps must stay below 2^32"; return SNG_ERR_ARG; }
        DeviceGuard guard(device);
        const long long blocks = ((long long)n_steps * p.n_envs * ((p.A + 3) / 4) + 255) / 256;
        sample_actions_kernel<real><<<(unsigned)(blocks < 148 * 64 ? blocks : 148 * 64), 256, 0, st>>>(p, (real *)actions, seed, step0, n_steps);
        ++launches;
        SNG_CUDA(cudaGetLastError());
        return SNG_OK;
    }

    int probe_arrival_gap(const uint32_t *x, uint32_t *g, long long n, cudaStream_t st) override
    {
        if (!x || !g || n < 1) { error = "sng_debug_arrival_gap: bad arguments"; return SNG_ERR_ARG; }
        DeviceGuard guard(device);
        gap_probe_kernel<real><<<148, 128, 0, st>>>(p, x, g, n);
        ++launches;
        SNG_CUDA(cudaGetLastError());
        return SNG_OK;
    }

    int traffic_skeleton(int variant, cudaStream_t st) override
    {
        int rc = check_ready(true);
        if (rc) return rc;
        if constexpr (!EXACT) {
            const bool fixed = p.pv && p.H == 3 && p.pv_days == 1 && p.batt && !p.has_req && p.N == 10;
            if (variant < 0 || variant > 5) { error = "sng_debug_traffic_skeleton: variant must be in 0..5"; return SNG_ERR_ARG; }
            if (!fixed || p.n_envs % kBlock != 0 || !aligned16(p.actions) || !aligned16(p.obs)) {
                error = "sng_debug_traffic_skeleton: default 10-spot station, whole 32-env blocks, aligned buffers only";
                return SNG_ERR_UNSUPPORTED;
            }
            DeviceGuard guard(device);
            const int wpb = 2;
            const size_t smem = 128 + (size_t)wpb * row_bytes(1);
            auto kern = traffic_skeleton_kernel<real, 10>;
            rc = ensure_smem((const void *)kern, smem);
            if (rc) return rc;
            kern<<<(unsigned)((p.n_envs / kBlock + wpb - 1) / wpb), wpb * 32, smem, st>>>(p, variant);
            ++launches;
            SNG_CUDA(cudaGetLastError());
            return SNG_OK;
        } else {
            error = "sng_debug_traffic_skeleton: float32 build only";
            return SNG_ERR_UNSUPPORTED;
        }
    }

    int error_flags(uint32_t *out, cudaStream_t st) override
    {
        int rc = check_ready(false);
        if (rc) return rc;
        if (!out) { error = "sng_error_flags: null output"; return SNG_ERR_ARG; }
        *out = 0;
        if (!buf.err) return SNG_OK;
        DeviceGuard guard(device);
        SNG_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), st));
        or_reduce_kernel<<<148, 256, 0, st>>>(buf.err, p.n_envs, d_flag);
        ++launches;
        SNG_CUDA(cudaMemcpyAsync(out, d_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SNG_CUDA(cudaStreamSynchronize(st));
        return SNG_OK;
    }

    // The fused kernel exists for the reference's default station at 4 and 10 spots (what the policy of BASELINE config 3
    // drives); everything else answers SNG_ERR_UNSUPPORTED and the caller launches policy and step separately.
    int policy_step(const PolicyStepArgs &a, cudaStream_t st) override
    {
        int rc = check_ready(true);
        if (rc) return rc;
        if constexpr (!EXACT && std::is_same<real, float>::value) {
            const bool fixed = p.pv && p.H == 3 && p.pv_days == 1 && p.batt && !p.has_req && (p.N == 4 || p.N == 10);
            if (!fixed || use_generic || p.n_envs % 128 != 0) { error = "sng_policy_step: unsupported station shape or batch size"; return SNG_ERR_UNSUPPORTED; }
            if (!a.packed || !a.obs || !a.low || !a.high || !a.raw_actions || !a.actions || !a.values || !a.log_probs || !a.obs_next ||
                !a.reward || !a.done || (!a.noise && !a.step_counter)) { error = "sng_policy_step: null argument"; return SNG_ERR_ARG; }
            DeviceGuard guard(device);
            rc = launch_policy_step(p, a, st);
            if (rc == SNG_ERR_UNSUPPORTED) error = "sng_policy_step: buffers must be 16-byte aligned";
            else if (rc) error = std::string("sng_policy_step: ") + cudaGetErrorString(cudaGetLastError());
            else ++launches;
            return rc;
        } else {
            error = "sng_policy_step: float32 build only";
            return SNG_ERR_UNSUPPORTED;
        }
    }

    int set_tuning(int w, int g, int b, int hc) override
    {
        if (w != 0 && w != 1 && w != 2 && w != 4 && w != 8) {
            error = "sng_set_tuning: warps_per_cta must be 0, 1, 2, 4 or 8";
            return SNG_ERR_ARG;
        }
        if (hc < 0 || hc > 64) { error = "sng_set_tuning: host_chunks must be in 0..64"; return SNG_ERR_ARG; }
        warps_per_cta = w; use_generic = g; use_bulk = b; host_chunks = hc;
        return SNG_OK;
    }

    int set_launch_mode(int mode) override
    {
        if (mode < 0 || mode > 2) { error = "sng_set_launch_mode: mode must be 0, 1 or 2"; return SNG_ERR_ARG; }
        launch_mode = mode;
        return SNG_OK;
    }

    int set_pipeline(int up, int cps) override
    {
        if (cps < 0 || up < 0 || up > 5) { error = "sng_set_pipeline: bad arguments"; return SNG_ERR_ARG; }
        kernel_variant = up >= 2 ? 0 : up;
        lanes_per_env = up == 2 ? 1 : (up == 3 ? 2 : 0);
        lanes_kernel = up == 4 ? 1 : (up == 0 ? 0 : -1);      // 2, 3, 5: the one-block-per-warp kernel at every batch size
        ctas_per_sm = cps;
        return SNG_OK;
    }
};

}  // namespace sng
