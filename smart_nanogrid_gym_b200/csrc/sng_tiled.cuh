// sng_tiled.cuh -- the production step kernel for B200: persistent CTAs stream tiles of
// environments through shared memory with the TMA bulk-copy engine.
//
//   HBM --cp.async.bulk (global->shared, mbarrier complete_tx)--> input stage (actions, SoC,
//   vehicle records, env scalars)  --env_step() per env-->  output stage (obs, SoC, env scalars,
//   reward, done)  --cp.async.bulk (shared->global, bulk_group)--> HBM
//
// Every per-env array is dense and env-major, so a tile of EPB consecutive envs is ONE
// contiguous, 16-byte-aligned byte range per array: rows of 11 or 29 floats need no padding and
// every DRAM access is a full-line burst issued by the copy engine, not by the SM's LSUs.
// Threads then work out of shared memory with any lanes-per-env mapping (L = 1: one thread per
// env, spots in registers; L = 16/32: the warp-per-env mapping with shuffle reductions).
// The input ring (IN_STAGES deep) keeps IN_STAGES-1 tiles of loads in flight per CTA while one
// tile is computed; the output ring lets the stores of tile k drain while tile k+1 is computed.
#pragma once
#include "sng_device.cuh"

namespace sng {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// TMA 1-D bulk copies (SASS: UBLKCP).  Addresses and sizes must be multiples of 16 bytes.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Shared-memory carve-up for a tile of `epb` envs (all offsets multiples of 16).
struct TileLayout {
    uint32_t act, soc, rec, envst, in_bytes;       // input stage
    uint32_t obs, osoc, oenvst, rew, done, out_bytes;  // output stage
    uint32_t tab_real, tab_dep, tables_bytes;
    uint32_t bars, total;
};

template <typename real>
__host__ __device__ inline TileLayout make_tile_layout(int epb, int N, int A, int D, int table_len, int in_stages,
                                                       int out_stages)
{
    TileLayout t;
    uint32_t o = 0;
    t.act = o;   o += (uint32_t)align16((size_t)epb * A * sizeof(real));
    t.soc = o;   o += (uint32_t)align16((size_t)epb * N * sizeof(real));
    t.rec = o;   o += (uint32_t)align16((size_t)epb * N * sizeof(Rec<real>));
    t.envst = o; o += (uint32_t)align16((size_t)epb * sizeof(EnvSt<real>));
    t.in_bytes = o;
    o = 0;
    t.obs = o;    o += (uint32_t)align16((size_t)epb * D * sizeof(float));
    t.osoc = o;   o += (uint32_t)align16((size_t)epb * N * sizeof(real));
    t.oenvst = o; o += (uint32_t)align16((size_t)epb * sizeof(EnvSt<real>));
    t.rew = o;    o += (uint32_t)align16((size_t)epb * sizeof(real));
    t.done = o;   o += (uint32_t)align16((size_t)epb);
    t.out_bytes = o;
    t.tab_real = 0;
    t.tab_dep = (uint32_t)align16((size_t)4 * table_len * sizeof(real));
    t.tables_bytes = t.tab_dep + (uint32_t)align16((size_t)kDepTab * sizeof(float));
    t.bars = t.tables_bytes + in_stages * t.in_bytes + out_stages * t.out_bytes;
    t.total = t.bars + (uint32_t)align16((size_t)in_stages * sizeof(uint64_t));
    return t;
}

// Requires: every bound array 16-byte aligned, epb a multiple of 16, num_tiles * epb <= n_envs
// (the host runs the remainder through the direct kernel).
template <typename real, int L, int OUT_STAGES>
__global__ void __launch_bounds__(1024) step_tiled_kernel(const Params<real> p, int epb, int num_tiles, int in_stages,
                                                          int table_len)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const TileLayout lay = make_tile_layout<real>(epb, p.N, p.A, p.D, table_len, in_stages, OUT_STAGES);
    real *tab = reinterpret_cast<real *>(smem + lay.tab_real);
    float *tab_dep = reinterpret_cast<float *>(smem + lay.tab_dep);
    unsigned char *in_base = smem + lay.tables_bytes;
    unsigned char *out_base = in_base + (size_t)in_stages * lay.in_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + lay.bars);

    const int tid = threadIdx.x;
    const int env_local = tid / L;
    const int lane = tid % L;
    const uint32_t act_bytes = (uint32_t)(epb * p.A * sizeof(real));
    const uint32_t soc_bytes = (uint32_t)(epb * p.N * sizeof(real));
    const uint32_t rec_bytes = (uint32_t)(epb * p.N * sizeof(Rec<real>));
    const uint32_t est_bytes = (uint32_t)(epb * sizeof(EnvSt<real>));
    const uint32_t obs_bytes = (uint32_t)(epb * p.D * sizeof(float));
    const uint32_t rew_bytes = (uint32_t)(epb * sizeof(real));

    auto issue_loads = [&](int tile, int stage) {   // one elected thread
        unsigned char *st = in_base + (size_t)stage * lay.in_bytes;
        const size_t e0 = (size_t)tile * epb;
        mbar_expect_tx(&full[stage], act_bytes + soc_bytes + rec_bytes + est_bytes);
        bulk_g2s(st + lay.act, p.actions + e0 * p.A, act_bytes, &full[stage]);
        bulk_g2s(st + lay.soc, p.soc + e0 * p.N, soc_bytes, &full[stage]);
        bulk_g2s(st + lay.rec, p.rec + e0 * p.N, rec_bytes, &full[stage]);
        bulk_g2s(st + lay.envst, p.envst + e0, est_bytes, &full[stage]);
    };

    if (tid == 0) {
        for (int s = 0; s < in_stages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < in_stages; ++s) {
            const int tile = blockIdx.x + s * gridDim.x;
            if (tile < num_tiles) issue_loads(tile, s);
        }
    }
    // shared tables (2 x 48 .. 2 x 192 entries each): staged once per CTA
    for (int k = tid; k < 4 * table_len; k += blockDim.x) tab[k] = p.pv_power[k];   // the 4 tables are contiguous
    for (int k = tid; k < kDepTab; k += blockDim.x) tab_dep[k] = p.dep_norm[k];
    Tables<real> tb;
    tb.pv_power = tab; tb.irr_norm = tab + table_len; tb.price = tab + 2 * table_len; tb.price_norm = tab + 3 * table_len;
    tb.dep_norm = tab_dep;
    __syncthreads();

    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int s = it % in_stages;
        const uint32_t parity = (uint32_t)((it / in_stages) & 1);
        const int o = it % OUT_STAGES;
        unsigned char *st = in_base + (size_t)s * lay.in_bytes;
        unsigned char *ot = out_base + (size_t)o * lay.out_bytes;
        // the bulk stores that last used output stage `o` must have finished READING shared memory
        if (tid == 0) bulk_wait_read<OUT_STAGES - 1>();
        __syncthreads();
        mbar_wait(&full[s], parity);

        const long long e = (long long)tile * epb + env_local;
        const real *act = reinterpret_cast<const real *>(st + lay.act) + (size_t)env_local * p.A;
        const real *soc = reinterpret_cast<const real *>(st + lay.soc) + (size_t)env_local * p.N;
        const Rec<real> *rec = reinterpret_cast<const Rec<real> *>(st + lay.rec) + (size_t)env_local * p.N;
        EnvSt<real> es = reinterpret_cast<const EnvSt<real> *>(st + lay.envst)[env_local];
        float *obs = reinterpret_cast<float *>(ot + lay.obs) + (size_t)env_local * p.D;
        real *soc_out = reinterpret_cast<real *>(ot + lay.osoc) + (size_t)env_local * p.N;
        real reward;
        uint8_t done;
        uint32_t err;
        env_step<real, L, false>(p, tb, e, lane, act, soc, soc_out, rec, p.rec + (size_t)e * p.N, es, obs,
                                 p.tobs ? p.tobs + (size_t)e * p.D : nullptr, reward, done, err,
                                 p.diag ? p.diag + (size_t)e * D_COUNT : nullptr);
        if (lane == 0) {
            reinterpret_cast<EnvSt<real> *>(ot + lay.oenvst)[env_local] = es;
            reinterpret_cast<real *>(ot + lay.rew)[env_local] = reward;
            (ot + lay.done)[env_local] = done;
            if (err && p.err) atomicOr(p.err + e, err);
        }
        fence_proxy_async();   // make this thread's shared-memory writes visible to the copy engine
        __syncthreads();
        if (tid == 0) {
            const size_t e0 = (size_t)tile * epb;
            bulk_s2g(p.obs + e0 * p.D, ot + lay.obs, obs_bytes);
            bulk_s2g(p.soc + e0 * p.N, ot + lay.osoc, soc_bytes);
            bulk_s2g(p.envst + e0, ot + lay.oenvst, est_bytes);
            bulk_s2g(p.reward + e0, ot + lay.rew, rew_bytes);
            bulk_s2g(p.done + e0, ot + lay.done, (uint32_t)epb);
            bulk_commit();
            const int next = tile + in_stages * gridDim.x;   // input stage `s` is free again: refill it
            if (next < num_tiles) issue_loads(next, s);
        }
    }
    if (tid == 0) bulk_wait_read<0>();   // shared memory must stay valid until the last stores have read it
}

}  // namespace sng
