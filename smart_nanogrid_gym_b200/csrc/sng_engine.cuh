// sng_engine.cuh -- kernels and host-side launch logic, templated on the arithmetic type.
// Instantiated as Engine<float,false> (sng_f32.cu) and Engine<double,true> (sng_f64.cu, built
// with -fmad=false so that no multiply-add is contracted).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/sng.h"
#include "sng_device.cuh"
#include "sng_tiled.cuh"

namespace sng {

struct EngineBase {
    virtual ~EngineBase() {}
    virtual int bind(const sng_buffers *b) = 0;
    virtual int reset(uint64_t seed, const uint8_t *mask, int reset_battery, cudaStream_t st) = 0;
    virtual int load_schedule(const sng_schedule_view *v, cudaStream_t st) = 0;
    virtual int step(cudaStream_t st) = 0;
    virtual int rollout(const void *actions, float *obs, void *reward, uint8_t *done, int n_steps, cudaStream_t st) = 0;
    virtual int step_host(const void *a, float *obs, void *rew, uint8_t *done, cudaStream_t st) = 0;
    virtual int sample_plan(cudaStream_t st) = 0;
    virtual int error_flags(uint32_t *out, cudaStream_t st) = 0;
    virtual int set_tuning(int lanes, int tile, int bulk) = 0;
    virtual int set_pipeline(int in_stages, int out_stages, int ctas_per_sm) = 0;
    int64_t launches = 0;
    std::string error;
};

EngineBase *make_engine_f32(const sng_config &cfg, int device, std::string &err);
EngineBase *make_engine_f64(const sng_config &cfg, int device, std::string &err);

#define SNG_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (call);                                                                \
        if (_e != cudaSuccess) {                                                                \
            error = std::string(#call) + ": " + cudaGetErrorString(_e);                         \
            return SNG_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

// ------------------------------------------------------------------------------------------
// Kernels, direct-global variant: L lanes per env, rows addressed in global memory.
// ------------------------------------------------------------------------------------------
template <typename real, int L, bool EXACT>
__global__ void __launch_bounds__(256) step_direct_kernel(const Params<real> p)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = gtid / L;
    const int lane = (int)(gtid % L);
    if (e >= p.n_envs) return;
    EnvSt<real> es = p.envst[e];
    real reward;
    uint8_t done;
    uint32_t err;
    real *soc = p.soc + (size_t)e * p.N;
    Rec<real> *rec = p.rec + (size_t)e * p.N;
    env_step<real, L, EXACT>(p, global_tables(p), e, lane, p.actions + (size_t)e * p.A, soc, soc, rec, rec, es,
                             p.obs + (size_t)e * p.D, p.tobs ? p.tobs + (size_t)e * p.D : nullptr, reward, done, err,
                             p.diag ? p.diag + (size_t)e * D_COUNT : nullptr);
    if (lane == 0) {
        p.envst[e] = es;
        p.reward[e] = reward;
        p.done[e] = done;
        if (err && p.err) atomicOr(p.err + e, err);
    }
}

template <typename real, int L, bool EXACT>
__global__ void __launch_bounds__(256) rollout_direct_kernel(const Params<real> p, const real *actions, float *obs,
                                                            real *reward, uint8_t *done, int n_steps)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = gtid / L;
    const int lane = (int)(gtid % L);
    if (e >= p.n_envs) return;
    EnvSt<real> es = p.envst[e];
    uint32_t err_all = 0;
    for (int s = 0; s < n_steps; ++s) {
        const size_t row = (size_t)s * p.n_envs + e;
        real r;
        uint8_t d;
        uint32_t err;
        real *soc = p.soc + (size_t)e * p.N;
        Rec<real> *rec = p.rec + (size_t)e * p.N;
        env_step<real, L, EXACT>(p, global_tables(p), e, lane, actions + row * p.A, soc, soc, rec, rec, es,
                                 obs + row * p.D, nullptr, r, d, err, nullptr);
        group_sync<L>();
        err_all |= err;
        if (lane == 0) {
            reward[row] = r;
            done[row] = d;
        }
    }
    if (lane == 0) {
        p.envst[e] = es;
        if (err_all && p.err) atomicOr(p.err + e, err_all);
    }
}

template <typename real, int L>
__global__ void __launch_bounds__(256) reset_kernel(const Params<real> p, const uint8_t *mask, int init,
                                                   int new_episode, int reset_battery)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = gtid / L;
    const int lane = (int)(gtid % L);
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
    if (p.mode == MODE_SAMPLE) shift = sample_pv_shift(p, p.gid0 + (unsigned long long)e, episode);
    begin_episode<real, L>(p, global_tables(p), e, lane, episode, shift, soc_b, p.soc + (size_t)e * p.N,
                           p.rec + (size_t)e * p.N, p.obs + (size_t)e * p.D);
    if (lane == 0) {
        es.soc_b = soc_b;
        es.pv_shift = shift;
        es.ep_ret = 0;
        es.t_ep = episode << 8;
        p.envst[e] = es;
        if (init && p.err) p.err[e] = 0;
    }
}

// Whole-day schedule of the current episode, one thread per (env, spot): the reference's
// generator loop (charging_station.py:200-255) over the same Philox trials the lazy sampler uses.
template <typename real>
__global__ void __launch_bounds__(256) sample_plan_kernel(const Params<real> p, Rec<real> *plan)
{
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid >= p.n_envs * p.N) return;
    const long long e = gtid / p.N;
    const int i = (int)(gtid % p.N);
    const uint32_t episode = p.envst[e].t_ep >> 8;
    Rec<real> *pl = plan + (size_t)gtid * kMaxVehicles;
    Rec<real> cur;
    cur.hdr = make_hdr(kNoVehicle, 0, 0, kNoVehicle);
    cur.soc0 = 0;
    cur.req = 0;
    int nv = 0;
    for (int t = 0; t < p.T; ++t) {
        Rec<real> r = cur;
        if (advance_spot(p, e, i, episode, t, r) && nv < kMaxVehicles) {
            if (nv > 0) pl[nv - 1].hdr = (pl[nv - 1].hdr & 0x00FFFFFFu) | ((uint32_t)t << 24);
            pl[nv++] = r;
            cur = r;
        }
    }
    for (int v = nv; v < kMaxVehicles; ++v) {
        Rec<real> z;
        z.hdr = make_hdr(kNoVehicle, 0, 0, kNoVehicle);
        z.soc0 = 0;
        z.req = 0;
        pl[v] = z;
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
    int lanes = 0;  // 0 = auto
    int tile = 0, bulk = 1;
    int in_stages = 0, out_stages = 0, ctas_per_sm = 0;  // 0 = auto
    int num_sms = 148;
    size_t smem_optin = 0;
    void *d_tables = nullptr;
    uint32_t *d_flag = nullptr;

    int init(const sng_config &c, int dev)
    {
        cfg = c;
        device = dev;
        if (c.n_spots < 1 || c.n_spots > SNG_MAX_SPOTS || c.n_steps < 1 || c.n_steps > 250 || c.n_envs < 1 ||
            c.table_len < c.n_steps + c.horizon || c.table_len > SNG_MAX_TABLE || !c.price || !c.price_norm ||
            (c.pv && (!c.pv_power || !c.irr_norm)) || c.penalty_mode < 0 || c.penalty_mode > 3 || c.horizon < 0) {
            error = "sng_create: invalid configuration";
            return SNG_ERR_ARG;
        }
        if (EXACT && c.n_spots > 128) {
            error = "sng_create: the float64 validation build supports at most 128 spots";
            return SNG_ERR_UNSUPPORTED;
        }
        SNG_CUDA(cudaSetDevice(dev));
        {
            int v = 0;
            SNG_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
            num_sms = v;
            SNG_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
            smem_optin = (size_t)v;
        }
        memset(&p, 0, sizeof(p));
        p.n_envs = c.n_envs;
        p.gid0 = (unsigned long long)c.env_gid0;
        p.N = c.n_spots; p.T = c.n_steps; p.H = c.horizon;
        p.pv = c.pv != 0; p.batt = c.batt != 0; p.v2x = c.v2x != 0;
        p.A = p.N + p.batt;
        const int nd = (1 + p.pv) * (1 + p.H);
        p.off_soc = nd; p.off_dep = nd + p.N; p.off_batt = nd + 2 * p.N;
        p.D = nd + 2 * p.N + p.batt;
        p.pen_mode = c.penalty_mode; p.diff_cap = c.diff_cap != 0; p.req_soc = c.req_soc != 0;
        p.default_cap = c.default_cap; p.auto_reset = c.auto_reset != 0; p.mode = MODE_SAMPLE;
        p.i4 = (int)(4.0 / c.dt); p.i10 = (int)(10.0 / c.dt); p.i1 = (int)(1.0 / c.dt);
        p.dt = (real)c.dt; p.ev_pmax = (real)c.ev_pmax; p.ev_eff = (real)c.ev_eff;
        p.b_cap = (real)c.b_cap; p.b_pmax = (real)c.b_pmax; p.b_eff = (real)c.b_eff; p.b_dod = (real)c.b_dod;
        p.b_soc0 = (real)c.b_soc0; p.sell = (real)c.sell_coeff; p.cost_w = (real)c.cost_weight;
        p.batt_w = (real)c.batt_pen_w; p.margin = (real)c.margin;
        // shared tables: 4 x table_len reals + the departure-normalisation table
        const int n = c.table_len;
        std::vector<real> h(4 * (size_t)n, (real)0);
        for (int k = 0; k < n; ++k) {
            h[k] = c.pv ? (real)c.pv_power[k] : (real)0;
            h[n + k] = c.pv ? (real)c.irr_norm[k] : (real)0;
            h[2 * n + k] = (real)c.price[k];
            h[3 * n + k] = (real)c.price_norm[k];
        }
        std::vector<float> dn(kDepTab);
        for (int k = 0; k < kDepTab; ++k) dn[k] = (float)((double)k / c.dep_norm);  // ...environment.py:208,228
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
        if (d_tables) cudaFree(d_tables);
        if (d_flag) cudaFree(d_flag);
    }

    int bind(const sng_buffers *b) override
    {
        if (!b || b->struct_size != sizeof(sng_buffers)) { error = "sng_bind: bad struct_size"; return SNG_ERR_ARG; }
        if (!b->actions || !b->obs || !b->reward || !b->done || !b->soc || !b->rec || !b->envst) {
            error = "sng_bind: actions, obs, reward, done, soc, rec and envst are required";
            return SNG_ERR_ARG;
        }
        buf = *b;
        p.actions = (const real *)b->actions; p.obs = b->obs; p.reward = (real *)b->reward; p.done = b->done;
        p.tobs = b->terminal_obs; p.soc = (real *)b->soc; p.rec = (Rec<real> *)b->rec;
        p.envst = (EnvSt<real> *)b->envst; p.plan = (const Rec<real> *)b->plan; p.err = b->err;
        p.diag = (real *)b->diag; p.last_ret = (real *)b->last_return;
        bound = true;
        return SNG_OK;
    }

    int auto_lanes() const
    {
        if (EXACT) return 1;
        if (lanes > 0) return lanes;
        int l = 1;
        while (l < p.N && l < 32) l <<= 1;  // one (sub-)warp per env, spots on lanes
        return l;
    }

    template <typename F> int dispatch_lanes(F &&f)
    {
        if constexpr (EXACT) {
            return f(std::integral_constant<int, 1>());
        } else {
            switch (auto_lanes()) {
            case 1: return f(std::integral_constant<int, 1>());
            case 2: return f(std::integral_constant<int, 2>());
            case 4: return f(std::integral_constant<int, 4>());
            case 8: return f(std::integral_constant<int, 8>());
            case 16: return f(std::integral_constant<int, 16>());
            default: return f(std::integral_constant<int, 32>());
            }
        }
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
        return dispatch_lanes([&](auto lc) -> int {
            constexpr int L = decltype(lc)::value;
            reset_kernel<real, L><<<grid_for(p.n_envs * L), 256, 0, st>>>(p, mask, init, new_episode, reset_battery);
            ++launches;
            SNG_CUDA(cudaGetLastError());
            return (int)SNG_OK;
        });
    }

    int reset(uint64_t seed, const uint8_t *mask, int reset_battery, cudaStream_t st) override
    {
        int rc = check_ready(false);
        if (rc) return rc;
        SNG_CUDA(cudaSetDevice(device));
        const bool first = !started;
        if (first && mask) { error = "sng_reset: the first reset must cover all envs (mask = NULL)"; return SNG_ERR_STATE; }
        p.seed_lo = (uint32_t)seed;
        p.seed_hi = (uint32_t)(seed >> 32);
        p.mode = MODE_SAMPLE;
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
        SNG_CUDA(cudaSetDevice(device));
        const long long E = p.n_envs;
        const int N = p.N, V = v->n_slots;
        std::vector<Rec<real>> plan((size_t)E * N * kMaxVehicles);
        for (long long e = 0; e < E; ++e)
            for (int i = 0; i < N; ++i) {
                const size_t sp = (size_t)e * N + i;
                const int nv = v->n_veh[sp];
                if (nv < 0 || nv > V) { error = "sng_load_schedule: n_veh out of range"; return SNG_ERR_ARG; }
                int prev_dep = -1;
                for (int k = 0; k < kMaxVehicles; ++k) {
                    Rec<real> r;
                    memset(&r, 0, sizeof(r));
                    r.hdr = make_hdr_host(kNoVehicle, 0, 0, kNoVehicle);
                    if (k < nv) {
                        const size_t q = sp * V + k;
                        const int a = v->arr[q], d = v->dep[q], c = v->cap[q];
                        // invariants of generated schedules the step kernel relies on (schedule.py validate())
                        if (a < 0 || a >= p.T || d <= a || d > 250 || c < 1 || c > 255 || a <= prev_dep) {
                            error = "sng_load_schedule: invalid vehicle record (arrival/departure/capacity)";
                            return SNG_ERR_ARG;
                        }
                        prev_dep = d;
                        const uint32_t nxt = (k + 1 < nv) ? (uint32_t)v->arr[q + 1] : kNoVehicle;
                        r.hdr = make_hdr_host((uint32_t)a, (uint32_t)d, (uint32_t)c, nxt);
                        r.soc0 = (real)v->soc0[q];
                        r.req = (real)v->req[q];
                    }
                    plan[sp * kMaxVehicles + k] = r;
                }
            }
        SNG_CUDA(cudaMemcpyAsync(buf.plan, plan.data(), plan.size() * sizeof(Rec<real>), cudaMemcpyHostToDevice, st));
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
        rc = launch_reset(nullptr, 0, 0, 0, st);
        SNG_CUDA(cudaStreamSynchronize(st));  // `plan` staging vector goes out of scope
        if (rc == SNG_OK) started = true;
        return rc;
    }

    static uint32_t make_hdr_host(uint32_t arr, uint32_t dep, uint32_t cap, uint32_t next)
    {
        return arr | (dep << 8) | (cap << 16) | (next << 24);
    }

    // Parameters with every per-env pointer advanced by e0 envs (tail of a tiled launch).
    Params<real> offset_params(long long e0) const
    {
        Params<real> q = p;
        q.n_envs = p.n_envs - e0;
        q.gid0 = p.gid0 + (unsigned long long)e0;
        q.actions += (size_t)e0 * p.A; q.obs += (size_t)e0 * p.D; q.reward += e0; q.done += e0;
        if (q.tobs) q.tobs += (size_t)e0 * p.D;
        q.soc += (size_t)e0 * p.N; q.rec += (size_t)e0 * p.N; q.envst += e0;
        if (q.plan) q.plan += (size_t)e0 * p.N * kMaxVehicles;
        if (q.err) q.err += e0;
        if (q.diag) q.diag += (size_t)e0 * D_COUNT;
        if (q.last_ret) q.last_ret += e0;
        return q;
    }

    int launch_direct(const Params<real> &q, cudaStream_t st)
    {
        return dispatch_lanes([&](auto lc) -> int {
            constexpr int L = decltype(lc)::value;
            step_direct_kernel<real, L, EXACT><<<grid_for(q.n_envs * L), 256, 0, st>>>(q);
            ++launches;
            SNG_CUDA(cudaGetLastError());
            return (int)SNG_OK;
        });
    }

    static bool aligned16(const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

    template <int L, int OS> int launch_tiled(int epb, int num_tiles, int is, cudaStream_t st)
    {
        auto kern = step_tiled_kernel<real, L, OS>;
        const TileLayout lay = make_tile_layout<real>(epb, p.N, p.A, p.D, cfg.table_len, is, OS);
        SNG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
        int per_sm = 0;
        SNG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, epb * L, lay.total));
        if (per_sm < 1) { error = "tiled kernel does not fit on an SM"; return SNG_ERR_CUDA; }
        if (ctas_per_sm > 0 && per_sm > ctas_per_sm) per_sm = ctas_per_sm;
        long long grid = (long long)num_sms * per_sm;
        if (grid > num_tiles) grid = num_tiles;
        kern<<<(unsigned)grid, epb * L, lay.total, st>>>(p, epb, num_tiles, is, cfg.table_len);
        ++launches;
        SNG_CUDA(cudaGetLastError());
        return SNG_OK;
    }

    // Tile geometry of the bulk-copy path; returns false when the direct kernel must be used.
    bool plan_tiles(int &L, int &epb, int &is, int &os) const
    {
        if (EXACT || !bulk) return false;
        if (!aligned16(p.actions) || !aligned16(p.obs) || !aligned16(p.reward) || !aligned16(p.done) ||
            !aligned16(p.soc) || !aligned16(p.rec) || !aligned16(p.envst))
            return false;
        L = lanes > 0 ? lanes : 1;
        epb = tile > 0 ? tile : (L == 1 ? 128 : (L <= 4 ? 64 : (L <= 16 ? 32 : 16)));
        if (epb % 16 != 0 || epb * L > 1024 || epb * L < 32) return false;
        if (p.n_envs < epb) return false;
        os = out_stages > 0 ? out_stages : 1;
        if (os > 2) os = 2;
        is = in_stages > 0 ? in_stages : 2;
        while (is > 1 && make_tile_layout<real>(epb, p.N, p.A, p.D, cfg.table_len, is, os).total > smem_optin) --is;
        return make_tile_layout<real>(epb, p.N, p.A, p.D, cfg.table_len, is, os).total <= smem_optin;
    }

    int step(cudaStream_t st) override
    {
        int rc = check_ready(true);
        if (rc) return rc;
        if constexpr (!EXACT) {
            int L, epb, is, os;
            if (plan_tiles(L, epb, is, os)) {
                const int num_tiles = (int)(p.n_envs / epb);
                auto go = [&](auto lc) -> int {
                    constexpr int LL = decltype(lc)::value;
                    return os == 2 ? launch_tiled<LL, 2>(epb, num_tiles, is, st) : launch_tiled<LL, 1>(epb, num_tiles, is, st);
                };
                switch (L) {
                case 1: rc = go(std::integral_constant<int, 1>()); break;
                case 2: rc = go(std::integral_constant<int, 2>()); break;
                case 4: rc = go(std::integral_constant<int, 4>()); break;
                case 8: rc = go(std::integral_constant<int, 8>()); break;
                case 16: rc = go(std::integral_constant<int, 16>()); break;
                default: rc = go(std::integral_constant<int, 32>()); break;
                }
                if (rc) return rc;
                const long long done_envs = (long long)num_tiles * epb;
                if (done_envs < p.n_envs) return launch_direct(offset_params(done_envs), st);
                return SNG_OK;
            }
        }
        return launch_direct(p, st);
    }

    int rollout(const void *actions, float *obs, void *reward, uint8_t *done, int n_steps, cudaStream_t st) override
    {
        int rc = check_ready(true);
        if (rc) return rc;
        if (!actions) { error = "sng_rollout: in-kernel random actions are not implemented yet"; return SNG_ERR_UNSUPPORTED; }
        if (!obs || !reward || !done || n_steps < 1) { error = "sng_rollout: bad arguments"; return SNG_ERR_ARG; }
        return dispatch_lanes([&](auto lc) -> int {
            constexpr int L = decltype(lc)::value;
            rollout_direct_kernel<real, L, EXACT><<<grid_for(p.n_envs * L), 256, 0, st>>>(
                p, (const real *)actions, obs, (real *)reward, done, n_steps);
            ++launches;
            SNG_CUDA(cudaGetLastError());
            return (int)SNG_OK;
        });
    }

    int step_host(const void *a, float *obs, void *rew, uint8_t *done, cudaStream_t st) override
    {
        int rc = check_ready(true);
        if (rc) return rc;
        if (!a || !obs || !rew || !done) { error = "sng_step_host: null host buffer"; return SNG_ERR_ARG; }
        const size_t E = (size_t)p.n_envs;
        SNG_CUDA(cudaMemcpyAsync((void *)buf.actions, a, E * p.A * sizeof(real), cudaMemcpyHostToDevice, st));
        rc = step(st);
        if (rc) return rc;
        SNG_CUDA(cudaMemcpyAsync(obs, buf.obs, E * p.D * sizeof(float), cudaMemcpyDeviceToHost, st));
        SNG_CUDA(cudaMemcpyAsync(rew, buf.reward, E * sizeof(real), cudaMemcpyDeviceToHost, st));
        SNG_CUDA(cudaMemcpyAsync(done, buf.done, E, cudaMemcpyDeviceToHost, st));
        SNG_CUDA(cudaStreamSynchronize(st));
        return SNG_OK;
    }

    int sample_plan(cudaStream_t st) override
    {
        int rc = check_ready(true);
        if (rc) return rc;
        if (!buf.plan) { error = "sng_sample_plan: no `plan` buffer bound"; return SNG_ERR_STATE; }
        if (p.mode != MODE_SAMPLE) { error = "sng_sample_plan: handle is in replay mode"; return SNG_ERR_STATE; }
        sample_plan_kernel<real><<<grid_for(p.n_envs * p.N), 256, 0, st>>>(p, (Rec<real> *)buf.plan);
        ++launches;
        SNG_CUDA(cudaGetLastError());
        return SNG_OK;
    }

    int error_flags(uint32_t *out, cudaStream_t st) override
    {
        int rc = check_ready(false);
        if (rc) return rc;
        if (!out) { error = "sng_error_flags: null output"; return SNG_ERR_ARG; }
        *out = 0;
        if (!buf.err) return SNG_OK;
        SNG_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), st));
        or_reduce_kernel<<<148, 256, 0, st>>>(buf.err, p.n_envs, d_flag);
        ++launches;
        SNG_CUDA(cudaMemcpyAsync(out, d_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SNG_CUDA(cudaStreamSynchronize(st));
        return SNG_OK;
    }

    int set_tuning(int l, int t, int b) override
    {
        if (l != 0 && l != 1 && l != 2 && l != 4 && l != 8 && l != 16 && l != 32) {
            error = "sng_set_tuning: lanes_per_env must be 0 or a power of two <= 32";
            return SNG_ERR_ARG;
        }
        lanes = l; tile = t; bulk = b;
        return SNG_OK;
    }

    int set_pipeline(int is, int os, int cps) override
    {
        if (is < 0 || is > 8 || os < 0 || os > 2 || cps < 0) { error = "sng_set_pipeline: bad arguments"; return SNG_ERR_ARG; }
        in_stages = is; out_stages = os; ctas_per_sm = cps;
        return SNG_OK;
    }
};

}  // namespace sng
