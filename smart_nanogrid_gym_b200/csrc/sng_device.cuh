// sng_device.cuh -- device-side types and the per-environment step body.
//
// One body, shared by every kernel variant (direct-global / TMA-staged, float / double,
// any lanes-per-env mapping).  File:line citations refer to the reference tree
// (smart_nanogrid_gym/...).  SURVEY.md section 2.3 is the step-by-step specification.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sng {

// ------------------------------------------------------------------------------------------
// State layout in HBM (DESIGN.md "Data layout")
// ------------------------------------------------------------------------------------------
// Current-vehicle record of one charging spot.  hdr = arr | dep << 8 | cap << 16 | next << 24
// (arrival step, departure step, capacity kWh, arrival step of the next planned vehicle).
// arr == 0xFF: no vehicle has been assigned to the spot yet.
template <typename real> struct Rec;
template <> struct Rec<float> { uint32_t hdr; float soc0; float req; };                     // 12 B
template <> struct Rec<double> { uint32_t hdr; uint32_t pad; double soc0; double req; };    // 24 B

// Per-env scalars.  t_ep = t | episode << 8.
template <typename real> struct EnvSt;
template <> struct __align__(16) EnvSt<float> { float soc_b, pv_shift, ep_ret; uint32_t t_ep; };          // 16 B
template <> struct __align__(16) EnvSt<double> { double soc_b, pv_shift, ep_ret; uint32_t t_ep, pad; };   // 32 B

constexpr uint32_t kNoVehicle = 0xFFu;
constexpr int kMaxVehicles = 8;
constexpr int kDepTab = 256;

enum : int { PEN_NONE = 0, PEN_ON_DEPARTURE = 1, PEN_SPARSE = 2, PEN_DENSE = 3 };
enum : int { MODE_SAMPLE = 0, MODE_REPLAY = 1 };
enum : uint32_t { FLAG_NEG_DEMAND = 1u, FLAG_BATT_SOC_GT1 = 2u, FLAG_NAN_ACTION = 4u };
enum : int { D_TOTAL_CH = 0, D_TOTAL_DIS, D_SOLAR, D_BATT_POWER, D_GRID_POWER, D_GRID_COST, D_PEN_VEH, D_PEN_BATT, D_COUNT };

template <typename real> struct Params {
    long long n_envs;
    unsigned long long gid0;
    uint32_t seed_lo, seed_hi;
    int N, T, H, A, D;
    int pv, batt, v2x, pen_mode, diff_cap, req_soc, default_cap, auto_reset, mode;
    int i4, i10, i1;  // int(4/dt), int(10/dt), int(1/dt): charging_station.py:271-279
    int off_soc, off_dep, off_batt;
    real dt, ev_pmax, ev_eff, b_cap, b_pmax, b_eff, b_dod, b_soc0, sell, cost_w, batt_w, margin;
    // shared read-only tables in global memory
    const real *pv_power, *irr_norm, *price, *price_norm;  // [table_len]
    const float *dep_norm;                                 // [kDepTab]: float(k / 24.0)
    // caller-owned buffers
    const real *actions;
    float *obs;
    real *reward;
    uint8_t *done;
    float *tobs;
    real *soc;
    Rec<real> *rec;
    EnvSt<real> *envst;
    const Rec<real> *plan;
    uint32_t *err;
    real *diag;
    real *last_ret;
};

// Shared read-only tables (global memory in the direct kernels, a shared-memory copy in the tiled one).
template <typename real> struct Tables {
    const real *pv_power, *irr_norm, *price, *price_norm;  // [table_len]
    const float *dep_norm;                                 // [kDepTab]
};
template <typename real> __device__ __forceinline__ Tables<real> global_tables(const Params<real> &p)
{
    Tables<real> tb;
    tb.pv_power = p.pv_power; tb.irr_norm = p.irr_norm; tb.price = p.price; tb.price_norm = p.price_norm;
    tb.dep_norm = p.dep_norm;
    return tb;
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ uint32_t make_hdr(uint32_t arr, uint32_t dep, uint32_t cap, uint32_t next)
{
    return arr | (dep << 8) | (cap << 16) | (next << 24);
}

// One arrival trial for (global env, spot, episode, step tn): the reference's per-step draw
// `round(rand() - 0.1) == 1` (p = 0.4) and, on arrival, SoC / requested SoC / capacity /
// departure (charging_station.py:213-237, 257-279).  Mirrors oracle ngo_sample_episode bit for bit.
template <typename real>
__device__ __forceinline__ bool sample_arrival(const Params<real> &p, unsigned long long stream, uint32_t episode,
                                               int tn, Rec<real> &r)
{
    uint32_t x[4];
    philox4x32_10((uint32_t)stream, (uint32_t)(stream >> 32), episode, (uint32_t)tn, p.seed_lo, p.seed_hi, x);
    if (x[0] <= 0x99999999u) return false;
    const float u1 = __fmul_rn((float)(x[1] >> 8), 5.9604644775390625e-08f);
    const float soc0 = __fmaf_rn(0.8f, u1, 0.1f);
    float rq = 1.0f;
    if (p.req_soc) {
        const float u2 = __fmul_rn((float)(x[2] >> 8), 5.9604644775390625e-08f);
        const float lo = (soc0 <= 0.9f) ? __fadd_rn(soc0, 0.1f) : 1.0f;
        rq = __fmaf_rn(__fsub_rn(1.0f, lo), u2, lo);
    }
    const uint32_t cap = p.diff_cap ? 15u + (((x[3] >> 16) * 105u) >> 16) : (uint32_t)p.default_cap;
    const int low = tn + p.i4;
    const int up = min(tn + p.i10, p.T + p.i1);
    const int dep = (low >= up) ? low : low + (int)(((x[3] & 0xFFFFu) * (uint32_t)(up - low)) >> 16);
    r.hdr = make_hdr((uint32_t)tn, (uint32_t)dep, cap, kNoVehicle);
    r.soc0 = (real)soc0;
    r.req = (real)rq;
    return true;
}

// random.randint(0, 180) / 100 (envs/smart_nanogrid_environment.py:181,349)
template <typename real>
__device__ __forceinline__ real sample_pv_shift(const Params<real> &p, unsigned long long gid, uint32_t episode)
{
    const unsigned long long stream = gid * (unsigned long long)p.N;
    uint32_t x[4];
    philox4x32_10((uint32_t)stream, (uint32_t)(stream >> 32), episode, 0xFFFFFFFFu, p.seed_lo, p.seed_hi, x);
    const uint32_t k = __umulhi(x[0], 181u);
    return (real)__fdiv_rn((float)k, 100.0f);
}

// Make `r` the record that governs step tn: sample a new arrival (sampling mode) or fetch the
// planned vehicle arriving at tn (replay mode).  Returns true when `r` changed.
template <typename real>
__device__ __forceinline__ bool advance_spot(const Params<real> &p, long long e, int i, uint32_t episode, int tn,
                                             Rec<real> &r)
{
    const uint32_t arr = r.hdr & 0xFFu, dep = (r.hdr >> 8) & 0xFFu;
    if (p.mode == MODE_SAMPLE) {
        // the generator draws only while the spot is free; on the departure step itself
        // (tn == dep) no draw happens (charging_station.py:213,239-251)
        const bool is_free = (arr == kNoVehicle) || (tn > (int)dep);
        if (!is_free) return false;
        const unsigned long long stream = (p.gid0 + (unsigned long long)e) * (unsigned long long)p.N + (unsigned)i;
        return sample_arrival(p, stream, episode, tn, r);
    }
    if ((r.hdr >> 24) != (uint32_t)tn || p.plan == nullptr) return false;
    const Rec<real> *pl = p.plan + ((size_t)e * p.N + i) * kMaxVehicles;
    for (int v = 0; v < kMaxVehicles; ++v) {
        const Rec<real> c = pl[v];
        if ((c.hdr & 0xFFu) == (uint32_t)tn) { r = c; return true; }
    }
    return false;
}

// numpy's pairwise float64 sum (n <= 128 branch) for the bit-faithful double build:
// charger_power_values[mask].sum(), utils/charging_station.py:293-294.
__device__ inline double numpy_sum(const double *a, int n)
{
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; i++) res = __dadd_rn(res, a[i]);
        return res;
    }
    double r[8];
    int i;
    for (int j = 0; j < 8; j++) r[j] = a[j];
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], a[i + j]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; i++) res = __dadd_rn(res, a[i]);
    return res;
}

// Reductions over the L lanes that share one env (L divides 32; every group uses its own mask,
// so groups of a warp whose env index is out of range may have exited).
template <int L> __device__ __forceinline__ uint32_t group_mask()
{
    if (L >= 32) return 0xffffffffu;
    const uint32_t lane = threadIdx.x & 31u;
    return ((1u << (L & 31)) - 1u) << (lane & ~(uint32_t)(L - 1));
}
template <int L, typename T> __device__ __forceinline__ T group_sum(T v)
{
    const uint32_t m = group_mask<L>();
#pragma unroll
    for (int off = L / 2; off > 0; off >>= 1) v += __shfl_xor_sync(m, v, off, L);
    return v;
}
template <int L> __device__ __forceinline__ uint32_t group_or(uint32_t v)
{
    const uint32_t m = group_mask<L>();
#pragma unroll
    for (int off = L / 2; off > 0; off >>= 1) v |= __shfl_xor_sync(m, v, off, L);
    return v;
}
template <int L> __device__ __forceinline__ void group_sync()
{
    if (L > 1) __syncwarp(group_mask<L>());
}

// Env-level part of the observation (envs/smart_nanogrid_environment.py:197-205,
// central_management_system.py:53-60): disturbances now and `H` steps ahead, battery SoC.
template <typename real>
__device__ __forceinline__ void write_obs_env(const Params<real> &p, const Tables<real> &tb, float *obs, int t,
                                              real shift, real soc_b)
{
    int k = 0;
    if (p.pv) {
        obs[k++] = (float)(tb.irr_norm[t] * shift);
        obs[k++] = (float)tb.price_norm[t];
        for (int j = 1; j <= p.H; ++j) obs[k++] = (float)(tb.irr_norm[t + j] * shift);
        for (int j = 1; j <= p.H; ++j) obs[k++] = (float)tb.price_norm[t + j];
    } else {
        obs[k++] = (float)tb.price_norm[t];
        for (int j = 1; j <= p.H; ++j) obs[k++] = (float)tb.price_norm[t + j];
    }
    if (p.batt) obs[p.off_batt] = (float)soc_b;
}

// Begin an episode at t = 0 (SmartNanogridEnv.reset, envs/smart_nanogrid_environment.py:311-351):
// sampling mode runs the step-0 arrival trial of every spot, replay mode rewinds to the first
// planned vehicle; per-spot SoC state is cleared (clear_initialisation_variables,
// charging_station.py:138-150) and the reset observation is written.
template <typename real, int L>
__device__ __forceinline__ void begin_episode(const Params<real> &p, const Tables<real> &tb, long long e, int lane,
                                              uint32_t episode, real shift, real soc_b, real *soc, Rec<real> *rec,
                                              float *obs)
{
    const int N = p.N;
    for (int i = lane; i < N; i += L) {
        Rec<real> r;
        if (p.mode == MODE_SAMPLE || p.plan == nullptr) {
            r.hdr = make_hdr(kNoVehicle, 0, 0, kNoVehicle);
            r.soc0 = 0;
            r.req = 0;
            advance_spot(p, e, i, episode, 0, r);
        } else {
            r = p.plan[((size_t)e * N + i) * kMaxVehicles];   // slot 0 = first vehicle of the day
        }
        rec[i] = r;
        soc[i] = 0;
        // the dense SoC array holds the arrival SoC at slot `arr` (charging_station.py:257-259),
        // so the reset observation shows it for vehicles arriving at t = 0
        const int arr = (int)(r.hdr & 0xFFu), dep = (int)((r.hdr >> 8) & 0xFFu);
        const bool present = ((uint32_t)arr != kNoVehicle) && arr == 0 && 0 < dep;
        obs[p.off_soc + i] = present ? (float)r.soc0 : 0.0f;
        obs[p.off_dep + i] = present ? tb.dep_norm[dep] : 0.0f;
    }
    if (lane == 0) write_obs_env(p, tb, obs, 0, shift, soc_b);   // battery SoC survives resets (quirk Q8)
}

// ------------------------------------------------------------------------------------------
// The step of ONE environment, executed cooperatively by L lanes (lane l owns spots
// l, l+L, ...).  All row pointers may point to global or shared memory.
//   act [A] in     soc [N] in -> soc_out [N]     rec [N] in -> rec_out [N] (written only when a
//   vehicle arrives)     es in/out (same value in all lanes)     obs [D] out     tobs [D] out or null
// In the direct kernels soc == soc_out and rec == rec_out; the tiled kernel reads from the staged
// input tile and writes SoC to the output tile and changed records straight to global memory.
// EXACT (double, L == 1 only): reproduces numpy's summation order.
// ------------------------------------------------------------------------------------------
template <typename real, int L, bool EXACT>
__device__ __forceinline__ void env_step(const Params<real> &p, const Tables<real> &tb, long long e, int lane,
                                         const real *act, const real *soc, real *soc_out, const Rec<real> *rec,
                                         Rec<real> *rec_out, EnvSt<real> &es, float *obs, float *tobs,
                                         real &reward_out, uint8_t &done_out, uint32_t &err_out, real *diag)
{
    const int N = p.N;
    const int t = (int)(es.t_ep & 0xFFu);
    uint32_t episode = es.t_ep >> 8;
    real pos = 0, neg = 0, pen_veh = 0;
    uint32_t err = 0;
    double cpos[EXACT ? 256 : 1], cneg[EXACT ? 256 : 1];
    int npos = 0, nneg = 0;

    // ---- per-spot phase: ChargingStation.simulate_vehicle_charging (charging_station.py:281-300),
    //      Charger.charge_or_discharge_vehicle (charger.py:37-140) and the lagged undercharge
    //      penalty (penaliser.py:39-87) ----
    for (int i = lane; i < N; i += L) {
        const Rec<real> r = rec[i];
        const real s_prev = soc[i];
        const real a = act[i];
        const int arr = (int)(r.hdr & 0xFFu), dep = (int)((r.hdr >> 8) & 0xFFu);
        const bool has = (uint32_t)arr != kNoVehicle;
        if (a != a) err |= FLAG_NAN_ACTION;

        // check set computed by the previous observe() at t_obs = t-1; column t-1 of soc / req
        if (t >= 1 && has && arr <= t - 1 && t - 1 < dep) {
            const int togo = dep - (t - 1);
            const bool allowed = (p.pen_mode == PEN_DENSE) || (p.pen_mode == PEN_SPARSE && togo <= 3) ||
                                 (p.pen_mode == PEN_ON_DEPARTURE && togo == 1);
            if (allowed) {
                const real lower = p.margin * r.req;       // penaliser.py:72
                if (s_prev < r.req - lower) {              // :78
                    const real d = (r.req - s_prev) * (real)10;
                    pen_veh = pen_veh + d * d;             // :79 (Python `** 2`)
                }
            }
        }

        const bool present = has && arr <= t && t < dep;   // charger.occupancy[t] == 1
        real P = 0, s_new = 0;
        if (present) {
            const real s_in = (arr == t) ? r.soc0 : s_prev;  // `timestep in arrivals`, charger.py:62-67
            const real cap = (real)((r.hdr >> 16) & 0xFFu);
            if (a == (real)0) {                              // charger.py:38-45
                s_new = s_in;
            } else if (a > (real)0) {                        // charge_vehicle, charger.py:58-90
                const real power = a * p.ev_pmax * p.ev_eff;
                const real calc = s_in + (power * p.dt) / cap;
                s_new = ((real)1 < calc) ? (real)1 : calc;   // power is NOT reduced when clamped
                P = power;
            } else {                                         // discharge_vehicle, charger.py:108-140
                const real power = a * p.ev_pmax * p.ev_eff;
                const real calc = s_in + (power * p.dt) / cap;
                // flag = ceil(0.5 * (1 + sign(calc))) == 1 iff calc >= 0 (quirk Q1)
                P = (calc >= (real)0) ? -((s_in * cap) / p.dt) : power;
                s_new = (calc > (real)0) ? calc : (real)0;
            }
        }
        if (EXACT) {
            if (P < 0) cneg[nneg++] = (double)P;
            if (P > 0) cpos[npos++] = (double)P;
        } else {
            if (P < 0) neg += P;
            if (P > 0) pos += P;
        }
        soc_out[i] = s_new;
        obs[p.off_soc + i] = (float)s_new;                               // charging_station.py:114-117
        obs[p.off_dep + i] = present ? tb.dep_norm[dep - t] : 0.0f;       // :92-112, "/ 24" env:208
    }
    if (EXACT) {
        neg = (real)numpy_sum(cneg, nneg);
        pos = (real)numpy_sum(cpos, npos);
    } else if (L > 1) {
        pos = group_sum<L>(pos);
        neg = group_sum<L>(neg);
        pen_veh = group_sum<L>(pen_veh);
        err = group_or<L>(err);
    }

    // ---- env-level phase: CentralManagementSystem.manage_nanogrid (central_management_system.py:99-113) ----
    const real total_power = pos + neg;                                   // :105
    if (total_power < (real)0 && !p.v2x) err |= FLAG_NEG_DEMAND;          // reference raises, :158-159
    const real solar = p.pv ? tb.pv_power[t] * es.pv_shift : (real)0;      // :99-103
    real rem = total_power - solar;                                       // :167
    real soc_b = es.soc_b, batt_power = 0, pen_b = 0;
    if (p.batt) {                                                         // battery_energy_storage_system.py:30-106
        const real ab = act[N];                                           // actions[-1], :88-89
        if (ab != ab) err |= FLAG_NAN_ACTION;
        if (ab > (real)0) {                                               // charge, :46-74
            const real power = ab * p.b_pmax * p.b_eff;
            const real calc = soc_b + (power * p.dt) / p.b_cap;
            soc_b = ((real)1 < calc) ? (real)1 : calc;
            batt_power = power;
            rem = rem + power;                                            // -((-rem) - power)
        } else if (ab < (real)0) {                                        // discharge, :76-106
            real power = ab * p.b_pmax * p.b_eff;
            const real calc = soc_b + (power * p.dt) / p.b_cap;
            if (calc < (real)0) power = -((soc_b * p.b_cap) / p.dt);      // :82-94
            soc_b = (calc > (real)0) ? calc : (real)0;                    // :98
            batt_power = power;
            rem = rem + power;                                            // :102
        }
        if (soc_b < p.b_dod) {                                            // penaliser.py:104-111
            const real d = (p.b_dod - soc_b) * (real)10;
            pen_b = d * d;
        } else if (!(soc_b <= (real)1)) {
            err |= FLAG_BATT_SOC_GT1;
        }
    }
    const real energy = rem * p.dt;                                       // central_management_system.py:107
    const real price = tb.price[t];
    const real cost = (energy < (real)0) ? energy * p.sell * price : energy * price;   // accountant.py:26-32
    const real total_pen = p.batt_w * pen_b + pen_veh;                    // penaliser.py:181
    const real total_cost = p.cost_w * fabs(cost) + total_pen;            // accountant.py:35
    const real reward = -total_cost;                                      // ...environment.py:183

    if (lane == 0) {
        write_obs_env(p, tb, obs, t, es.pv_shift, soc_b);                 // obs at the pre-increment t, :173
        if (diag) {
            diag[D_TOTAL_CH] = pos; diag[D_TOTAL_DIS] = neg; diag[D_SOLAR] = solar;
            diag[D_BATT_POWER] = batt_power; diag[D_GRID_POWER] = rem; diag[D_GRID_COST] = cost;
            diag[D_PEN_VEH] = pen_veh; diag[D_PEN_BATT] = pen_b;
        }
    }

    // ---- t += 1, termination, auto-reset (...environment.py:174-181, 311-351) ----
    const int tn = t + 1;
    const bool is_done = (tn == p.T);
    real ep_ret = es.ep_ret + reward;
    real shift = es.pv_shift;
    if (!is_done) {
        for (int i = lane; i < N; i += L) {
            Rec<real> r = rec[i];
            if (advance_spot(p, e, i, episode, tn, r)) rec_out[i] = r;
        }
        es.t_ep = (episode << 8) | (uint32_t)tn;
    } else {
        if (p.last_ret && lane == 0) p.last_ret[e] = ep_ret;
        ep_ret = 0;
        if (p.auto_reset) {
            group_sync<L>();
            if (tobs) {
                for (int k = lane; k < p.D; k += L) tobs[k] = obs[k];
                group_sync<L>();
            }
            episode = (episode + 1u) & 0xFFFFFFu;
            if (p.mode == MODE_SAMPLE) shift = sample_pv_shift(p, p.gid0 + (unsigned long long)e, episode);
            begin_episode<real, L>(p, tb, e, lane, episode, shift, soc_b, soc_out, rec_out, obs);
        }
        es.t_ep = (episode << 8);                                         // t wraps to 0, :178
    }
    es.soc_b = soc_b;
    es.pv_shift = shift;
    es.ep_ret = ep_ret;
    reward_out = reward;
    done_out = is_done ? 1 : 0;
    err_out = err;
}

}  // namespace sng
