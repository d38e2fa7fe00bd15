// sng_device.cuh -- device-side types, the counter-based schedule sampler and the per-environment
// step body.
//
// Thread mapping: ONE THREAD PER ENVIRONMENT, 32 consecutive envs per warp (four lanes per env, 8 envs
// per warp, for 64-spot stations whose rows would otherwise leave 8 warps per SM).  Per-spot state is a
// structure of arrays, plane-major and blocked by 32 envs ([3 planes][E/32][N][32]), so lane l of a warp
// reads plane f of spot i of its env from word  f*plane + (block*N + i)*32 + l : every state load /
// store of a warp is one full 128-byte line, a block's N lines of one plane are contiguous (1.25 KB at
// 10 spots: r3 -- with the planes interleaved per spot the same traffic cost 7 % more DRAM time), and all
// of a thread's accesses are constant offsets from one base pointer per plane.
// Spots are walked sequentially inside the thread; the station sums are kept per spot parity and combined
// in a fixed order, so results do not depend on the kernel variant or on how envs are split over GPUs.
// DESIGN.md section 3.1 has the measurements behind this choice (a warp-per-env mapping leaves 22 of 32
// lanes idle at 10 spots and is issue-bound).
//
// File:line citations refer to the reference tree (smart_nanogrid_gym/...).  SURVEY.md section 2.3
// is the step-by-step specification.
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

namespace sng {

// ------------------------------------------------------------------------------------------
// State layout in HBM (DESIGN.md "Data layout")
// ------------------------------------------------------------------------------------------
constexpr int kBlock = 32;          // envs per state block = lanes of a warp
constexpr int kPlanes = 3;          // per-spot state planes: header word, SoC, requested SoC
enum : int { PL_HDR = 0, PL_SOC = 1, PL_REQ = 2 };   // header and SoC (read every step) adjacent: 256-byte runs; the requested SoC last

// State words are as wide as `real` (uint32 / uint64) so that one array holds all three planes.
template <typename real> struct WordOf;
template <> struct WordOf<float> { typedef uint32_t type; };
template <> struct WordOf<double> { typedef unsigned long long type; };
__device__ __forceinline__ float word_to_real(uint32_t w, float) { return __uint_as_float(w); }
__device__ __forceinline__ double word_to_real(unsigned long long w, double) { return __longlong_as_double((long long)w); }
__device__ __forceinline__ uint32_t real_to_word(float x) { return __float_as_uint(x); }
__device__ __forceinline__ unsigned long long real_to_word(double x) { return (unsigned long long)__double_as_longlong(x); }
constexpr uint32_t kNoVehicle = 0xFFu;
constexpr int kMaxVehicles = 8;
constexpr int kDepTab = 208;          // departure-normalisation entries (dep - t <= 10 h / dt <= 93 steps)
constexpr int kGapTab = 48;           // thresholds of the geometric arrival gap (43 used)
constexpr int kSmemTab = kDepTab + kGapTab;   // words of the per-CTA shared-memory copy: [dep | gap]


// Per-spot header word: arr | dep << 8 | cap << 16 | next << 24
//   arr   arrival step of the current / last vehicle (0xFF: none yet this episode)
//   dep   its departure step (first step the spot is free again)
//   cap   its battery capacity in kWh
//   next  step at which the next vehicle arrives (0xFF: none left today)
__host__ __device__ __forceinline__ uint32_t make_hdr(uint32_t arr, uint32_t dep, uint32_t cap, uint32_t next)
{
    return arr | (dep << 8) | (cap << 16) | (next << 24);
}

// One planned vehicle of a replayed schedule (array of kMaxVehicles per spot, arrival order).
template <typename real> struct PlanRec;
template <> struct PlanRec<float> { uint32_t hdr; float soc0; float req; };                     // 12 B
template <> struct PlanRec<double> { uint32_t hdr; uint32_t pad; double soc0; double req; };    // 24 B

// Per-env scalars.  t_ep = t | episode << 8.
template <typename real> struct EnvSt;
template <> struct __align__(16) EnvSt<float> { float soc_b, pv_shift, ep_ret; uint32_t t_ep; };          // 16 B
template <> struct __align__(16) EnvSt<double> { double soc_b, pv_shift, ep_ret; uint32_t t_ep, pad; };   // 32 B

enum : int { PEN_NONE = 0, PEN_ON_DEPARTURE = 1, PEN_SPARSE = 2, PEN_DENSE = 3 };
enum : int { MODE_SAMPLE = 0, MODE_REPLAY = 1 };
enum : uint32_t { FLAG_NEG_DEMAND = 1u, FLAG_BATT_SOC_GT1 = 2u, FLAG_NAN_ACTION = 4u };
enum : int { D_TOTAL_CH = 0, D_TOTAL_DIS, D_SOLAR, D_BATT_POWER, D_GRID_POWER, D_GRID_COST, D_PEN_VEH, D_PEN_BATT, D_COUNT };
constexpr uint32_t kCtrFirstArrival = 0xFFFFFFFEu, kCtrPvShift = 0xFFFFFFFFu;

template <typename real> struct Params {
    long long n_envs;
    unsigned long long gid0;
    uint32_t seed_lo, seed_hi;
    int N, T, H, A, D;
    int pv, batt, v2x, pen_mode, diff_cap, req_soc, default_cap, auto_reset, mode;
    int i4, i10, i1;  // int(4/dt), int(10/dt), int(1/dt): charging_station.py:271-279
    int off_soc, off_dep, off_batt;
    int max_togo;     // penalty-check window: 0 none, 1 on_departure, 3 sparse, 1<<20 dense
    int pv_days;      // 1 (the reference), or D > 1: episode k reads the PV tables at offset (k % D) * T (ND == 0 kernels only)
    int has_req;      // 0: every vehicle requests SoC 1.0 (sampling without enable_requested_state_of_charge): the
                      //    requested-SoC plane is neither read nor written
    real dt, ev_pmax, ev_eff, b_cap, b_pmax, b_eff, b_dod, b_soc0, sell, cost_w, batt_w, margin;
    real dt_cap, cap_dt;   // dt / b_cap and b_cap / dt (float32 build: the battery update multiplies instead of dividing)
    // shared read-only tables in global memory (L1-resident; every env of a lock-stepped batch reads the same entry)
    const real *pv_power, *irr_norm, *price, *price_norm;  // [table_len]
    const float *dep_norm;                                 // [kSmemTab]: float(k / 24.0), k < kDepTab | gap thresholds (uint32 bits)
    // caller-owned buffers
    long long plane;       // words between the planes of `spot` (= ceil(E_bound / 32) * N * 32 of the WHOLE bound array: a
                           // slice of envs keeps it)
    const real *actions;   // [E][A]
    float *obs;            // [E][D]
    real *reward;          // [E]
    uint8_t *done;         // [E]
    float *tobs;           // [E][D] or null
    typename WordOf<real>::type *spot;  // [3][E/32][N][32]: header word | SoC column the next step starts from | requested SoC
    EnvSt<real> *envst;    // [E]
    const PlanRec<real> *plan;  // [E][N][kMaxVehicles] or null
    uint32_t *err;
    real *diag;
    real *last_ret;
    real *spot_power;      // [E][N] or null: per-spot power of the step (diagnostics)
};

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

// ------------------------------------------------------------------------------------------
// Schedule sampler.  The reference draws one Bernoulli(0.4) arrival trial per free spot and step
// (`round(rand() - 0.1) == 1`, charging_station.py:213-214) and, on arrival, SoC / requested SoC /
// capacity / departure (:218-237, 257-279); at the departure step itself no trial is drawn
// (:239-251).  The same process, sampled per VEHICLE instead of per trial: the number of failed
// trials before the next arrival is geometric, so one Philox block keyed by
// (seed, global env, spot, episode, arrival step) yields the vehicle AND the step at which its
// successor arrives (dep + 1 + gap).  The first arrival of a day is gap(block keyed kCtrFirstArrival).
// Bit-exact CPU mirror: oracle/nanogrid_oracle.c ngo_sample_episode.
// ------------------------------------------------------------------------------------------
// #{k >= 1 : x < th_k}, th_1 = floor(0.6 * 2^32), th_{k+1} = floor(th_k * 0x9999999A / 2^32)
__host__ __device__ __forceinline__ uint32_t geometric_gap(uint32_t x)
{
    uint32_t g = 0, th = 0x99999999u;
    while (x < th) {
        ++g;
        th = (uint32_t)(((unsigned long long)th * 0x9999999Aull) >> 32);
    }
    return g;
}

template <typename real> struct Vehicle { uint32_t hdr; real soc0, req; };

// The same count without the data-dependent loop (whose trip count diverges across the lanes of a warp: ~9 iterations
// per warp for a mean of 1.5 per lane): estimate k from log2(x), then settle it against the exact thresholds
// th_1 > th_2 > ... > th_43 = 0 (the CTA's shared-memory copy, staged next to the departure table).  The estimate is
// within one of the answer, so each fix-up loop runs at most once; the result is exact by construction.
__device__ __forceinline__ uint32_t geometric_gap_tab(uint32_t x, uint32_t tab_base)
{
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"((float)x));                  // x = 0: -inf -> clamped below
    int c = (int)fminf(fmaxf((32.0f - lg) * 1.35692f, 0.0f), 42.0f);              // 1 / log2(1 / 0.6)
    const uint32_t gap_base = tab_base + 4u * (uint32_t)kDepTab;
    auto th = [&](int k) {
        uint32_t v;
        asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(gap_base + 4u * (uint32_t)k));
        return v;
    };
    // the estimate is within one of the answer (verified on every threshold +- 2 and random words by
    // test_arrival_gap_table_form_equals_the_recurrence): one branch-free correction either way.  th(43..47) = 0, so
    // c + 1 needs no bound; th(0) is not a threshold
    const uint32_t up = th(c + 1), dn = th(c);
    c += (x < up) ? 1 : 0;
    c -= (c > 0 && x >= dn) ? 1 : 0;          // x >= th(c) excludes x < th(c + 1): at most one of the two corrections applies
    return (uint32_t)c;
}
template <bool SMEM> __device__ __forceinline__ uint32_t arrival_gap(uint32_t x, uint32_t tab_base)
{
    return SMEM ? geometric_gap_tab(x, tab_base) : geometric_gap(x);
}

template <typename real, bool SMEM = false>
__device__ __forceinline__ Vehicle<real> sample_vehicle(const Params<real> &p, unsigned long long stream,
                                                        uint32_t episode, int tn, uint32_t tab_base = 0u)
{
    uint32_t x[4];
    philox4x32_10((uint32_t)stream, (uint32_t)(stream >> 32), episode, (uint32_t)tn, p.seed_lo, p.seed_hi, x);
    const float u1 = __fmul_rn((float)(x[1] >> 8), 5.9604644775390625e-08f);
    const float soc0 = __fmaf_rn(0.8f, u1, 0.1f);                        // uniform(0.1, 0.9), :257-259
    float rq = 1.0f;
    if (p.req_soc) {                                                     // :227-229, 261-265
        const float u2 = __fmul_rn((float)(x[2] >> 8), 5.9604644775390625e-08f);
        const float lo = (soc0 <= 0.9f) ? __fadd_rn(soc0, 0.1f) : 1.0f;
        rq = __fmaf_rn(__fsub_rn(1.0f, lo), u2, lo);
    }
    const uint32_t cap = p.diff_cap ? 15u + (((x[3] >> 16) * 105u) >> 16) : (uint32_t)p.default_cap;  // :267-269
    const int low = tn + p.i4;                                           // :271-279
    const int up = min(tn + p.i10, p.T + p.i1);
    const int dep = (low >= up) ? low : low + (int)(((x[3] & 0xFFFFu) * (uint32_t)(up - low)) >> 16);
    const uint32_t next = (uint32_t)dep + 1u + arrival_gap<SMEM>(x[0], tab_base);
    Vehicle<real> v;
    v.hdr = make_hdr((uint32_t)tn, (uint32_t)dep, cap, next < (uint32_t)p.T ? next : kNoVehicle);
    v.soc0 = (real)soc0;
    v.req = (real)rq;
    return v;
}

// random.randint(0, 180) / 100 (envs/smart_nanogrid_environment.py:181,349)
template <typename real>
__device__ __forceinline__ real sample_pv_shift(const Params<real> &p, int N, unsigned long long gid, uint32_t episode)
{
    const unsigned long long stream = gid * (unsigned long long)N;
    uint32_t x[4];
    philox4x32_10((uint32_t)stream, (uint32_t)(stream >> 32), episode, kCtrPvShift, p.seed_lo, p.seed_hi, x);
    const uint32_t k = __umulhi(x[0], 181u);
    return (real)__fdiv_rn((float)k, 100.0f);
}

// Step at which the first vehicle of the day arrives at spot i (0xFF: none).
template <typename real, bool SMEM = false>
__device__ __forceinline__ uint32_t first_arrival(const Params<real> &p, int N, long long e, int i, uint32_t episode, uint32_t tab_base = 0u)
{
    if (p.mode == MODE_SAMPLE) {
        const unsigned long long stream = (p.gid0 + (unsigned long long)e) * (unsigned long long)N + (unsigned)i;
        uint32_t x[4];
        philox4x32_10((uint32_t)stream, (uint32_t)(stream >> 32), episode, kCtrFirstArrival, p.seed_lo, p.seed_hi, x);
        const uint32_t g = arrival_gap<SMEM>(x[0], tab_base);
        return g < (uint32_t)p.T ? g : kNoVehicle;
    }
    if (p.plan == nullptr) return kNoVehicle;
    return p.plan[((size_t)e * N + i) * kMaxVehicles].hdr & 0xFFu;   // slot 0 = first vehicle of the day
}

// The vehicle that arrives at spot i at step tn: sampled, or looked up in the replayed plan.
template <typename real, bool SMEM = false>
__device__ __forceinline__ Vehicle<real> fetch_vehicle(const Params<real> &p, int N, long long e, int i,
                                                       uint32_t episode, int tn, uint32_t tab_base = 0u)
{
    if (p.mode == MODE_SAMPLE) {
        const unsigned long long stream = (p.gid0 + (unsigned long long)e) * (unsigned long long)N + (unsigned)i;
        return sample_vehicle<real, SMEM>(p, stream, episode, tn, tab_base);
    }
    Vehicle<real> v;
    v.hdr = make_hdr(kNoVehicle, 0, 0, kNoVehicle);
    v.soc0 = 0;
    v.req = 0;
    const PlanRec<real> *pl = p.plan + ((size_t)e * N + i) * kMaxVehicles;
    for (int k = 0; k < kMaxVehicles; ++k) {
        const PlanRec<real> c = pl[k];
        if ((c.hdr & 0xFFu) == (uint32_t)tn) {
            v.hdr = c.hdr; v.soc0 = c.soc0; v.req = c.req;
            break;
        }
    }
    return v;
}

// Install vehicle `v` at the spot whose plane-0 word is *sp (planes are `plane` words apart).
template <typename real>
__device__ __forceinline__ void store_vehicle(typename WordOf<real>::type *sp, size_t plane, const Vehicle<real> &v, bool has_req)
{
    sp[PL_HDR * plane] = v.hdr;
    if (has_req) sp[PL_REQ * plane] = real_to_word(v.req);
    sp[PL_SOC * plane] = real_to_word(v.soc0);   // the step at `arr` starts from the arrival SoC (charger.py:62-67)
}

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

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

// (power * dt) / capacity: IEEE division in the float64 validation build; in the float32 production
// build one MUFU.RCP (rcp.approx, <= 1 ulp; the capacity is an integer in 1..255) and a multiply.
__device__ __forceinline__ float div_cap(float x, float cap)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(cap));
    return x * r;
}
__device__ __forceinline__ double div_cap(double x, double cap) { return x / cap; }

// Departure-time normalisation float(k / 24.0) (…environment.py:208,228): a 256-entry table, read from
// the CTA's shared-memory copy in the step kernels (SMEM) or from global memory elsewhere.
__device__ __forceinline__ float *dep_table_smem()
{
    __shared__ float tab[kSmemTab];
    return tab;
}
// 32-bit shared-space address of the table, made opaque so that it stays in one register instead of
// being re-derived (S2R + LEA) at every use.
__device__ __forceinline__ uint32_t dep_table_base()
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(dep_table_smem());
    asm volatile("" : "+r"(a));
    return a;
}
template <bool SMEM, typename real> __device__ __forceinline__ float dep_lookup(const Params<real> &p, uint32_t base, int k)
{
    if (SMEM) {
        float v;
        asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + 4u * (uint32_t)k));
        return v;
    }
    return __ldg(p.dep_norm + k);
}

// Observation offsets: ND = number of disturbance entries = (1 + pv) * (1 + H) when known at compile
// time (0 = read the offsets from the parameters).
template <int NCT, int ND> struct Offsets {
    template <typename real> static __device__ __forceinline__ int soc(const Params<real> &p) { return (NCT && ND) ? ND : p.off_soc; }
    template <typename real> static __device__ __forceinline__ int dep(const Params<real> &p) { return (NCT && ND) ? ND + NCT : p.off_dep; }
    template <typename real> static __device__ __forceinline__ int batt(const Params<real> &p) { return (NCT && ND) ? ND + 2 * NCT : p.off_batt; }
};

// Offset of the episode's day in the PV tables.  The reference always reads row 0 of solar_irradiance_2
// (pv_system_manager.py:81-91); with pv_days = D > 1 episode k reads day k % D of the (D + 1)-day series.  Only the
// instantiations with runtime observation offsets (ND == 0) support it: the dispatch keeps such configurations off
// the kernels specialised for the reference's observation shape.
template <int ND, typename real> __device__ __forceinline__ int pv_day_offset(const Params<real> &p, uint32_t episode)
{
    return (ND == 0 && p.pv_days > 1) ? (int)(episode % (uint32_t)p.pv_days) * p.T : 0;
}

// Env-level part of the observation (envs/smart_nanogrid_environment.py:197-205,
// central_management_system.py:53-60): disturbances now and `H` steps ahead, battery SoC.
// ND == 8 is the reference's own shape (PV on, NUMBER_OF_HOURS_AHEAD = 3).
template <typename real, int NCT, int ND, bool FIXED = false>
__device__ __forceinline__ void write_obs_env(const Params<real> &p, float *obs, int t, real shift, real soc_b, int pvo = 0)
{
    if (NCT && ND == 8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) obs[j == 0 ? 0 : 1 + j] = (float)(__ldg(p.irr_norm + t + j) * shift);
#pragma unroll
        for (int j = 0; j < 4; ++j) obs[j == 0 ? 1 : 4 + j] = (float)__ldg(p.price_norm + t + j);
    } else {
        int k = 0;
        if (p.pv) {
            obs[k++] = (float)(__ldg(p.irr_norm + pvo + t) * shift);
            obs[k++] = (float)__ldg(p.price_norm + t);
            for (int j = 1; j <= p.H; ++j) obs[k++] = (float)(__ldg(p.irr_norm + pvo + t + j) * shift);
            for (int j = 1; j <= p.H; ++j) obs[k++] = (float)__ldg(p.price_norm + t + j);
        } else {
            obs[k++] = (float)__ldg(p.price_norm + t);
            for (int j = 1; j <= p.H; ++j) obs[k++] = (float)__ldg(p.price_norm + t + j);
        }
    }
    if (FIXED || p.batt) obs[Offsets<NCT, ND>::batt(p)] = (float)soc_b;
}

// Begin an episode at t = 0 (SmartNanogridEnv.reset, envs/smart_nanogrid_environment.py:311-351):
// per spot, schedule the first arrival of the day (and admit it when it is at step 0), clear the SoC
// state (clear_initialisation_variables, charging_station.py:138-150) and write the reset observation.
// `spot` points at (this env, spot 0, plane 0).
// With L lanes per env, lane `sub` handles the spots sub, sub + L, ... (`spot` points at its first one)
// and lane 0 writes the env-level entries.
template <typename real, int NCT, int ND, bool SMEM, int L = 1, bool FIXED = false>
__device__ __forceinline__ void begin_episode(const Params<real> &p, int N, long long e,
                                              typename WordOf<real>::type *spot, uint32_t episode, real shift,
                                              real soc_b, float *obs, int sub = 0)
{
    const int off_soc = Offsets<NCT, ND>::soc(p), off_dep = Offsets<NCT, ND>::dep(p);
    const uint32_t dep_base = SMEM ? dep_table_base() : 0u;
#pragma unroll 1
    for (int i = sub; i < N; i += L) {
        const uint32_t next = first_arrival<real, SMEM>(p, N, e, i, episode, dep_base);
        Vehicle<real> v;
        v.hdr = make_hdr(kNoVehicle, 0, 0, next);
        v.soc0 = 0;
        v.req = 0;
        if (next == 0u) v = fetch_vehicle<real, SMEM>(p, N, e, i, episode, 0, dep_base);
        // the dense SoC array holds the arrival SoC at slot `arr` (charging_station.py:257-259),
        // so the reset observation shows it for vehicles arriving at t = 0
        store_vehicle<real>(spot + (size_t)(i / L) * (L * kBlock), (size_t)p.plane, v, !FIXED && p.has_req != 0);
        const bool present = (v.hdr & 0xFFu) == 0u;
        obs[off_soc + i] = present ? (float)v.soc0 : 0.0f;
        obs[off_dep + i] = present ? dep_lookup<SMEM>(p, dep_base, (int)((v.hdr >> 8) & 0xFFu)) : 0.0f;
    }
    if (L == 1 || sub == 0) write_obs_env<real, NCT, ND, FIXED>(p, obs, 0, shift, soc_b, pv_day_offset<ND>(p, episode));   // battery SoC survives resets (quirk Q8)
}

// Spots whose state loads are issued together (and, in the pipelined kernel, one block ahead).
template <int NCT> struct Chunk {
    static constexpr int value = NCT == 0 ? 1 : (NCT <= 16 ? NCT : 16);
    static_assert(NCT == 0 || NCT % value == 0, "the chunk must divide the number of spots");
};

// Registers holding the per-env scalars and the first chunk of per-spot state words of one env.
template <typename real, int NCT> struct StateRegs {
    typedef typename WordOf<real>::type word;
    EnvSt<real> es;
    word h[Chunk<NCT>::value], r[Chunk<NCT>::value], s[Chunk<NCT>::value];
};

// L lanes share an env (L = 1, or 4 / 2 for large stations): lane `sub` owns the spots sub, sub + L, ...;
// `spot` points at its first one and its k-th ("virtual") spot is L * kBlock words further.
template <typename real, int CH, int L>
__device__ __forceinline__ void load_spots(const typename WordOf<real>::type *spot, size_t plane, int c, bool has_req,
                                           typename WordOf<real>::type (&h)[CH], typename WordOf<real>::type (&r)[CH],
                                           typename WordOf<real>::type (&s)[CH])
{
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        const typename WordOf<real>::type *sp = spot + (size_t)(c + j) * (L * kBlock);
        h[j] = sp[PL_HDR * plane];
        r[j] = has_req ? sp[PL_REQ * plane] : real_to_word((typename std::conditional<sizeof(typename WordOf<real>::type) == 4, float, double>::type)1);
        s[j] = sp[PL_SOC * plane];
    }
}

// Issue the loads of one env's scalars and first state chunk (consumed by env_step, possibly one block later).
// (MCT = spots per lane at compile time = N / L.)
template <typename real, int MCT, int L = 1, bool FIXED = false>
__device__ __forceinline__ void load_state(const Params<real> &p, long long e, const typename WordOf<real>::type *spot,
                                           StateRegs<real, MCT> &st)
{
    st.es = p.envst[e];
    load_spots<real, Chunk<MCT>::value, L>(spot, (size_t)p.plane, 0, !FIXED && p.has_req != 0, st.h, st.r, st.s);
}

// Discharging an EV (V2X) -- Charger.discharge_vehicle, charger.py:108-140.  Kept out of line: the
// default action space has no negative charger actions, so this is cold code in the hot loop.
template <typename real> struct PowerSoc { real P, soc; };
template <typename real>
__device__ __noinline__ PowerSoc<real> discharge_vehicle(real power, real dt, real s_prev, real cap)
{
    const real calc = s_prev + div_cap(power * dt, cap);
    PowerSoc<real> r;
    // flag = ceil(0.5 * (1 + sign(calc))) == 1 iff calc >= 0 (quirk Q1)
    r.P = (calc >= (real)0) ? -((s_prev * cap) / dt) : power;
    r.soc = (calc > (real)0) ? calc : (real)0;
    return r;
}

// ------------------------------------------------------------------------------------------
// The step of ONE environment, executed by one thread.
//   e       local env index          spot    (this env, spot 0, plane 0) in the blocked state array
//   act     [A] action row           obs     [D] observation row   (both in shared memory in the step kernels)
//   SMEM    the departure table is read from the CTA's shared-memory copy (step kernels)
//   reward_out, done_out  [E] outputs of this step
//   st      the env scalars and first state chunk, loaded by load_state()
// NCT: number of spots at compile time (0 = runtime p.N); ND: see Offsets.  EXACT (double only):
// reproduces numpy's summation order of the station power sums.
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// How a thread reaches its action row and its observation row.
//   RowIO   the whole rows of the warp's 32 envs sit in shared memory; everything is a plain
//           shared-memory access.  (A variant that moved the rows of large stations through [32 envs][16
//           spots] tiles to save shared memory was measured 30 % slower at N = 64 -- its 64-byte row
//           pieces are partial-sector writes -- and was dropped; the hooks remain.)
// ------------------------------------------------------------------------------------------
template <typename real, int L = 1> struct RowIO {
    const real *act;   // this env's action row, advanced by `sub`: the lane's k-th spot is act[k * L]
    const real *act0;  // this env's action row (battery action, cold paths)
    float *obs;        // [D] this env's observation row
    int off_soc, off_dep;   // offsets of the lane's first spot (advanced by `sub`)
    __device__ __forceinline__ void begin_chunk(int) const {}
    __device__ __forceinline__ void end_chunk(int) const {}
    __device__ __forceinline__ real action(int c, int j) const { return act[(c + j) * L]; }
    __device__ __forceinline__ real action_at(int i) const { return act0[i]; }     // i: real spot / action index
    __device__ __forceinline__ void put_spot(int c, int j, float soc, float dep) const
    {
        obs[off_soc + (c + j) * L] = soc;
        obs[off_dep + (c + j) * L] = dep;
    }
    __device__ __forceinline__ void fix_soc(int k, float soc) const { obs[off_soc + k * L] = soc; }   // k: the lane's k-th spot
    __device__ __forceinline__ float *row() const { return obs; }
    __device__ __forceinline__ float read_back(int k) const { return obs[k]; }
};

// What a thread hands to the warp-cooperative admission of arriving vehicles (admit_arrivals_warp).
struct Arrivals {
    uint32_t mask;      // spots of this env whose next vehicle arrives at tn (0 on the last step of the day)
    uint32_t episode;
    int tn;
};

// L > 1 (specialised float32 kernels of large stations): L lanes of the warp share the env, lane `sub`
// (= lane / (32 / L)) owns the spots sub, sub + L, ...; `spot` and the RowIO are advanced to its first
// spot, the partial station sums are combined with shuffles, every lane computes the env-level phase
// and lane 0 alone writes its results.  All 32 lanes of the warp must be in env_step together then.
// FIXED: the reference's default station besides the observation shape -- battery on, every vehicle requests SoC 1.0
// (no requested-SoC plane) -- known at compile time.
template <typename real, int NCT, int ND, bool EXACT, bool SMEM, bool COOP = false, int L = 1, bool FIXED = false, typename IO = RowIO<real, L>>
__device__ __forceinline__ Arrivals env_step(const Params<real> &p, long long e, typename WordOf<real>::type *spot,
                                             const StateRegs<real, NCT / L> &st, const IO &io, real *reward_out,
                                             uint8_t *done_out, int sub = 0)
{
    float *const obs = io.row();   // env-level entries, reset observation
    static_assert(!COOP || (NCT > 0 && NCT <= 32 && L == 1), "cooperative admission needs a 32-bit arrival mask");
    static_assert(L == 1 || ((L == 2 || L == 4) && NCT > 0 && NCT % L == 0 && !EXACT), "several lanes per env: compile-time N divisible by L, float32");
    static_assert(!FIXED || (NCT > 0 && ND > 0 && !EXACT), "FIXED implies a compile-time observation shape");
    const bool has_req = !FIXED && p.has_req != 0;
    typedef typename WordOf<real>::type word;
    constexpr int MCT = NCT / L;                               // spots per lane at compile time
    const int N = NCT ? NCT : p.N;
    const int M = NCT ? MCT : p.N;                             // spots this lane walks
    constexpr int CH = Chunk<MCT>::value;
    constexpr int SP = L * kBlock;                             // words between consecutive spots of this lane
    const size_t plane = (size_t)p.plane;                      // words between the state planes
    const bool lead = (L == 1) || sub == 0;                    // writes the env-level results
    const int off_soc = Offsets<NCT, ND>::soc(p), off_dep = Offsets<NCT, ND>::dep(p);
    EnvSt<real> es = st.es;
    const int t = (int)(es.t_ep & 0xFFu);
    uint32_t episode = es.t_ep >> 8;
    const int tn = t + 1;
    const bool is_done = (tn == p.T);
    const uint32_t tn_key = is_done ? 0x100u : (uint32_t)tn;   // never equals a header's `next` byte when done
    // Station sums are kept per spot CLASS (spot index mod 2; mod 4 for stations of more than 32 spots) and the class
    // sums are combined at the end as (c0 + c2) + (c1 + c3): a fixed association order that a mapping with several
    // lanes per env (a lane owns the spots of one or two classes) reproduces bit for bit.
    constexpr int NCLS = NCT > 32 ? 4 : 2;                     // specialised kernels; the generic kernel decides at run time
    constexpr int NLC = NCT ? NCLS / L : 4;                    // classes this lane accumulates (its k-th spot: class k % NLC)
    static_assert(NLC >= 1 && (NCT == 0 || NLC * L == NCLS), "lanes per env must divide the number of spot classes");
    const int cls_mask = NCT ? NCLS - 1 : (p.N > 32 ? 3 : 1);
    real pos_l[NLC], neg_l[NLC], pen_l[NLC];
#pragma unroll
    for (int k = 0; k < NLC; ++k) { pos_l[k] = 0; neg_l[k] = 0; pen_l[k] = 0; }
    // a[class of the lane's i-th spot] += v  (lc: that class when known at compile time; adding +0 changes nothing)
    auto accumulate = [&](real (&a)[NLC], int lc, int i, real v) {
        if (NCT) {
#pragma unroll
            for (int k = 0; k < NLC; ++k)
                if (k == lc) a[k] += v;
        } else {
#pragma unroll
            for (int k = 0; k < NLC; ++k) a[k] += ((i & cls_mask) == k) ? v : (real)0;
        }
    };
    uint32_t err = 0;
    // spots whose next vehicle arrives at tn (specialised kernels only: N <= 64)
    typename std::conditional<(MCT > 32), unsigned long long, uint32_t>::type arrivals = 0, discharging = 0;   // bit k: the lane's k-th spot
    bool special = false;   // some occupied spot of this lane has a negative or NaN action (finished by the cold pass)
    constexpr bool DEFER = NCT > 0 && !EXACT;   // specialised kernels: V2X discharges are finished after the (branch-free) hot loop
    double cpos[EXACT ? 256 : 1], cneg[EXACT ? 256 : 1];
    int npos = 0, nneg = 0;
    // float32 build: fold the constant factors of the power (a * 22 * 0.95) and of the SoC change
    // (power * dt) into one multiplier each (<= 1 ulp from the reference's operation order);
    // the float64 validation build keeps the reference's order
    const real kw = EXACT ? p.ev_pmax : p.ev_pmax * p.ev_eff;
    const real kwh = kw * p.dt;
    const uint32_t dep_base = SMEM ? dep_table_base() : 0u;
    real nan_probe = 0;   // float32 build: sum of a * 0 over all actions is NaN iff some action is NaN or infinite

    // ---- per-spot phase: ChargingStation.simulate_vehicle_charging (charging_station.py:281-300),
    //      Charger.charge_or_discharge_vehicle (charger.py:37-140) and the lagged undercharge
    //      penalty (penaliser.py:39-87) ----
#pragma unroll 1
    for (int c = 0; c < M; c += CH) {
        io.begin_chunk(c);
        word wh[CH], wr[CH], ws[CH];
        if (c == 0) {
#pragma unroll
            for (int j = 0; j < CH; ++j) { wh[j] = st.h[j]; wr[j] = st.r[j]; ws[j] = st.s[j]; }
        } else {
            load_spots<real, CH, L>(spot, plane, c, has_req, wh, wr, ws);
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int i = c + j;                               // the lane's i-th spot = real spot i * L + sub
            const int lc = NCT ? j % NLC : 0;                  // class of this spot among the lane's own (chunk sizes are
                                                               // multiples of NLC, so j decides); generic kernel: see accumulate
            const uint32_t hd = (uint32_t)wh[j];
            const real rq = word_to_real(wr[j], (real)0);
            const real s_prev = word_to_real(ws[j], (real)0);  // SoC column t-1 (the arrival SoC when arr == t, charger.py:62-67)
            const int arr = (int)(hd & 0xFFu), dep = (int)((hd >> 8) & 0xFFu);   // arr == 0xFF: no vehicle yet
            const real a = io.action(c, j);
            if (EXACT) {
                if (a != a) err |= FLAG_NAN_ACTION;
            } else {
                nan_probe = fma(a, (real)0, nan_probe);
            }

            // check set computed by the previous observe() at t_obs = t-1; column t-1 of soc / req
            const bool checked = arr < t && t <= dep && dep - t < p.max_togo;   // arr <= t-1 < dep, dep-(t-1) <= window
            const real lower = p.margin * rq;                  // penaliser.py:72
            if (checked && s_prev < rq - lower) {              // :78
                const real d = (rq - s_prev) * (real)10;
                const real dd = mul_rn(d, d);                  // :79 (Python `** 2`); never contracted into the sum, so that every
                                                               // kernel variant rounds the square the same way
                if (EXACT) pen_l[0] = pen_l[0] + dd;           // the float64 build sums in spot order like the reference
                else accumulate(pen_l, lc, i, dd);
            }

            const bool present = arr <= t && t < dep;          // charger.occupancy[t] == 1
            real P = 0, s_new = 0;
            if (DEFER) {
                // branch-free hot loop: a negative (V2X) or NaN action is computed as a zero action here -- the spot
                // keeps its SoC exactly (s + 0 * kwh / cap == s <= 1) and contributes no power -- and is finished by
                // the cold pass below, which the env enters only when one of its occupied spots has such an action
                const real ae = fmax(a, (real)0);              // NaN -> 0
                const real cap = (real)((hd >> 16) & 0xFFu);
                const real calc = s_prev + div_cap(ae * kwh, cap);
                const real clamped = ((real)1 < calc) ? (real)1 : calc;
                s_new = present ? clamped : (real)0;
                P = present ? ae * kw : (real)0;
                special = special || (present && !(a >= (real)0));
                accumulate(pos_l, lc, i, P);
            } else if (a >= (real)0) {
                // a == 0: soc[t] = soc[p], power 0 (charger.py:38-45) -- the same as charging with zero power;
                // a > 0: charge_vehicle, charger.py:58-90: min(soc + P*dt/cap, 1); power is NOT reduced when clamped
                const real cap = (real)((hd >> 16) & 0xFFu);
                const real power = EXACT ? a * kw * p.ev_eff : a * kw;
                const real calc = s_prev + div_cap(EXACT ? power * p.dt : a * kwh, cap);
                const real clamped = ((real)1 < calc) ? (real)1 : calc;   // a == 0 gives calc == s_prev <= 1 exactly
                s_new = present ? clamped : (real)0;
                P = present ? power : (real)0;
                if (EXACT) {
                    if (P > 0) cpos[npos++] = (double)P;
                } else {
                    accumulate(pos_l, lc, i, P);               // P >= 0 here
                }
            } else if (present) {                              // a < 0 (or NaN): V2X discharge
                const PowerSoc<real> r = discharge_vehicle(a * p.ev_pmax * p.ev_eff, p.dt, s_prev, (real)((hd >> 16) & 0xFFu));
                P = r.P;
                s_new = r.soc;
                if (p.spot_power) p.spot_power[(size_t)e * N + (i * L + sub)] = P;
                if (EXACT) {
                    if (P < 0) cneg[nneg++] = (double)P;
                    if (P > 0) cpos[npos++] = (double)P;
                } else {
                    if (P < 0) accumulate(neg_l, lc, i, P);
                    if (P > 0) accumulate(pos_l, lc, i, P);
                }
            }
            word *sp = spot + (size_t)i * SP;
            sp[PL_SOC * plane] = real_to_word(s_new);
            io.put_spot(c, j, (float)s_new,                                    // charging_station.py:114-117
                        present ? dep_lookup<SMEM>(p, dep_base, dep - t) : 0.0f);     // :92-112, "/ 24" env:208
            if ((hd >> 24) == tn_key) {
                if (NCT) {
                    arrivals |= (decltype(arrivals))1 << i;
                } else {                                   // generic kernel: admit the arriving vehicle in place
                    store_vehicle<real>(sp, plane, fetch_vehicle<real, SMEM>(p, N, e, i * L + sub, episode, tn, dep_base), has_req);
                }
            }
        }
        io.end_chunk(c + CH);
    }
    real pos, neg;
    if (EXACT) {
        neg = (real)numpy_sum(cneg, nneg);
        pos = (real)numpy_sum(cpos, npos);
    }
    if (DEFER && special) {   // cold: find those spots again (their headers are unchanged so far)
        for (int i = 0; i < M; ++i) {
            const uint32_t hd = (uint32_t)spot[(size_t)i * SP + PL_HDR * plane];
            const bool present = (int)(hd & 0xFFu) <= t && t < (int)((hd >> 8) & 0xFFu);
            if (present && !(io.action(i, 0) >= (real)0)) discharging |= (decltype(discharging))1 << i;
        }
    }
    while (DEFER && discharging) {   // cold pass: the V2X discharges (and NaN actions) the hot loop computed as zero actions
        const int i = (MCT > 32) ? __ffsll((long long)discharging) - 1 : __ffs((int)discharging) - 1;
        discharging &= discharging - 1;
        word *sp = spot + (size_t)i * SP;
        const uint32_t hd = (uint32_t)sp[PL_HDR * plane];
        const real s_prev = word_to_real(sp[PL_SOC * plane], (real)0);   // the hot loop stored it back unchanged
        const PowerSoc<real> r = discharge_vehicle(io.action(i, 0) * p.ev_pmax * p.ev_eff, p.dt, s_prev, (real)((hd >> 16) & 0xFFu));
#pragma unroll
        for (int k = 0; k < NLC; ++k) {                   // DEFER implies NCT > 0: the class is i % NLC
            if (k == i % NLC) {
                if (r.P < 0) neg_l[k] += r.P;
                if (r.P > 0) pos_l[k] += r.P;
            }
        }
        sp[PL_SOC * plane] = real_to_word(r.soc);
        io.fix_soc(i, (float)r.soc);
        if (p.spot_power) p.spot_power[(size_t)e * N + (i * L + sub)] = r.P;
    }

    // ---- env-level phase: CentralManagementSystem.manage_nanogrid (central_management_system.py:99-113) ----
    // the four class sums in canonical order (class of a spot = its index mod 4, or mod 2 with classes 2 and 3 empty)
    real pos4[4] = {0, 0, 0, 0}, neg4[4] = {0, 0, 0, 0}, pen4[4] = {0, 0, 0, 0};
    if (L == 1) {
#pragma unroll
        for (int k = 0; k < NLC; ++k) { pos4[k] = pos_l[k]; neg4[k] = neg_l[k]; pen4[k] = pen_l[k]; }
    } else {
        // lane `sub` of the env holds the classes k * L + sub: every lane fetches every class from the lane that owns it
        constexpr uint32_t FULL = 0xffffffffu;
        constexpr int EPW = kBlock / L;
        const int el = (int)(threadIdx.x & 31) % EPW;
#pragma unroll
        for (int k = 0; k < NLC; ++k) {
#pragma unroll
            for (int q = 0; q < L; ++q) {
                pos4[k * L + q] = __shfl_sync(FULL, pos_l[k], el + q * EPW);
                neg4[k * L + q] = __shfl_sync(FULL, neg_l[k], el + q * EPW);
                pen4[k * L + q] = __shfl_sync(FULL, pen_l[k], el + q * EPW);
            }
        }
        real probe = nan_probe;
#pragma unroll
        for (int q = 1; q < L; ++q) probe += __shfl_xor_sync(FULL, nan_probe, q * EPW);
        nan_probe = probe;
    }
    // (two classes: classes 2 and 3 are +0 and none of these sums can be -0, so c0 + c1 is the same number)
    constexpr bool TWO = NCT > 0 && NCLS == 2;
    if (!EXACT) {
        pos = TWO ? pos4[0] + pos4[1] : (pos4[0] + pos4[2]) + (pos4[1] + pos4[3]);
        neg = TWO ? neg4[0] + neg4[1] : (neg4[0] + neg4[2]) + (neg4[1] + neg4[3]);
    }
    const real pen_veh = TWO ? pen4[0] + pen4[1] : (pen4[0] + pen4[2]) + (pen4[1] + pen4[3]);
    const real total_power = pos + neg;                                   // :105
    if (total_power < (real)0 && !p.v2x) err |= FLAG_NEG_DEMAND;          // reference raises, :158-159
    const int pvo = pv_day_offset<ND>(p, episode);
    const real solar = p.pv ? __ldg(p.pv_power + pvo + t) * es.pv_shift : (real)0;   // :99-103
    real rem = total_power - solar;                                       // :167
    real soc_b = es.soc_b, batt_power = 0, pen_b = 0;
    if (FIXED || p.batt) {                                                // battery_energy_storage_system.py:30-106
        const real ab = io.action_at(N);                                  // actions[-1], :88-89
        if (EXACT) {
            if (ab != ab) err |= FLAG_NAN_ACTION;
        } else {
            nan_probe = fma(ab, (real)0, nan_probe);
        }
        if (!EXACT) {
            // float32 build: one multiply-add with dt / b_cap instead of an IEEE division in each (divergent) branch
            const real power0 = ab * p.b_pmax * p.b_eff;
            const real calc = soc_b + power0 * p.dt_cap;
            const bool chg = ab > (real)0, dis = ab < (real)0;
            const real power = (dis && calc < (real)0) ? -(soc_b * p.cap_dt) : power0;   // :82-94
            const real soc_c = ((real)1 < calc) ? (real)1 : calc;                        // :46-74
            const real soc_d = (calc > (real)0) ? calc : (real)0;                        // :98
            soc_b = chg ? soc_c : (dis ? soc_d : soc_b);
            batt_power = (chg || dis) ? power : (real)0;
            rem = rem + batt_power;                                                      // :102, -((-rem) - power)
        } else if (ab > (real)0) {                                        // charge, :46-74
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
    if (!EXACT && nan_probe != nan_probe) err |= FLAG_NAN_ACTION;
    const real energy = rem * p.dt;                                       // central_management_system.py:107
    const real price = __ldg(p.price + t);
    const real cost = (energy < (real)0) ? energy * p.sell * price : energy * price;   // accountant.py:26-32
    const real total_pen = p.batt_w * pen_b + pen_veh;                    // penaliser.py:181
    const real total_cost = p.cost_w * fabs(cost) + total_pen;            // accountant.py:35
    const real reward = -total_cost;                                      // ...environment.py:183

    if (lead) write_obs_env<real, NCT, ND, FIXED>(p, obs, t, es.pv_shift, soc_b, pvo);   // obs at the pre-increment t, :173
    if (p.spot_power) {
        // diagnostics only (cold): the power of the charging / idle spots is a function of the action and of the
        // header, which is unchanged until the arrivals are admitted below; discharging spots were written above
        for (int i = 0; i < M; ++i) {
            const uint32_t hd = (uint32_t)spot[(size_t)i * SP + PL_HDR * plane];
            const real a = io.action(i, 0);
            const bool present = (int)(hd & 0xFFu) <= t && t < (int)((hd >> 8) & 0xFFu);
            if (!present || a >= (real)0)
                p.spot_power[(size_t)e * N + (i * L + sub)] = present ? (EXACT ? a * kw * p.ev_eff : a * kw) : (real)0;
        }
    }
    if (lead && p.diag) {
        real *diag = p.diag + (size_t)e * D_COUNT;
        diag[D_TOTAL_CH] = pos; diag[D_TOTAL_DIS] = neg; diag[D_SOLAR] = solar;
        diag[D_BATT_POWER] = batt_power; diag[D_GRID_POWER] = rem; diag[D_GRID_COST] = cost;
        diag[D_PEN_VEH] = pen_veh; diag[D_PEN_BATT] = pen_b;
    }

    // ---- t += 1, termination, auto-reset (...environment.py:174-181, 311-351) ----
    real ep_ret = es.ep_ret + reward;
    real shift = es.pv_shift;
    Arrivals out;
    out.mask = (COOP && !is_done) ? (uint32_t)arrivals : 0u;
    out.episode = episode;
    out.tn = tn;
    if (!is_done) {
        // admit the vehicles that arrive at tn (the observation above does not show them: quirk Q5);
        // COOP: left to admit_arrivals_warp(), which spreads the warp's arrivals evenly over its lanes
        while (!COOP && arrivals) {
            const int i = (MCT > 32) ? __ffsll((long long)arrivals) - 1 : __ffs((int)arrivals) - 1;
            arrivals &= arrivals - 1;
            store_vehicle<real>(spot + (size_t)i * SP, plane, fetch_vehicle<real, SMEM>(p, N, e, i * L + sub, episode, tn, dep_base), has_req);
        }
        es.t_ep = (episode << 8) | (uint32_t)tn;
    } else {
        if (lead && p.last_ret) p.last_ret[e] = ep_ret;
        ep_ret = 0;
        if (p.auto_reset) {
            // lanes sharing an env take this branch together: pair barriers keep the partner's row entries
            // complete before they are copied and untouched until they have been
            uint32_t pair = 0u;                                           // the L lanes of this env
            if (L > 1) {
#pragma unroll
                for (int q = 0; q < L; ++q) pair |= 1u << ((int)(threadIdx.x & 31) % (kBlock / L) + q * (kBlock / L));
            }
            if (p.tobs) {
                if (L > 1) __syncwarp(pair);
                float *tobs = p.tobs + (size_t)e * p.D;
                for (int k = sub; k < p.D; k += L) tobs[k] = io.read_back(k);
            }
            if (L > 1) __syncwarp(pair);
            episode = (episode + 1u) & 0xFFFFFFu;
            if (p.mode == MODE_SAMPLE) shift = sample_pv_shift(p, N, p.gid0 + (unsigned long long)e, episode);
            begin_episode<real, NCT, ND, SMEM, L, FIXED>(p, N, e, spot, episode, shift, soc_b, obs, sub);
        }
        es.t_ep = (episode << 8);                                         // t wraps to 0, :178
    }
    es.soc_b = soc_b;
    es.pv_shift = shift;
    es.ep_ret = ep_ret;
    if (lead) {
        p.envst[e] = es;
        reward_out[e] = reward;
        done_out[e] = is_done ? 1 : 0;
    }
    if (err && p.err) atomicOr(p.err + e, err);
    return out;
}

// Warp-cooperative admission of the vehicles arriving at the next step.  Each lane (= env) holds a
// mask of its arriving spots; a lane may have 0..N of them while the warp has ~1.25 per env, so letting
// every lane work through its own list costs max-over-lanes Philox rounds (~3.3 at N = 10).  Instead the
// arrivals of the whole 32-env block are compacted into a shared-memory queue (prefix sum over the
// lanes) and dealt out 32 at a time: ceil(total / 32) rounds (~1.9).  Any lane can admit any (env, spot)
// of the block: the vehicle is a pure function of (seed, global env, spot, episode, step) and its state
// words live at block_spot[plane * p.plane + spot * 32 + env_lane].
// All 32 lanes must call this (lanes without a valid env pass mask = 0); queue holds 32 * N entries.
template <typename real, int NCT, bool FIXED = false>
__device__ __forceinline__ void admit_arrivals_warp(const Params<real> &p, long long e0, int lane,
                                                    typename WordOf<real>::type *block_spot, const Arrivals &a,
                                                    uint16_t *queue)
{
    constexpr uint32_t FULL = 0xffffffffu;
    const uint32_t tab_base = dep_table_base();
    const int cnt = __popc(a.mask);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (total == 0) return;                       // warp-uniform
    int pos = incl - cnt;
    for (uint32_t m = a.mask; m; m &= m - 1) queue[pos++] = (uint16_t)((lane << 8) | (__ffs((int)m) - 1));
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
        const int idx = base + lane;
        const bool mine = idx < total;
        const uint32_t ev = mine ? queue[idx] : 0u;
        const int src = (int)(ev >> 8), i = (int)(ev & 0xFFu);
        const uint32_t episode = __shfl_sync(FULL, a.episode, src);
        const int tn = __shfl_sync(FULL, a.tn, src);
        if (mine)
            store_vehicle<real>(block_spot + (size_t)i * kBlock + src, (size_t)p.plane,
                                fetch_vehicle<real, true>(p, NCT, e0 + src, i, episode, tn, tab_base), !FIXED && p.has_req != 0);
    }
    __syncwarp();                                 // the queue may be refilled by the next step
}

}  // namespace sng
