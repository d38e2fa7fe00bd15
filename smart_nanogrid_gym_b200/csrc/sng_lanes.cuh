// sng_lanes.cuh -- the step with ONE LANE PER CHARGING SPOT: the kernel of latency-bound batches.
//
// step_simple_kernel maps one thread to one env: every byte a warp touches is part of a full 128-byte line, which is what
// 1M-env batches need (0.97 of the HBM roofline), but one warp-step is then ~1,080 dependent-ish instructions, and a batch
// of 4,096 envs is 128 warps on 148 SMs -- the step takes as long as that one chain (3.3 us), whatever the memory system
// could do.  This kernel is BASELINE.json's own mapping ("one warp per env, spots mapped to lanes, station power sum by
// warp-shuffle reductions"), used where it wins: 16 lanes per env (a warp = two envs), lane `sub` owns spot `sub`, the
// env-level phase is computed redundantly by all 16 lanes, arriving vehicles are drawn by the lane that owns the spot
// (one Philox block, all spots at once), and the observation entries leave from the lanes that computed them.  In a
// rollout (MULTI) the state stays in registers between the steps.
//
// Results are BIT-IDENTICAL to step_simple_kernel's (tests/test_gpu_parity.py::test_kernel_variants_are_bit_identical):
// the per-spot expressions are the same, and the station sums are formed in that kernel's association order -- two class
// sums (spot index mod 2), each accumulated in spot order starting from +0, V2X discharges added behind the charging
// powers, then c0 + c1 -- by walking the spots with shuffles instead of reducing by butterfly.
//
// Scope: float32, the reference's default station shape (PV on, 3 steps ahead, battery on), N <= 16 spots, sampled or
// replayed schedules, with or without individual requested SoCs (the third state plane, kept in a register like the other
// two).  Reference lines as in env_step (sng_device.cuh).
#pragma once
#include "sng_device.cuh"

namespace sng {

// (included by sng_engine.cuh behind pdl_wait / pdl_launch_dependents / publish_dep_table_warp)
constexpr int kLaneGroup = 16;   // lanes per env

// REQ: vehicles carry individual requested SoCs (the third state plane); false: every vehicle requests 1.0 and the plane is
// neither read nor written (compile time: the common case keeps its registers and its instruction count).
template <int NCT, bool MULTI, bool REQ>
__global__ void __launch_bounds__(128)
    step_lanes_kernel(const Params<float> p, const float *actions, float *obs_out, float *reward_out, uint8_t *done_out, int n_steps)
{
    static_assert(NCT > 0 && NCT <= kLaneGroup, "one lane per spot: at most 16 spots");
    constexpr int ND = 8, A = NCT + 1, D = ND + 2 * NCT + 1;
    constexpr uint32_t FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, sub = lane & (kLaneGroup - 1);
    const int n_envs = (int)p.n_envs;
    const int e0 = (int)((blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2);   // the warp's two envs: e0, e0 + 1
    if (e0 >= n_envs) return;                                    // warp-uniform
    const int e = e0 + (lane >> 4);
    const bool valid = e < n_envs;                               // the upper half of the last warp of an odd batch idles along
    const bool own = valid && sub < NCT;                         // this lane owns spot `sub` of env e
    publish_dep_table_warp(p);                                   // constant tables: ahead of the dependency wait
    pdl_wait();
    pdl_launch_dependents();
    const size_t plane = (size_t)p.plane;
    uint32_t *const sp = p.spot + (size_t)(e / kBlock) * (size_t)(NCT * kBlock) + (size_t)sub * kBlock + (size_t)(e % kBlock);
    // lanes without a spot carry an empty header: never present, never checked, never arriving
    constexpr bool has_req = REQ;
    uint32_t hd = make_hdr(kNoVehicle, 0, 0, kNoVehicle);
    float soc = 0.0f, rq = 1.0f;
    EnvSt<float> es = {0.0f, 0.0f, 0.0f, 0u};
    if (own) {
        hd = sp[PL_HDR * plane];
        soc = __uint_as_float(sp[PL_SOC * plane]);
        if (has_req) rq = __uint_as_float(sp[PL_REQ * plane]);
    }
    if (valid) es = p.envst[e];
    bool new_vehicle = false;                                    // header (and requested SoC) changed since they were loaded
    // the action of this lane's spot and the battery action (actions[-1], read by every lane of the env); a rollout
    // requests the next step's pair one step ahead, so that their way from L2 is off the step-to-step chain
    float a_next = own ? actions[(size_t)e * A + sub] : 0.0f;
    float ab_next = valid ? actions[(size_t)e * A + NCT] : 0.0f;
    const uint32_t dep_base = dep_table_base();
    const float kw = p.ev_pmax * p.ev_eff;
    const float kwh = kw * p.dt;

    // class sums in step_simple_kernel's order: c[k] = ((0 + v[k]) + v[k + 2]) + ... over the spots of class k
    auto class_sums = [&](float v, float &c0, float &c1) {
#pragma unroll
        for (int i = 0; i < NCT; ++i) {
            const float x = __shfl_sync(FULL, v, i, kLaneGroup);
            if (i & 1) c1 += x; else c0 += x;
        }
    };

#pragma unroll 1
    for (int s = 0; s < (MULTI ? n_steps : 1); ++s) {
        const size_t slab = MULTI ? (size_t)s * (size_t)n_envs : 0;
        float *row = obs_out + (slab + (size_t)e) * D;
        const float a = a_next, ab = ab_next;
        if (MULTI && s + 1 < n_steps) {
            const float *act = actions + (slab + (size_t)n_envs + (size_t)e) * A;
            if (own) a_next = act[sub];
            if (valid) ab_next = act[NCT];
        }

        const int t = (int)(es.t_ep & 0xFFu);
        uint32_t episode = es.t_ep >> 8;
        const int tn = t + 1;
        const bool is_done = (tn == p.T);
        const uint32_t tn_key = is_done ? 0x100u : (uint32_t)tn;
        uint32_t err = 0;

        // ---- per-spot phase (charging_station.py:281-300, charger.py:37-140, penaliser.py:39-87) ----
        const float s_prev = soc;
        const int arr = (int)(hd & 0xFFu), dep = (int)((hd >> 8) & 0xFFu);
        const bool checked = arr < t && t <= dep && dep - t < p.max_togo;
        const float lower = p.margin * rq;
        float pen = 0.0f;
        if (checked && s_prev < rq - lower) {
            const float d = (rq - s_prev) * 10.0f;
            pen = mul_rn(d, d);
        }
        const bool present = arr <= t && t < dep;
        const float ae = fmax(a, 0.0f);
        const float cap = (float)((hd >> 16) & 0xFFu);
        const float calc = s_prev + div_cap(ae * kwh, cap);
        const float clamped = (1.0f < calc) ? 1.0f : calc;
        float s_new = present ? clamped : 0.0f;
        const float P = present ? ae * kw : 0.0f;
        const bool special = present && !(a >= 0.0f);            // V2X discharge (or a NaN action)
        float P_spot = P, v_neg = 0.0f, v_pos = 0.0f;
        if (special) {
            const PowerSoc<float> r = discharge_vehicle(a * p.ev_pmax * p.ev_eff, p.dt, s_prev, cap);
            if (r.P < 0.0f) v_neg = r.P;
            if (r.P > 0.0f) v_pos = r.P;
            s_new = r.soc;
            P_spot = r.P;
        }
        float pos0 = 0.0f, pos1 = 0.0f, neg0 = 0.0f, neg1 = 0.0f, pen0 = 0.0f, pen1 = 0.0f;
        class_sums(P, pos0, pos1);
        class_sums(pen, pen0, pen1);
        if (__any_sync(FULL, special)) {                         // cold; warp-uniform
            class_sums(v_neg, neg0, neg1);
            class_sums(v_pos, pos0, pos1);
        }
        const bool bad_action = (own && !(fabsf(a) <= 3.4028235e38f)) || (valid && !(fabsf(ab) <= 3.4028235e38f));
        const uint32_t bad = __ballot_sync(FULL, bad_action);
        if ((bad >> (lane & 16)) & 0xFFFFu) err |= FLAG_NAN_ACTION;
        const float dep_obs = present ? dep_lookup<true>(p, dep_base, dep - t) : 0.0f;

        // ---- env-level phase (central_management_system.py:99-113), computed by every lane of the env ----
        const float pos = pos0 + pos1, neg = neg0 + neg1;
        const float pen_veh = pen0 + pen1;
        const float total_power = pos + neg;
        if (total_power < 0.0f && !p.v2x) err |= FLAG_NEG_DEMAND;
        const float solar = p.pv ? __ldg(p.pv_power + t) * es.pv_shift : 0.0f;
        float rem = total_power - solar;
        float soc_b = es.soc_b, batt_power = 0.0f, pen_b = 0.0f;
        {
            const float power0 = ab * p.b_pmax * p.b_eff;
            const float calc_b = soc_b + power0 * p.dt_cap;
            const bool chg = ab > 0.0f, dis = ab < 0.0f;
            const float power = (dis && calc_b < 0.0f) ? -(soc_b * p.cap_dt) : power0;
            const float soc_c = (1.0f < calc_b) ? 1.0f : calc_b;
            const float soc_d = (calc_b > 0.0f) ? calc_b : 0.0f;
            soc_b = chg ? soc_c : (dis ? soc_d : soc_b);
            batt_power = (chg || dis) ? power : 0.0f;
            rem = rem + batt_power;
            if (soc_b < p.b_dod) {
                const float d = (p.b_dod - soc_b) * 10.0f;
                pen_b = d * d;
            } else if (!(soc_b <= 1.0f)) {
                err |= FLAG_BATT_SOC_GT1;
            }
        }
        const float energy = rem * p.dt;
        const float price = __ldg(p.price + t);
        const float cost = (energy < 0.0f) ? energy * p.sell * price : energy * price;
        const float total_pen = p.batt_w * pen_b + pen_veh;
        const float total_cost = p.cost_w * fabs(cost) + total_pen;
        const float reward = -total_cost;

        // one row of observations: spot entries from the lanes that own them, the nine env-level entries from lanes 0..8
        auto put_row = [&](float *r, float soc_i, float dep_i, int tt, float shift, float sb) {
            if (own) {
                r[ND + sub] = soc_i;
                r[ND + NCT + sub] = dep_i;
            }
            if (valid) {
                if (sub < 4) r[sub == 0 ? 0 : 1 + sub] = __ldg(p.irr_norm + tt + sub) * shift;
                else if (sub < 8) r[sub == 4 ? 1 : sub] = __ldg(p.price_norm + tt + (sub - 4));
                else if (sub == 8) r[ND + 2 * NCT] = sb;
            }
        };
        if (p.spot_power && own) p.spot_power[(size_t)e * NCT + sub] = P_spot;
        if (p.diag && valid && sub == 9) {
            float *diag = p.diag + (size_t)e * D_COUNT;
            diag[D_TOTAL_CH] = pos; diag[D_TOTAL_DIS] = neg; diag[D_SOLAR] = solar;
            diag[D_BATT_POWER] = batt_power; diag[D_GRID_POWER] = rem; diag[D_GRID_COST] = cost;
            diag[D_PEN_VEH] = pen_veh; diag[D_PEN_BATT] = pen_b;
        }

        // ---- t += 1, termination, auto-reset (...environment.py:174-181, 311-351) ----
        float ep_ret = es.ep_ret + reward;
        float shift = es.pv_shift;
        soc = s_new;
        if (!is_done || !p.auto_reset) {
            put_row(row, s_new, dep_obs, t, es.pv_shift, soc_b);
            if (!is_done) {
                if ((hd >> 24) == tn_key) {                      // the vehicle arriving at tn (not shown above: quirk Q5)
                    const Vehicle<float> v = fetch_vehicle<float, true>(p, NCT, e, sub, episode, tn, dep_base);
                    hd = v.hdr;
                    soc = v.soc0;
                    if (has_req) rq = v.req;
                    new_vehicle = true;
                }
                es.t_ep = (episode << 8) | (uint32_t)tn;
            } else {
                if (p.last_ret && valid && sub == 0) p.last_ret[e] = ep_ret;
                ep_ret = 0.0f;
                es.t_ep = (episode << 8);
            }
        } else {
            if (p.last_ret && valid && sub == 0) p.last_ret[e] = ep_ret;
            ep_ret = 0.0f;
            if (p.tobs) put_row(p.tobs + (size_t)e * D, s_new, dep_obs, t, es.pv_shift, soc_b);
            episode = (episode + 1u) & 0xFFFFFFu;
            if (p.mode == MODE_SAMPLE) shift = sample_pv_shift(p, NCT, p.gid0 + (unsigned long long)e, episode);
            float soc_i = 0.0f, dep_i = 0.0f;
            if (own) {                                           // begin_episode for this lane's spot
                const uint32_t next = first_arrival<float, true>(p, NCT, e, sub, episode, dep_base);
                Vehicle<float> v;
                v.hdr = make_hdr(kNoVehicle, 0, 0, next);
                v.soc0 = 0.0f;
                v.req = 0.0f;
                if (next == 0u) v = fetch_vehicle<float, true>(p, NCT, e, sub, episode, 0, dep_base);
                hd = v.hdr;
                soc = v.soc0;
                if (has_req) rq = v.req;
                new_vehicle = true;
                const bool present0 = (v.hdr & 0xFFu) == 0u;
                soc_i = present0 ? v.soc0 : 0.0f;
                dep_i = present0 ? dep_lookup<true>(p, dep_base, (int)((v.hdr >> 8) & 0xFFu)) : 0.0f;
            }
            put_row(row, soc_i, dep_i, 0, shift, soc_b);
            es.t_ep = (episode << 8);
        }
        es.soc_b = soc_b;
        es.pv_shift = shift;
        es.ep_ret = ep_ret;
        if (valid && sub == 10) (reward_out + slab)[e] = reward;
        if (valid && sub == 11) (done_out + slab)[e] = is_done ? 1 : 0;
        if (err && p.err && valid && sub == 0) atomicOr(p.err + e, err);
    }
    // ---- the state goes back once per launch ----
    if (own) {
        if (new_vehicle) {
            sp[PL_HDR * plane] = hd;
            if (has_req) sp[PL_REQ * plane] = __float_as_uint(rq);
        }
        sp[PL_SOC * plane] = __float_as_uint(soc);
    }
    if (valid && sub == 0) p.envst[e] = es;
}

}  // namespace sng
