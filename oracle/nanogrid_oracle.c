/* TEST INFRASTRUCTURE ONLY -- see nanogrid_oracle.h.  float64 scalar restatement of the
 * reference step; every function cites the reference file:line it follows.
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile). */
#include "nanogrid_oracle.h"

#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX3(e, i, k, N, W) ((((int64_t)(e)) * (N) + (i)) * (W) + (k))

int ngo_act_dim(const ngo_config *c) { return c->n_spots + (c->batt ? 1 : 0); }

/* envs/smart_nanogrid_environment.py:90-96 */
int ngo_obs_dim(const ngo_config *c)
{
    int observed = 1 + (c->pv ? 1 : 0);
    int states = observed + c->horizon * observed;
    return states + 2 * c->n_spots + (c->batt ? 1 : 0);
}

/* numpy/core/src/umath/loops_utils.h.src pairwise_sum (numpy 1.24 / 2.3): n < 8 is a
 * plain left-to-right loop starting at 0.0; 8 <= n <= 128 uses 8 strided accumulators.
 * Used by charger_power_values[mask].sum(), utils/charging_station.py:293-294. */
double ngo_numpy_sum(const double *a, int n)
{
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        int i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return ngo_numpy_sum(a, n2) + ngo_numpy_sum(a + n2, n - n2);
}

static int in_list(const int32_t *lst, int n, int v)
{
    for (int k = 0; k < n; k++)
        if (lst[k] == v) return 1;
    return 0;
}

/* ChargingStation.simulate (utils/charging_station.py:34-40) + SmartNanogridEnv.__get_observations
 * (envs/smart_nanogrid_environment.py:190-231) + CentralManagementSystem.observe
 * (utils/central_management_system.py:45-78) for one env. */
static void observe_one(const ngo_config *c, ngo_state *s, int64_t e, float *obs)
{
    const int N = c->n_spots, T = c->n_steps, W = T + 1;
    const int t = s->t[e];
    /* find_vehicles_for_penalty_check, charging_station.py:42-63 */
    if (!(t >= T)) {
        for (int i = 0; i < N; i++) {
            const int32_t *dep = s->dep + IDX3(e, i, 0, N, NGO_MAX_VEHICLES);
            const int nv = s->n_veh[e * N + i];
            int allowed;
            switch (c->penalty_mode) {
            case NGO_PEN_NONE: allowed = 0; break;
            case NGO_PEN_ON_DEPARTURE: allowed = in_list(dep, nv, t + 1); break; /* :79-84 */
            case NGO_PEN_SPARSE: /* :86-90, the n argument is ignored */
                allowed = in_list(dep, nv, t + 1) || in_list(dep, nv, t + 2) || in_list(dep, nv, t + 3);
                break;
            default: allowed = 1; break;
            }
            const double occupied = s->occ[IDX3(e, i, t, N, W)];
            s->check[e * N + i] = (occupied != 0.0 && allowed) ? 1 : 0;
        }
    }
    if (!obs) return;
    int k = 0;
    const double shift = s->pv_shift[e];
    const int pvo = s->pv_base ? s->pv_base[e] : 0;
    const int lo = t + 1, hi = lo + c->horizon; /* ...environment.py:191-192 */
    if (c->pv) {
        obs[k++] = (float)(c->irr_norm[pvo + t] * shift);  /* central_management_system.py:58 */
        obs[k++] = (float)(c->price_norm[t]);        /* :53 */
        for (int j = lo; j < hi; j++) obs[k++] = (float)(c->irr_norm[pvo + j] * shift); /* :59-60 */
        for (int j = lo; j < hi; j++) obs[k++] = (float)(c->price_norm[j]);       /* :54-55 */
    } else {
        obs[k++] = (float)(c->price_norm[t]);
        for (int j = lo; j < hi; j++) obs[k++] = (float)(c->price_norm[j]);
    }
    /* extract_current_state_of_charge_per_vehicle, charging_station.py:114-117 */
    for (int i = 0; i < N; i++) obs[k++] = (float)s->soc[IDX3(e, i, t, N, W)];
    /* calculate_departure_times, charging_station.py:92-112; "/ 24" ...environment.py:208 */
    for (int i = 0; i < N; i++) {
        double dep_time = 0.0;
        if (s->occ[IDX3(e, i, t, N, W)] != 0.0) {
            const int32_t *dep = s->dep + IDX3(e, i, 0, N, NGO_MAX_VEHICLES);
            const int nv = s->n_veh[e * N + i];
            for (int q = 0; q < nv; q++)
                if (t <= dep[q]) { dep_time = (double)(dep[q] - t); break; }
        }
        obs[k++] = (float)(dep_time / c->dep_norm);
    }
    if (c->batt) obs[k++] = (float)s->soc_b[e];
}

void ngo_observe(const ngo_config *c, ngo_state *s, float *obs, int n_threads)
{
    const int D = ngo_obs_dim(c);
    (void)n_threads;
#pragma omp parallel for num_threads(n_threads > 0 ? n_threads : 1) schedule(static)
    for (int64_t e = 0; e < s->n_envs; e++) observe_one(c, s, e, obs ? obs + e * D : (float *)0);
}

/* Charger.charge_or_discharge_vehicle (utils/charger.py:37-56) for an occupied spot. */
static double charger_act(const ngo_config *c, ngo_state *s, int64_t e, int i, int t, double a)
{
    const int N = c->n_spots, W = c->n_steps + 1;
    const int nv = s->n_veh[e * N + i];
    const int is_arrival = in_list(s->arr + IDX3(e, i, 0, N, NGO_MAX_VEHICLES), nv, t);
    /* "timestep - 1" at t = 0 is Python index -1 = the last slot (charger.py:45,66-67) */
    const int p = is_arrival ? t : (t == 0 ? W - 1 : t - 1);
    double *soc = s->soc + IDX3(e, i, 0, N, W);
    const double *cap = s->cap + IDX3(e, i, 0, N, W);
    if (a == 0) { /* charger.py:38-45 */
        soc[t] = soc[p];
        return 0.0;
    }
    if (a > 0) { /* charge_vehicle, charger.py:58-90; power = a*22*0.95 at :92-94 */
        const double power = a * c->ev_pmax * c->ev_eff;
        const double change = (power * c->dt) / cap[p];
        const double calc = soc[p] + change;
        soc[t] = (1.0 < calc) ? 1.0 : calc; /* min(calc, 1.0), :86 ; power NOT reduced */
        return power;
    }
    { /* discharge_vehicle, charger.py:108-140 */
        double power = a * c->ev_pmax * c->ev_eff; /* :142-144 */
        const double s_prev = soc[p], cp = cap[p];
        const double change = (power * c->dt) / cp;
        const double calc = s_prev + change;
        /* over_discharging_flag = ceil(0.5 * (1 + sign(calc))), :122 -> 1 iff calc >= 0 (quirk Q1) */
        const double sgn = (calc > 0) ? 1.0 : ((calc < 0) ? -1.0 : 0.0);
        const double flag = ceil(0.5 * (1 + sgn));
        if (flag * c->ev_pmax != 0.0) { /* :128-132 */
            const double possible = (s_prev * cp) / c->dt;
            power = -possible;
        }
        soc[t] = (calc > 0.0) ? calc : 0.0; /* max(0.0, calc), :136 */
        return power;
    }
}

/* CentralManagementSystem.manage_nanogrid (utils/central_management_system.py:84-155) */
static void step_one(const ngo_config *c, ngo_state *s, int64_t e, const double *act, float *obs,
                     double *reward, uint8_t *done, double *spot_power, double *diag)
{
    const int N = c->n_spots, T = c->n_steps, W = T + 1;
    const int t = s->t[e];
    double P[256];
    double pos[256] = {0}, neg[256] = {0};
    int npos = 0, nneg = 0;

    /* ChargingStation.simulate_vehicle_charging, charging_station.py:281-300 */
    for (int i = 0; i < N; i++) {
        const double a = act[i];
        if (a != a) s->err[e] |= NGO_ERR_NAN;
        if (s->occ[IDX3(e, i, t, N, W)] == 1.0) P[i] = charger_act(c, s, e, i, t, a);
        else P[i] = 0.0; /* reset_info_values only touches diagnostics, charger.py:146-156 */
        if (P[i] < 0) neg[nneg++] = P[i];
        if (P[i] > 0) pos[npos++] = P[i];
        if (spot_power) spot_power[i] = P[i];
    }
    const double total_dis = ngo_numpy_sum(neg, nneg); /* :293 */
    const double total_ch = ngo_numpy_sum(pos, npos);  /* :294 */

    /* Penaliser.penalise_charging_vehicles_outside_bounds, penaliser.py:39-87, on the check set
     * left by the PREVIOUS observe() (charging_station.py:35,321-322).  `timestep in arrivals`
     * compares an int with a list of lists -> always False -> column t-1 (penaliser.py:59-69). */
    double pen_veh = 0.0; /* builtin sum(): 0 + p0 + p1 ... */
    {
        const int col = (t == 0) ? W - 1 : t - 1;
        for (int v = 0; v < N; v++) {
            if (!s->check[e * N + v]) continue;
            const double sv = s->soc[IDX3(e, v, col, N, W)];
            const double rv = s->req[IDX3(e, v, col, N, W)];
            const double lower = c->margin * rv; /* :72 */
            double pen = 0.0;
            if (sv < rv - lower) pen = pow((rv - sv) * 10, 2.0); /* :78-79 */
            pen_veh = pen_veh + pen;
        }
    }

    /* central_management_system.py:99-103 */
    double solar = 0.0;
    if (c->pv) solar = c->pv_power[(s->pv_base ? s->pv_base[e] : 0) + t] * s->pv_shift[e];

    const double total_power = total_ch + total_dis; /* :105 */
    /* calculate_grid_power, :157-185 */
    if (total_power < 0 && !c->v2x) s->err[e] |= NGO_ERR_NEG_DEMAND; /* reference raises ValueError :158-159 */
    double rem = total_power - solar; /* :167 */
    double batt_power = 0.0, pen_b = 0.0;
    if (c->batt) {
        const double ab = act[N]; /* actions[-1], :88-89 */
        if (ab != ab) s->err[e] |= NGO_ERR_NAN;
        double sb = s->soc_b[e];
        if (ab == 0) { /* battery_energy_storage_system.py:31-35 */
            batt_power = 0.0;
        } else if (ab > 0) { /* charge, :46-74 */
            const double available = -rem; /* :37 */
            const double power = ab * c->b_pmax * c->b_eff;
            const double calc = sb + (power * c->dt) / c->b_cap;
            sb = (1.0 < calc) ? 1.0 : calc; /* :67 */
            batt_power = power;
            const double remaining_available = available - power; /* :70 */
            rem = -remaining_available;                           /* :72 */
        } else { /* discharge, :76-106 */
            double power = ab * c->b_pmax * c->b_eff;
            const double calc = sb + (power * c->dt) / c->b_cap;
            const double sgn = (calc > 0) ? 1.0 : ((calc < 0) ? -1.0 : 0.0);
            const double flag = 1 - ceil(0.5 * (1 + sgn)); /* :82 -> 1 iff calc < 0 */
            if (flag * c->b_pmax != 0.0) {
                const double possible = (sb * c->b_cap) / c->dt; /* :87 */
                power = -possible;
            }
            sb = (calc > 0.0) ? calc : 0.0; /* :98 */
            batt_power = power;
            rem = rem + power; /* :102 */
        }
        s->soc_b[e] = sb;
        /* penalise_battery_state_below_depth_of_discharge, penaliser.py:104-111 */
        if (sb < c->b_dod) pen_b = pow((c->b_dod - sb) * 10, 2.0);
        else if (sb <= 1.0) pen_b = 0.0;
        else s->err[e] |= NGO_ERR_BATT_SOC_GT1;
    }
    const double grid_power = rem;
    const double grid_energy = grid_power * c->dt; /* :107 */
    const double price = c->price[t];               /* accountant.py:38-39 */
    double cost;                                    /* accountant.py:26-32 */
    if (grid_energy < 0) cost = grid_energy * c->sell_coeff * price;
    else cost = grid_energy * price;
    const double total_pen = c->batt_pen_w * pen_b + 1 * pen_veh;  /* penaliser.py:181 */
    const double total_cost = c->cost_weight * fabs(cost) + total_pen; /* accountant.py:35 */
    if (reward) *reward = -total_cost; /* ...environment.py:183 */
    if (diag) {
        diag[NGO_D_TOTAL_CH] = total_ch;
        diag[NGO_D_TOTAL_DIS] = total_dis;
        diag[NGO_D_SOLAR] = solar;
        diag[NGO_D_BATT_POWER] = batt_power;
        diag[NGO_D_BATT_SOC] = c->batt ? s->soc_b[e] : 0.0;
        diag[NGO_D_GRID_POWER] = grid_power;
        diag[NGO_D_GRID_ENERGY] = grid_energy;
        diag[NGO_D_GRID_COST] = cost;
        diag[NGO_D_PEN_VEH] = pen_veh;
        diag[NGO_D_PEN_BATT] = pen_b;
        diag[NGO_D_PEN_TOTAL] = total_pen;
        diag[NGO_D_TOTAL_COST] = total_cost;
    }
    /* ...environment.py:173-181: obs at the pre-increment t, then t += 1, done check */
    observe_one(c, s, e, obs);
    int tn = t + 1;
    int is_done = (tn == T);
    if (is_done) tn = 0;
    s->t[e] = tn;
    if (done) *done = (uint8_t)is_done;
}

void ngo_step(const ngo_config *c, ngo_state *s, const double *actions, float *obs, double *reward,
              uint8_t *done, double *spot_power, double *diag, int n_threads)
{
    const int A = ngo_act_dim(c), D = ngo_obs_dim(c), N = c->n_spots;
#pragma omp parallel for num_threads(n_threads > 0 ? n_threads : 1) schedule(static)
    for (int64_t e = 0; e < s->n_envs; e++)
        step_one(c, s, e, actions + e * A, obs ? obs + e * D : (float *)0, reward ? reward + e : (double *)0,
                 done ? done + e : (uint8_t *)0, spot_power ? spot_power + e * N : (double *)0,
                 diag ? diag + e * NGO_D_COUNT : (double *)0);
}

/* solvers/RBC/rbc.py:6-29 with generic offsets: departure obs index (1+pv)(1+H)+N+i (the reference's
 * literal 8+N is PV on, H = 3), radiation obs[0], next-step radiation obs[2]; battery action 0. */
void ngo_rbc_actions(const ngo_config *c, int64_t n_envs, const float *obs, double *actions)
{
    const int N = c->n_spots, A = ngo_act_dim(c), D = ngo_obs_dim(c);
    const int off = (1 + (c->pv ? 1 : 0)) * (1 + c->horizon) + N;
    for (int64_t e = 0; e < n_envs; e++) {
        const float *o = obs + e * D;
        double *a = actions + e * A;
        for (int i = 0; i < N; i++) {
            const float d = o[off + i];
            if (d == 0) a[i] = 0;
            else if (d > 0 && d < 0.16667) a[i] = 1; /* rbc.py:14 */
            else a[i] = ((double)o[0] + (double)o[2]) / 2; /* rbc.py:26 */
        }
        if (c->batt) a[N] = 0;
    }
}

/* ---------------- Philox4x32-10 (Salmon et al., SC'11) ---------------- */
void ngo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Sampler spec shared with the CUDA kernel (DESIGN.md "Sampling"): the arrival process of
 * ChargingStation.generate_initial_vehicle_presence_per_charger (charging_station.py:200-255):
 * one Bernoulli(0.4) arrival trial per free spot and step (:213-214), none on the departure step
 * itself (:239-251).  Sampled per VEHICLE: the count of failed trials before the next arrival is
 * geometric, so the Philox block keyed by (seed, global env id, spot, episode, arrival step)
 * yields the vehicle and the step its successor arrives at (dep + 1 + gap); the block keyed
 * 0xFFFFFFFE yields the first arrival of the day. */
static uint32_t geometric_gap(uint32_t x)
{
    uint32_t g = 0, th = 0x99999999u; /* floor(0.6 * 2^32); th_{k+1} = floor(th_k * 0x9999999A / 2^32) */
    while (x < th) {
        g++;
        th = (uint32_t)(((uint64_t)th * 0x9999999Aull) >> 32);
    }
    return g;
}

void ngo_sample_episode(const ngo_config *c, ngo_state *s, int64_t e, uint64_t seed, uint64_t env_gid,
                        uint32_t episode)
{
    const int N = c->n_spots, T = c->n_steps, W = T + 1;
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const int i4 = (int)(4.0 / c->dt), i10 = (int)(10.0 / c->dt), i1 = (int)(1.0 / c->dt);
    for (int i = 0; i < N; i++) {
        double *occ = s->occ + IDX3(e, i, 0, N, W), *soc = s->soc + IDX3(e, i, 0, N, W);
        double *cap = s->cap + IDX3(e, i, 0, N, W), *req = s->req + IDX3(e, i, 0, N, W);
        int32_t *arr = s->arr + IDX3(e, i, 0, N, NGO_MAX_VEHICLES);
        int32_t *dep = s->dep + IDX3(e, i, 0, N, NGO_MAX_VEHICLES);
        memset(occ, 0, sizeof(double) * W); /* clear_initialisation_variables, :138-150 */
        memset(soc, 0, sizeof(double) * W);
        memset(cap, 0, sizeof(double) * W);
        memset(req, 0, sizeof(double) * W);
        int nv = 0;
        const uint64_t stream = env_gid * (uint64_t)N + (uint64_t)i;
        uint32_t x[4];
        {
            const uint32_t ctr[4] = {(uint32_t)stream, (uint32_t)(stream >> 32), episode, 0xFFFFFFFEu};
            ngo_philox4x32_10(ctr, key, x);
        }
        uint32_t next = geometric_gap(x[0]); /* failed trials at t = 0, 1, ... before the first arrival */
        while (next < (uint32_t)T && nv < NGO_MAX_VEHICLES) {
            const int t = (int)next;
            const uint32_t ctr[4] = {(uint32_t)stream, (uint32_t)(stream >> 32), episode, (uint32_t)t};
            ngo_philox4x32_10(ctr, key, x);
            const float u1 = (float)(x[1] >> 8) * 5.9604644775390625e-08f;
            const float soc0 = fmaf(0.8f, u1, 0.1f); /* uniform(0.1, 0.9), :257-259 */
            float rq = 1.0f;
            if (c->req_soc) { /* :227-229, 261-265 */
                const float u2 = (float)(x[2] >> 8) * 5.9604644775390625e-08f;
                const float lo = (soc0 <= 0.9f) ? soc0 + 0.1f : 1.0f;
                rq = fmaf(1.0f - lo, u2, lo);
            }
            const int cp = c->diff_cap ? 15 + (int)(((x[3] >> 16) * 105u) >> 16) : 40; /* :267-269 */
            const int low = t + i4; /* :271-279 */
            const int up = (t + i10 < T + i1) ? t + i10 : T + i1;
            const int dp = (low >= up) ? low : low + (int)(((x[3] & 0xFFFFu) * (uint32_t)(up - low)) >> 16);
            soc[t] = (double)soc0;
            for (int q = t; q < dp && q < T; q++) { /* :239-242 */
                occ[q] = 1;
                cap[q] = cp;
                req[q] = (double)rq;
            }
            arr[nv] = t;
            dep[nv] = dp;
            nv++;
            next = (uint32_t)dp + 1u + geometric_gap(x[0]); /* no trial on the departure step, :243-251 */
        }
        s->n_veh[e * N + i] = nv;
    }
    { /* random.randint(0, 180) / 100, ...environment.py:349 */
        const uint64_t stream = env_gid * (uint64_t)N;
        const uint32_t ctr[4] = {(uint32_t)stream, (uint32_t)(stream >> 32), episode, 0xFFFFFFFFu};
        uint32_t x[4];
        ngo_philox4x32_10(ctr, key, x);
        const uint32_t k = (uint32_t)(((uint64_t)x[0] * 181u) >> 32);
        s->pv_shift[e] = (double)((float)k / 100.0f);
        if (s->pv_base) s->pv_base[e] = c->pv_days > 1 ? (int32_t)(episode % (uint32_t)c->pv_days) * c->n_steps : 0;
    }
    s->t[e] = 0;
}

void ngo_sample_episode_batch(const ngo_config *c, ngo_state *s, uint64_t seed, uint64_t env_gid0,
                              const uint32_t *episode, const uint8_t *mask, int n_threads)
{
#pragma omp parallel for num_threads(n_threads > 0 ? n_threads : 1) schedule(static)
    for (int64_t e = 0; e < s->n_envs; e++)
        if (!mask || mask[e]) ngo_sample_episode(c, s, e, seed, env_gid0 + (uint64_t)e, episode[e]);
}
