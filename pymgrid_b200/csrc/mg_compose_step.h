/*
 * mg_compose_step.h -- one microgrid, one timestep, ANY module composition: the per-env body of the composed-step
 * kernels (mg_compose.cu).  Restates Microgrid.run (src/pymgrid/microgrid/microgrid.py:227-325) over the module table of
 * include/pymgrid_b200_compose.h; citations are relative to /root/reference/src/pymgrid/.
 *
 * Every function is __host__ __device__ and free of CUDA-only constructs, so the CPU test-suite compiles this very file
 * with g++ (tests/hostsim/, -ffp-contract=off to match nvcc -fmad=false) and pins the arithmetic against vectors recorded
 * from the live reference without a GPU; on the GPU the same source runs one thread per env.
 *
 * IEEE f64 throughout, operations in the reference's order: sums start from 0.0 in dispatch order, np.sum's pairwise
 * order is restated for the energy lists (MgcSum), no fused multiply-add.
 */
#ifndef MG_COMPOSE_STEP_H
#define MG_COMPOSE_STEP_H

#include <math.h>
#include <stdint.h>

#include "pymgrid_b200.h"
#include "pymgrid_b200_compose.h"

#if defined(__CUDACC__)
#define MGC_HD __host__ __device__ __forceinline__
#else
#define MGC_HD static inline
#endif

/* what one env's step reads */
struct MgcView {
    const MgcModule *mod;
    int32_t n_mod;
    const double *cfg;            /* this env's config record */
    const double *series;
    const int64_t *series_off;
    const double *series_nrm;     /* may be NULL */
    int32_t T;
};

MGC_HD double mgc_inf() { return HUGE_VAL; }

/* utils/space.py:183-231 */
MGC_HD double mgc_spread(double low, double high) {
    double s = high - low;
    return s == 0.0 ? 1.0 : s;
}
MGC_HD double mgc_normalize(double v, double low, double high) { return (v - low) / mgc_spread(low, high); }
MGC_HD double mgc_denormalize(double x, double low, double high) { return low + mgc_spread(low, high) * x; }

/* numpy.isclose(a, b) with the default rtol=1e-5, atol=1e-8 */
MGC_HD bool mgc_isclose(double a, double b) { return fabs(a - b) <= (1e-8 + 1e-5 * fabs(b)); }

/*
 * numpy.sum of a GROWING list of f64, in numpy's order (pairwise_sum, numpy/_core/src/umath/loops_utils.h.src): sequential
 * from 0.0 below 8 addends; from 8 on, eight running sums over the full blocks of eight, combined pairwise, then the
 * remaining n % 8 values added one by one (valid up to 128 addends, where numpy starts to recurse; a microgrid has at most
 * MGC_MAX_MODULES = 64).  microgrid/utils/step.py:33-36 sums the provided / absorbed energy lists with it after the fixed,
 * after the controllable and after the flex modules, so the sum is needed at three lengths of the same list.  Kept instead
 * of the list: the eight running sums (entry j receives every addend whose position is j mod 8, as it arrives -- the same
 * additions in the same order as adding a completed block at once) and `seq`, numpy's result for the current length: the
 * plain left-to-right sum while the list is shorter than 8, the pairwise combination of the running sums whenever a block
 * of eight completes, plus the addends of the incomplete block one by one.  Reading the sum is then free.
 */
struct MgcSum {
    double r[8];      /* running sums; entries at or past n % 8 still lack the current block's addend */
    double seq;       /* numpy.sum of the first n addends */
    int n;
};
MGC_HD void mgc_sum_init(MgcSum &S) { S.n = 0; S.seq = 0.0; }
MGC_HD void mgc_sum_append(MgcSum &S, double v) {
    const int j = S.n & 7;
    S.r[j] = (S.n < 8) ? v : S.r[j] + v;
    S.n += 1;
    if ((S.n & 7) == 0) S.seq = ((S.r[0] + S.r[1]) + (S.r[2] + S.r[3])) + ((S.r[4] + S.r[5]) + (S.r[6] + S.r[7]));
    else S.seq += v;
}
MGC_HD double mgc_sum_value(const MgcSum &S) { return S.seq; }

/* genset_module.py:216-233 */
MGC_HD void mgc_genset_reset_times(int cs, int U, int D, int &up, int &dn) {
    if (cs) { up = 0; dn = D; }
    else { dn = 0; up = U; }
}

/* genset_module.py:235-346 (update_status, _finish_in_progress_change, _non_instantaneous_update) */
MGC_HD void mgc_genset_update_status(int &cs, int &gs, int &up, int &dn, double goal_status, int U, int D, int abortion) {
    const int goal = (int)rint(goal_status);      /* Python round(): half to even */
    if (goal == cs && cs == gs) return;
    const bool instant_up = (U == 0 && goal == 1), instant_down = (D == 0 && goal == 0);
    if (goal != gs && (abortion || instant_up || instant_down)) gs = goal;
    if (up == 0 && gs == 1) { cs = 1; mgc_genset_reset_times(cs, U, D, up, dn); return; }
    if (dn == 0 && gs == 0) { cs = 0; mgc_genset_reset_times(cs, U, D, up, dn); return; }
    if (goal == cs && cs != gs && abortion) {
        gs = goal;
        mgc_genset_reset_times(cs, U, D, up, dn);
    } else if (cs == gs && gs != goal) {
        mgc_genset_reset_times(cs, U, D, up, dn);
        gs = goal;
    }
    if (gs != cs) {
        if (gs == 0) dn -= 1;
        else up -= 1;
    }
}

MGC_HD const double *mgc_series_of(const MgcView &V, const double *p) { return V.series + V.series_off[(int)p[0]]; }

/*
 * Element k of module m's normalised observation at step t (the module's state AFTER the step, base_module.py:157):
 *   time series  [current(C), forecast_0(C) ... forecast_{H-1}(C)]: ts[t] or the fill row past the end
 *                (base_timeseries_module.py:103-140), forecast rows clipped to the bounds and short windows padded with
 *                (high + low) / 2 (forecast/forecaster.py:95, 120-149), normalised per column (utils/space.py:207-218)
 *   battery      [soc, current_charge]                           battery_module.py:323-330
 *   genset       [current, goal, steps_until_up, steps_until_down]  genset_module.py:503-509
 */
MGC_HD double mgc_obs_element(const MgcView &V, int m, int k, int t, const double *fstate, const int32_t *istate) {
    const MgcModule &M = V.mod[m];
    const double *p = V.cfg + M.param_off;
    switch (M.kind) {
    case MGC_LOAD:
    case MGC_RENEWABLE:
    case MGC_GRID: {
        const int C = (M.kind == MGC_GRID) ? 4 : 1;
        const int row = k / C, c = k - row * C;
        const double low = (M.kind == MGC_GRID) ? p[4 + c] : p[1];
        const double high = (M.kind == MGC_GRID) ? p[8 + c] : p[2];
        const int idx = t + row;
        double v;
        if (idx < V.T && t < V.T) {
            if (V.series_nrm) return V.series_nrm[V.series_off[(int)p[0]] + (int64_t)idx * C + c];     /* pre-normalised gather */
            v = mgc_series_of(V, p)[(int64_t)idx * C + c];
            if (row > 0) {
                if (v < low) v = low;
                if (v > high) v = high;
            }
        } else {
            v = (high + low) / 2;
            if (row > 0 && t < V.T) {
                if (v < low) v = low;
                if (v > high) v = high;
            }
        }
        return mgc_normalize(v, low, high);
    }
    case MGC_BATTERY: {
        const double *s = fstate + M.fstate_off;
        if (k == 0) return mgc_normalize(s[1], p[0] / p[1], 1.0);
        return mgc_normalize(s[0], p[0], p[1]);
    }
    case MGC_GENSET: {
        const int32_t *s = istate + M.istate_off;
        if (k < 2) return mgc_normalize((double)s[k], 0.0, 1.0);
        return mgc_normalize((double)s[k], 0.0, p[3 + k]);      /* k = 2: start_up_time (p[5]); k = 3: wind_down_time (p[6]) */
    }
    default:
        return 0.0;
    }
}

MGC_HD int mgc_obs_len(const MgcModule &M) {
    switch (M.kind) {
    case MGC_LOAD:
    case MGC_RENEWABLE: return 1 + M.horizon;
    case MGC_GRID: return 4 * (1 + M.horizon);
    case MGC_BATTERY: return 2;
    case MGC_GENSET: return 4;
    default: return 0;
    }
}

MGC_HD bool mgc_is_timeseries(int kind) { return kind == MGC_LOAD || kind == MGC_RENEWABLE || kind == MGC_GRID; }
/* dispatch class: 0 fixed, 1 controllable, 2 flex (module_type[1] of the reference classes) */
MGC_HD int mgc_dispatch_class(int kind) {
    return kind == MGC_LOAD ? 0 : (kind == MGC_RENEWABLE || kind == MGC_UNBALANCED) ? 2 : 1;
}
MGC_HD bool mgc_is_source(int kind) { return kind != MGC_LOAD; }
MGC_HD bool mgc_is_sink(int kind) { return kind == MGC_LOAD || kind == MGC_BATTERY || kind == MGC_GRID || kind == MGC_UNBALANCED; }

/* accumulators of one Microgrid.run: microgrid/utils/step.py MicrogridStep */
struct MgcStepAcc {
    MgcSum provided, absorbed;
    double reward;
    int done;
    uint32_t flags;
};

MGC_HD double mgc_max_production(const MgcView &V, const MgcModule &M, const double *p, int t, const double *fstate,
                                 const int32_t *istate) {
    switch (M.kind) {
    case MGC_BATTERY: {      /* battery_module.py:283-286 */
        const double charge = fstate[M.fstate_off];
        return fmin(p[3], charge - p[0]) * p[4];
    }
    case MGC_GENSET: return istate[M.istate_off] * p[1];                          /* genset_module.py:466-482 */
    case MGC_GRID: return p[1] * mgc_series_of(V, p)[(int64_t)t * 4 + 3];         /* grid_module.py:314-316   */
    case MGC_RENEWABLE: return mgc_series_of(V, p)[t];                             /* renewable_module.py:95-110 */
    default: return mgc_inf();                                                     /* unbalanced_energy_module.py:99-101 */
    }
}

MGC_HD double mgc_max_consumption(const MgcView &V, const MgcModule &M, const double *p, int t, const double *fstate) {
    switch (M.kind) {
    case MGC_BATTERY: {      /* battery_module.py:288-291 */
        const double charge = fstate[M.fstate_off];
        return fmin(p[2], p[1] - charge) / p[4];
    }
    case MGC_GRID: return p[2] * mgc_series_of(V, p)[(int64_t)t * 4 + 3];         /* grid_module.py:318-320 */
    default: return mgc_inf();
    }
}

/*
 * BaseMicrogridModule.step for module m with the UNNORMALISED request `a` (> 0 source, < 0 sink; base_module.py:161-274)
 * followed by the module's update().  The genset's status update (genset_module.py:146-148) has already happened.
 * Appends to the accumulators; writes the module's info slots when `info` is not NULL.
 */
MGC_HD void mgc_module_step(const MgcView &V, int m, double a, int t, double *fstate, const int32_t *istate, MgcStepAcc &A,
                            double *info) {
    const MgcModule &M = V.mod[m];
    const double *p = V.cfg + M.param_off;
    const int kind = M.kind;
    const bool fixed = (kind == MGC_LOAD);
    /* base_module.py:161-171: a > 0 source, a < 0 sink, otherwise source when the module can be one */
    const bool as_source = (a > 0) ? true : (a < 0) ? false : mgc_is_source(kind);
    double energy = 0.0;
    bool clipped = false;
    if (!fixed) {
        if (as_source) {                                   /* as_source, base_module.py:173-226: upper test first */
            const double mx = mgc_max_production(V, M, p, t, fstate, istate);
            const double mn = (kind == MGC_GENSET) ? istate[M.istate_off] * p[0] : 0.0;      /* genset_module.py:484-501 */
            if (a > mx) { energy = mx; clipped = true; }
            else if (a < mn) { energy = mn; clipped = true; }
            else energy = a;
        } else {                                           /* as_sink, base_module.py:228-274 */
            if (!mgc_is_sink(kind)) {                      /* GensetModule / RenewableModule.update assert as_source */
                A.flags |= MGC_FLAG_NOT_A_SINK;
                energy = 0.0;
            } else {
                const double e = -1.0 * a;
                const double mc = mgc_max_consumption(V, M, p, t, fstate);
                if (e > mc) { energy = mc; clipped = true; }
                else energy = e;
                if (!(energy >= 0)) A.flags |= MG_FLAG_NEGATIVE_ABSORB;
            }
        }
    }
    if (clipped) A.flags |= MGC_FLAG_CLIP | (M.raise_errors ? MGC_FLAG_CLIP_RAISES : 0u);
    double reward = 0.0, extra = 0.0, provided = 0.0, absorbed = 0.0;
    bool sink = !as_source;
    switch (kind) {
    case MGC_LOAD:                                         /* load_module.py:86-91 */
        absorbed = -1 * mgc_series_of(V, p)[t];
        sink = true;
        break;
    case MGC_RENEWABLE: {                                  /* renewable_module.py:86-93 */
        const double cur = mgc_series_of(V, p)[t];
        provided = energy;
        extra = cur - energy;
        sink = false;
        break;
    }
    case MGC_BATTERY: {                                    /* battery_module.py:108-130, 244-278 */
        double internal;
        if (as_source) { provided = energy; internal = (-1.0 * energy) / p[4]; }
        else { absorbed = energy; internal = energy * p[4]; }
        double charge = fstate[M.fstate_off] + internal;
        if (charge < p[0]) {
            if (!mgc_isclose(charge, p[0])) A.flags |= MG_FLAG_BATTERY_MIN_CAP;
            charge = p[0];
        }
        fstate[M.fstate_off] = charge;
        fstate[M.fstate_off + 1] = charge / p[1];
        reward = -1.0 * (fabs(internal) * p[5]);
        break;
    }
    case MGC_GENSET: {                                     /* genset_module.py:151-214 */
        const double co2 = p[3] * energy;
        const double cost = p[2] * energy + p[4] * co2;
        reward = -1.0 * cost;
        provided = energy;
        extra = co2;
        sink = false;
        break;
    }
    case MGC_GRID: {                                       /* grid_module.py:134-228 */
        const double *row = mgc_series_of(V, p) + (int64_t)t * 4;
        if (as_source) {
            const double co2 = energy * row[2];
            reward = -1 * row[0] * energy + (-1.0 * p[3] * co2);
            provided = energy;
            extra = co2;
        } else {
            reward = row[1] * energy + (-1.0 * p[3] * 0.0);
            absorbed = energy;
            extra = 0.0;
        }
        break;
    }
    default:                                               /* unbalanced_energy_module.py:28-70 */
        if (as_source) { reward = -1.0 * (p[0] * energy); provided = energy; }
        else { reward = -1.0 * (p[1] * energy); absorbed = energy; }
        break;
    }
    /* MicrogridStep.append, microgrid/utils/step.py:13-31 */
    A.reward += reward;
    if (sink) mgc_sum_append(A.absorbed, absorbed);
    else mgc_sum_append(A.provided, provided);
    if (info) {
        double *r = info + (int64_t)M.listing * MGC_INFO_SLOTS;
        r[0] = sink ? 0.0 : provided;
        r[1] = sink ? absorbed : 0.0;
        r[2] = extra;
        r[3] = reward;
        r[4] = sink ? 1.0 : 0.0;
    }
}

/*
 * Microgrid.run for one env.  `t` is the env's current step (advanced on success), `fstate` / `istate` its state rows,
 * `action` its action row, `final_step` the end of its episode window.  Writes reward / done, ORs event bits into *flags, fills `info` (n_mod * MGC_INFO_SLOTS +
 * MGC_BALANCE_SLOTS doubles) when not NULL.
 */
MGC_HD void mgc_env_step(const MgcView &V, int32_t &t, double *fstate, int32_t *istate, const double *action, int normalized,
                         int final_step, double *reward_out, uint8_t *done_out, double *info, uint32_t *flags) {
    MgcStepAcc A;
    mgc_sum_init(A.provided);
    mgc_sum_init(A.absorbed);
    A.reward = 0.0;
    A.done = 0;
    A.flags = 0;
    const int t0 = t;
    const int n = V.n_mod;
    bool any_series = false;
    for (int m = 0; m < n; ++m) any_series |= mgc_is_timeseries(V.mod[m].kind);
    if (any_series && t0 >= V.T) {       /* the reference raises IndexError reading ts[t] (e.g. load_module.py:111) */
        *reward_out = NAN;
        *done_out = 1;
        *flags |= MG_FLAG_STEP_PAST_END;
        return;
    }
    if (info)
        for (int i = 0; i < n * MGC_INFO_SLOTS + MGC_BALANCE_SLOTS; ++i) info[i] = 0.0;
    /* BaseTimeSeriesMicrogridModule._done, base_timeseries_module.py:124-125, evaluated before t += 1 */
    const int ts_done = (t0 >= final_step - 1);
    int m = 0;
    /* ---- fixed modules: step(0.0, normalized=False), microgrid.py:255-257 ---- */
    for (; m < n && mgc_dispatch_class(V.mod[m].kind) == 0; ++m) {
        mgc_module_step(V, m, 0.0, t0, fstate, istate, A, info);
        A.done |= ts_done;
    }
    const double fixed_p = mgc_sum_value(A.provided), fixed_a = mgc_sum_value(A.absorbed);
    /* ---- controllable modules with the caller's control, microgrid.py:262-275 ---- */
    for (; m < n && mgc_dispatch_class(V.mod[m].kind) == 1; ++m) {
        const MgcModule &M = V.mod[m];
        const double *p = V.cfg + M.param_off;
        const double *c = action + M.act_col;
        double a;
        if (M.kind == MGC_GENSET) {
            const double goal = c[0];                    /* never denormalised: genset_module.py:146 */
            int32_t *s = istate + M.istate_off;
            if (!(0 <= goal && goal <= 1)) A.flags |= MG_FLAG_GENSET_GOAL_RANGE;      /* :147 */
            else {
                int cs = s[0], gs = s[1], up = s[2], dn = s[3];
                mgc_genset_update_status(cs, gs, up, dn, goal, (int)p[5], (int)p[6], p[7] != 0.0);
                s[0] = cs; s[1] = gs; s[2] = up; s[3] = dn;
            }
            a = normalized ? mgc_denormalize(c[1], 0.0, p[1]) : c[1];                 /* :511-517 */
        } else if (M.kind == MGC_BATTERY) {
            a = normalized ? mgc_denormalize(c[0], -p[3] / p[4], p[2] * p[4]) : c[0]; /* battery_module.py:332-338 */
        } else {
            a = normalized ? mgc_denormalize(c[0], -1 * p[2], p[1]) : c[0];           /* grid_module.py:125-132 */
        }
        mgc_module_step(V, m, a, t0, fstate, istate, A, info);
        if (M.kind == MGC_GRID) A.done |= ts_done;
    }
    const double ctl_p = mgc_sum_value(A.provided), ctl_a = mgc_sum_value(A.absorbed);
    const double difference = ctl_p - ctl_a;             /* microgrid.py:277-278 */
    /* ---- flex modules, microgrid.py:286-314 ---- */
    if (difference > 0) {
        double excess = difference;
        for (; m < n; ++m) {
            const MgcModule &M = V.mod[m];
            const double *p = V.cfg + M.param_off;
            double amt;
            if (!mgc_is_sink(M.kind)) amt = 0.0;
            else {
                const double mc = mgc_max_consumption(V, M, p, t0, fstate);
                amt = (mc < excess) ? -1.0 * mc : -1.0 * excess;
            }
            mgc_module_step(V, m, amt, t0, fstate, istate, A, info);
            if (M.kind == MGC_RENEWABLE) A.done |= ts_done;
            excess += amt;
        }
        A.flags |= MG_FLAG_EXCESS;
    } else {
        double needed = -difference;
        for (; m < n; ++m) {
            const MgcModule &M = V.mod[m];
            const double *p = V.cfg + M.param_off;
            double amt;
            if (!mgc_is_source(M.kind)) amt = 0.0;
            else {
                const double mp = mgc_max_production(V, M, p, t0, fstate, istate);
                amt = (mp < needed) ? mp : needed;
            }
            mgc_module_step(V, m, amt, t0, fstate, istate, A, info);
            if (M.kind == MGC_RENEWABLE) A.done |= ts_done;
            needed -= amt;
        }
    }
    const double all_p = mgc_sum_value(A.provided), all_a = mgc_sum_value(A.absorbed);
    if (!mgc_isclose(all_p, all_a)) A.flags |= MG_FLAG_BALANCE;      /* microgrid.py:321-323 */
    if (info) {                                                        /* the balance log, microgrid.py:259-260, 281, 317-319 */
        double *b = info + (int64_t)n * MGC_INFO_SLOTS;
        b[0] = fixed_p; b[1] = fixed_a;
        b[2] = ctl_p - fixed_p; b[3] = ctl_a - fixed_a;
        b[4] = all_p; b[5] = all_a;
    }
    t = t0 + 1;                                                        /* every module: _update_step, base_module.py:292-296 */
    *reward_out = A.reward;
    *done_out = (uint8_t)(A.done != 0);
    *flags |= A.flags;
}

/*
 * BaseMicrogridModule.step(action, normalized) (base_module.py:95-159; GensetModule.step, genset_module.py:100-149) for
 * every module of the composition ON ITS OWN -- the reference's operator API used without a Microgrid around it (its
 * module-level tests do): each module takes its own action columns (load: none; renewable, battery, grid, unbalanced: one;
 * genset: goal, energy), normalised actions are mapped through the module's action bounds (utils/space.py:220-231;
 * renewable: its series bounds, base_timeseries_module.py:81-88; unbalanced: (-inf, inf)), then as_source / as_sink /
 * update run as in mgc_module_step.  No energy balance, no flex split.  Columns are assigned in dispatch order.
 */
MGC_HD void mgc_modules_step(const MgcView &V, int32_t &t, double *fstate, int32_t *istate, const double *action, int normalized,
                             int final_step, double *reward_out, uint8_t *done_out, double *info, uint32_t *flags) {
    MgcStepAcc A;
    mgc_sum_init(A.provided);
    mgc_sum_init(A.absorbed);
    A.reward = 0.0;
    A.done = 0;
    A.flags = 0;
    const int t0 = t, n = V.n_mod;
    bool any_series = false;
    for (int m = 0; m < n; ++m) any_series |= mgc_is_timeseries(V.mod[m].kind);
    if (any_series && t0 >= V.T) {
        /* a renewable asked to ABSORB fails before it reads its series (as_sink compares with max_consumption first,
           base_module.py:265): report that too, the wrapper raises it ahead of the IndexError */
        int c0 = 0;
        for (int m = 0; m < n; ++m) {
            const MgcModule &M = V.mod[m];
            if (M.kind == MGC_RENEWABLE) {
                const double *p = V.cfg + M.param_off;
                const double a = normalized ? mgc_denormalize(action[c0], p[1], p[2]) : action[c0];
                if (a < 0) *flags |= MGC_FLAG_NOT_A_SINK;
            }
            c0 += (M.kind == MGC_GENSET) ? 2 : (M.kind == MGC_LOAD) ? 0 : 1;
        }
        *reward_out = NAN;
        *done_out = 1;
        *flags |= MG_FLAG_STEP_PAST_END;
        return;
    }
    if (info)
        for (int i = 0; i < n * MGC_INFO_SLOTS + MGC_BALANCE_SLOTS; ++i) info[i] = 0.0;
    const int ts_done = (t0 >= final_step - 1);
    int col = 0;
    for (int m = 0; m < n; ++m) {
        const MgcModule &M = V.mod[m];
        const double *p = V.cfg + M.param_off;
        const double *c = action + col;
        double a = 0.0;
        switch (M.kind) {
        case MGC_GENSET: {
            const double goal = c[0];
            int32_t *s = istate + M.istate_off;
            if (!(0 <= goal && goal <= 1)) A.flags |= MG_FLAG_GENSET_GOAL_RANGE;
            else {
                int cs = s[0], gs = s[1], up = s[2], dn = s[3];
                mgc_genset_update_status(cs, gs, up, dn, goal, (int)p[5], (int)p[6], p[7] != 0.0);
                s[0] = cs; s[1] = gs; s[2] = up; s[3] = dn;
            }
            a = normalized ? mgc_denormalize(c[1], 0.0, p[1]) : c[1];
            col += 2;
            break;
        }
        case MGC_BATTERY: a = normalized ? mgc_denormalize(c[0], -p[3] / p[4], p[2] * p[4]) : c[0]; col += 1; break;
        case MGC_GRID: a = normalized ? mgc_denormalize(c[0], -1 * p[2], p[1]) : c[0]; col += 1; break;
        case MGC_RENEWABLE: a = normalized ? mgc_denormalize(c[0], p[1], p[2]) : c[0]; col += 1; break;
        case MGC_UNBALANCED: a = normalized ? mgc_denormalize(c[0], -mgc_inf(), mgc_inf()) : c[0]; col += 1; break;
        default: break;      /* load: fixed, takes no action */
        }
        mgc_module_step(V, m, a, t0, fstate, istate, A, info);
        if (mgc_is_timeseries(M.kind)) A.done |= ts_done;
    }
    t = t0 + 1;
    *reward_out = A.reward;
    *done_out = (uint8_t)(A.done != 0);
    *flags |= A.flags;
}

/*
 * PriorityListAlgo._populate_action (algos/priority_list/priority_list.py:69-167) for one env: expand one priority list --
 * `width` elements (dispatch index of a controllable module, action number), negative module = padding -- into the
 * UNNORMALISED action row `ctl` (n_act doubles) that Microgrid.run(normalized=False) then takes.
 *   remaining = sum of the loads - np.sum of the renewables' production (:71-75, 157-167), then per element:
 *   |remaining| <= 1e-4 -> 0 (:92); remaining > 0 -> clip into the module's [min, max] production, a genset's taken
 *   from the status its goal leads to NEXT step (:138-155, genset_module.py:360-424); remaining < 0 -> absorb what a sink
 *   can, sources 0 (:118-136); remaining -= energy.  A genset appears once per goal in a list; only its first
 *   element counts (:84-87).
 */
MGC_HD void mgc_priority_control(const MgcView &V, int t, const double *fstate, const int32_t *istate, const int16_t *pl,
                                 int width, int n_act, double *ctl, uint32_t *flags) {
    double total_load = 0.0;
    MgcSum ren;
    mgc_sum_init(ren);
    for (int m = 0; m < V.n_mod; ++m) {
        const MgcModule &M = V.mod[m];
        const double *p = V.cfg + M.param_off;
        if (M.kind == MGC_LOAD) total_load += -1 * mgc_series_of(V, p)[t];          /* load_module.py:96-111 */
        else if (M.kind == MGC_RENEWABLE) mgc_sum_append(ren, mgc_series_of(V, p)[t]);
    }
    double remaining = total_load - mgc_sum_value(ren);
    for (int i = 0; i < n_act; ++i) ctl[i] = 0.0;
    uint64_t seen = 0;          /* gensets already given their goal by an earlier element */
    for (int i = 0; i < width; ++i) {
        const int m = pl[2 * i], act = pl[2 * i + 1];
        if (m < 0) continue;
        const MgcModule &M = V.mod[m];
        const double *p = V.cfg + M.param_off;
        if (M.kind == MGC_GENSET) {
            if (seen & (1ull << m)) continue;
            seen |= 1ull << m;
            ctl[M.act_col] = (double)act;
        }
        double energy;
        if (fabs(remaining - 0.0) <= 1e-4 + 1e-5 * fabs(0.0)) {       /* np.isclose(remaining, 0.0, atol=1e-4) */
            energy = 0.0;
        } else if (remaining > 0) {
            double mx, mn;
            if (M.kind == MGC_GENSET) {
                const int32_t *s = istate + M.istate_off;
                const int next = act ? ((s[0] || s[2] == 0) ? 1 : 0) : ((!s[0] || s[3] == 0) ? 0 : 1);
                mx = next * p[1];
                mn = next * p[0];
            } else {
                mx = mgc_max_production(V, M, p, t, fstate, istate);
                mn = 0.0;
            }
            if (mn <= remaining && remaining <= mx) energy = remaining;
            else if (remaining < mn) energy = mn;
            else energy = mx;
        } else {
            if (!mgc_is_sink(M.kind)) energy = 0.0;
            else {
                const double mc = mgc_max_consumption(V, M, p, t, fstate);
                if (!(mc >= 0)) *flags |= MG_FLAG_NEGATIVE_ABSORB;     /* :124 assert module_max_consumption >= 0 */
                energy = (-1 * remaining > mc) ? -1.0 * mc : remaining;
            }
        }
        ctl[M.act_col + (M.kind == MGC_GENSET ? 1 : 0)] = energy;
        remaining -= energy;
    }
}

/*
 * GaussianNoiseForecaster (forecast/forecaster.py:220-262) on a freshly written observation row: N(0, std_k) on every REAL
 * forecast row k of a time-series module (rows past the end of the series are padding and carry none, :120-132), clipped to
 * the column's bounds (:139-149) -- in normalised units  obs <- min(max(obs + z * sigma * scale_k, 0), 1)  with
 * scale_k = 1 + log(1 + k) under increase_uncertainty (:244-248).  `sigma` / `increase` are per ELEMENT of the row (the host
 * folds std, relative_noise and the column spread into sigma; 0 = leave the element alone).  z: Box-Muller over
 * Philox4x32-10 (Salmon et al., SC'11), a pure function of (seed, call, global env id, the env's step, element pair) -- the
 * same counter layout as mg_forecast_noise of the fused path.  Lane `lane` of a warp takes pairs lane, lane + 32, ...
 */
MGC_HD void mgc_philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t m0 = (uint64_t)0xD2511F53u * c[0], m1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t hi0 = (uint32_t)(m0 >> 32), lo0 = (uint32_t)m0, hi1 = (uint32_t)(m1 >> 32), lo1 = (uint32_t)m1;
        c[0] = hi1 ^ c[1] ^ k0;
        c[1] = lo1;
        c[2] = hi0 ^ c[3] ^ k1;
        c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

MGC_HD void mgc_noise_row(const MgcView &V, const int32_t *elem, int obs_dim, double *row, const double *sigma,
                          const double *increase, int t, uint64_t env, uint32_t k0, uint32_t k1, uint32_t c3, int lane) {
    for (int p = lane; 2 * p < obs_dim; p += 32) {
        double z[2];
        bool drawn = false;
        for (int h = 0; h < 2; ++h) {
            const int j = 2 * p + h;
            if (j >= obs_dim || sigma[j] == 0.0) continue;
            const int32_t d = elem[j];
            const MgcModule &M = V.mod[d >> 16];
            if (!mgc_is_timeseries(M.kind)) continue;
            const int C = (M.kind == MGC_GRID) ? 4 : 1;
            const int r = (d & 0xffff) / C;
            if (r == 0) continue;                       /* the current value is not a forecast */
            const int k = r - 1;
            if (t + 1 + k >= V.T) continue;             /* padding past the end of the series */
            if (!drawn) {
                uint32_t c[4] = {(uint32_t)env, ((uint32_t)(env >> 32) & 0xffffu) | ((uint32_t)p << 16), (uint32_t)t, c3};
                mgc_philox4x32_10(c, k0, k1);
                /* two 53-bit uniforms, u1 in (0, 1], u2 in [0, 1) */
                const double u1 = ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6) + 1.0) * (1.0 / 9007199254740992.0);
                const double u2 = ((double)(c[2] >> 5) * 67108864.0 + (double)(c[3] >> 6)) * (1.0 / 9007199254740992.0);
                const double radius = sqrt(-2.0 * log(u1));
#if defined(__CUDA_ARCH__)
                double sn, cs;
                sincospi(2.0 * u2, &sn, &cs);
#else
                const double sn = sin(6.283185307179586476925286766559 * u2), cs = cos(6.283185307179586476925286766559 * u2);
#endif
                z[0] = radius * cs;
                z[1] = radius * sn;
                drawn = true;
            }
            double s = sigma[j];
            if (increase[j] != 0.0) s = s * (1.0 + log(1.0 + (double)k));
            row[j] = fmin(fmax(row[j] + z[h] * s, 0.0), 1.0);
        }
    }
}

#endif /* MG_COMPOSE_STEP_H */
