/*
 * mg_oracle.c -- CPU restatement of pymgrid's Microgrid.run hot path (TEST INFRASTRUCTURE, not product).
 *
 * Every function cites the reference lines it follows (relative to /root/reference/src/pymgrid/).
 * Arithmetic is scalar IEEE f64 in the reference's operation order; compile with -ffp-contract=off so
 * that gcc never fuses a*b+c (Python/numpy never do).
 *
 * Pinned by tests/test_oracle_vs_golden.py against golden vectors recorded from the live reference and
 * against the reference's own known-answer tests.  The shipped CUDA path never calls into this file.
 */
#include "mg_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * utils/space.py:183-231  ModuleSpace: spread = high - low with 0 -> 1; normalize = (v - low) / spread;
 * denormalize = low + spread * v
 * ---------------------------------------------------------------------------------------------- */
static double space_spread(double low, double high) {
    double s = high - low;
    return s == 0.0 ? 1.0 : s;
}
static double space_normalize(double v, double low, double high) { return (v - low) / space_spread(low, high); }
static double space_denormalize(double x, double low, double high) { return low + space_spread(low, high) * x; }

/* numpy.isclose(a, b) with default rtol=1e-5, atol=1e-8 */
static int np_isclose(double a, double b, double rtol, double atol);
static int shaper_in_range(double v) {
    return (-1 <= v && v <= 1) || np_isclose(v, 1.0, 1e-5, 1e-8) || np_isclose(v, 0.0, 1e-5, 1e-8);
}

static int np_isclose(double a, double b, double rtol, double atol) {
    return fabs(a - b) <= (atol + rtol * fabs(b));
}

int orc_obs_dim(const OrcGrid *g) {
    int rows = 1 + g->horizon;
    return rows * (2 + 4 * (g->has_grid != 0)) + 2 + 4 * (g->has_genset != 0);
}
int orc_n_act(const OrcGrid *g) { return 1 + (g->has_grid != 0) + 2 * (g->has_genset != 0); }

/* ------------------------------------------------------------------------------------------------
 * Time-series bounds.  base_timeseries_module.py:81-88 (load / pv: scalar bounds, pulled to include 0)
 * and grid_module.py:125-132 (grid: per-column min / max, no zero adjustment).
 * ---------------------------------------------------------------------------------------------- */
static void scalar_series_bounds(const double *ts, int T, double *low, double *high) {
    double mn = ts[0], mx = ts[0];
    for (int i = 1; i < T; ++i) {
        if (ts[i] < mn) mn = ts[i];
        if (ts[i] > mx) mx = ts[i];
    }
    if (mn > 0) mn = 0;
    else if (mx < 0) mx = 0;
    *low = mn;
    *high = mx;
}
static void grid_series_bounds(const double *ts, int T, double low[4], double high[4]) {
    for (int c = 0; c < 4; ++c) {
        double mn = ts[c], mx = ts[c];
        for (int i = 1; i < T; ++i) {
            double v = ts[4 * (size_t)i + c];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        low[c] = mn;
        high[c] = mx;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Normalised state of a time-series module at step t: [current(C), forecast_0(C) ... forecast_{H-1}(C)].
 *   current  : ts[t]  or the fill row when t is past the end    base_timeseries_module.py:127-140
 *   forecast : ts[t+1 : t+1+H], short windows padded with (high+low)/2 rows, then clipped to bounds
 *              base_timeseries_module.py:103-122, forecast/forecaster.py:95,120-149,172-187,215-217
 *   normalise: (state - low) / spread                                      utils/space.py:207-218
 * ---------------------------------------------------------------------------------------------- */
static void series_obs(const double *ts, int T, int C, int t, int H, const double *low, const double *high,
                       double *out) {
    for (int k = 0; k <= H; ++k) {
        int idx = t + k;
        for (int c = 0; c < C; ++c) {
            double v;
            if (idx < T && t < T) {
                v = ts[(size_t)idx * C + c];
                if (k > 0) { /* Forecaster._clip: only forecast rows are clipped */
                    if (v < low[c]) v = low[c];
                    if (v > high[c]) v = high[c];
                }
            } else {
                v = (high[c] + low[c]) / 2; /* Forecaster._fill_arr */
                if (k > 0 && t < T) {
                    if (v < low[c]) v = low[c];
                    if (v > high[c]) v = high[c];
                }
            }
            out[k * C + c] = space_normalize(v, low[c], high[c]);
        }
    }
}

/* genset_module.py:503-509 obs bounds [0,0,0,0]..[1,1,U,D];  battery_module.py:323-330 */
static void genset_obs(const OrcGrid *g, double out[4]) {
    out[0] = space_normalize((double)g->cs, 0.0, 1.0);
    out[1] = space_normalize((double)g->gs, 0.0, 1.0);
    out[2] = space_normalize((double)g->up, 0.0, (double)g->start_up_time);
    out[3] = space_normalize((double)g->dn, 0.0, (double)g->wind_down_time);
}
static void battery_obs(const OrcGrid *g, double out[2]) {
    double min_soc = g->min_capacity / g->max_capacity; /* battery_module.py:88 */
    out[0] = space_normalize(g->soc, min_soc, 1.0);     /* _soc as stored, battery_module.py:89, 130 */
    out[1] = space_normalize(g->charge, g->min_capacity, g->max_capacity);
}

/* envs/base/base.py:211-223 flatten; module listing order microgrid.py / module_container.py:355-413 */
void orc_prepare(OrcGrid *g) {
    scalar_series_bounds(g->load_ts, g->T, &g->load_low, &g->load_high);
    scalar_series_bounds(g->pv_ts, g->T, &g->pv_low, &g->pv_high);
    if (g->has_grid) grid_series_bounds(g->grid_ts, g->T, g->grid_low, g->grid_high);
    g->prepared = 1;
}

void orc_observe(const OrcGrid *g, int order, double *obs) {
    int rows = 1 + g->horizon;
    double llow = g->load_low, lhigh = g->load_high, plow = g->pv_low, phigh = g->pv_high;
    const double *glow = g->grid_low, *ghigh = g->grid_high;
    double bat[2], gen[4];
    battery_obs(g, bat);
    if (g->has_genset) genset_obs(g, gen);
    double *p = obs;
    if (order == ORC_ORDER_GYM_SORTED) { /* battery, genset, grid, load, pv (alphabetical Dict keys) */
        memcpy(p, bat, sizeof bat); p += 2;
        if (g->has_genset) { memcpy(p, gen, sizeof gen); p += 4; }
        if (g->has_grid) { series_obs(g->grid_ts, g->T, 4, g->t, g->horizon, glow, ghigh, p); p += 4 * rows; }
        series_obs(g->load_ts, g->T, 1, g->t, g->horizon, &llow, &lhigh, p); p += rows;
        series_obs(g->pv_ts, g->T, 1, g->t, g->horizon, &plow, &phigh, p); p += rows;
    } else { /* container listing order: load, pv, (unbalanced: empty), genset, battery, grid */
        series_obs(g->load_ts, g->T, 1, g->t, g->horizon, &llow, &lhigh, p); p += rows;
        series_obs(g->pv_ts, g->T, 1, g->t, g->horizon, &plow, &phigh, p); p += rows;
        if (g->has_genset) { memcpy(p, gen, sizeof gen); p += 4; }
        memcpy(p, bat, sizeof bat); p += 2;
        if (g->has_grid) { series_obs(g->grid_ts, g->T, 4, g->t, g->horizon, glow, ghigh, p); p += 4 * rows; }
    }
}

void orc_reset(OrcGrid *g) { g->t = g->initial_step; }

/* ------------------------------------------------------------------------------------------------
 * Genset state machine.  genset_module.py:216-233 (_reset_up_down_times, _update_up_down_times),
 * :235-300 (update_status), :302-311 (_finish_in_progress_change), :327-346 (_non_instantaneous_update),
 * :360-390 (next_status)
 * ---------------------------------------------------------------------------------------------- */
static void genset_reset_up_down(OrcGrid *g) {
    if (g->cs) { g->up = 0; g->dn = g->wind_down_time; }
    else { g->dn = 0; g->up = g->start_up_time; }
}
int orc_genset_next_status(const OrcGrid *g, int goal) {
    if (goal) return (g->cs || g->up == 0) ? 1 : 0;
    return (!g->cs || g->dn == 0) ? 0 : 1;
}
void orc_genset_update_status(OrcGrid *g, double goal_status) {
    /* Python round(): half to even, so 0.5 -> 0; rint under the default rounding mode is the same */
    int goal = (int)rint(goal_status);
    if (goal == g->cs && g->cs == g->gs) return; /* :284-287 */
    int instant_up = (g->start_up_time == 0 && goal == 1);
    int instant_down = (g->wind_down_time == 0 && goal == 0);
    if (goal != g->gs && (g->allow_abortion || instant_up || instant_down)) g->gs = goal; /* :289-292 */
    /* _finish_in_progress_change :302-311 */
    if (g->up == 0 && g->gs == 1) { g->cs = 1; genset_reset_up_down(g); return; }
    if (g->dn == 0 && g->gs == 0) { g->cs = 0; genset_reset_up_down(g); return; }
    /* _non_instantaneous_update :327-346 */
    if (goal == g->cs && g->cs != g->gs && g->allow_abortion) {
        g->gs = goal;
        genset_reset_up_down(g);
    } else if (g->cs == g->gs && g->gs != goal) {
        genset_reset_up_down(g);
        g->gs = goal;
    }
    if (g->gs != g->cs) {
        if (g->gs == 0) g->dn -= 1;
        else g->up -= 1;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Microgrid.run  microgrid/microgrid.py:227-325 with MicrogridStep (microgrid/utils/step.py) sums.
 * Dispatch order: load -> genset -> battery -> grid -> (balance) -> pv -> unbalanced_energy.
 * ---------------------------------------------------------------------------------------------- */
void orc_run(OrcGrid *g, const double *control, int normalized, int order, double *obs, double *reward_out,
             int32_t *done_out, double *info, uint32_t *err_out) {
    uint32_t err = 0;
    double inf[ORC_N_INFO];
    for (int i = 0; i < ORC_N_INFO; ++i) inf[i] = 0.0;
    const int t = g->t;
    if (t >= g->T) { /* the reference raises IndexError in LoadModule.update */
        err |= ORC_ERR_STEP_PAST_END;
        if (reward_out) *reward_out = NAN;
        if (done_out) *done_out = 1;
        if (err_out) *err_out = err;
        return;
    }
    double reward = 0.0;   /* MicrogridStep._reward, step.py:9,18 */
    double provided = 0.0; /* np.sum(info['provided_energy']): sequential from 0.0 for < 8 items */
    double consumed = 0.0;
    int done = 0;
    /* BaseTimeSeriesMicrogridModule._done, base_timeseries_module.py:124-125 (evaluated before t += 1) */
    const int ts_done = (t >= g->final_step - 1);

    /* ---- fixed: LoadModule.update load_module.py:86-91 (absorbs -ts[t], reward 0.0) ---- */
    double load = -1 * g->load_ts[t];
    inf[ORC_INFO_LOAD_MET] = load;
    consumed += load;
    reward += 0.0;
    done |= ts_done;

    const double *c = control;
    /* ---- controllable: genset (source) ---- */
    if (g->has_genset) {
        double goal = c[0]; /* never denormalised: genset_module.py:146 */
        double a;
        if (!(0 <= goal && goal <= 1)) err |= ORC_ERR_GENSET_GOAL_RANGE;
        else orc_genset_update_status(g, goal); /* :148, BEFORE the production clamp */
        a = normalized ? space_denormalize(c[1], 0.0, g->running_max_production) : c[1]; /* :511-517 */
        c += 2;
        double p;
        if (a < 0) { /* base_module.py:164-165 as_sink -> GensetModule.update assert as_source */
            err |= ORC_ERR_GENSET_AS_SINK;
            p = 0.0;
        } else {
            double mx = g->cs * g->running_max_production; /* :482 */
            double mn = g->cs * g->running_min_production; /* :501 */
            /* base_module.py:213-224: upper test first */
            if (a > mx) { p = mx; err |= ORC_CLIP_GENSET; }
            else if (a < mn) { p = mn; err |= ORC_CLIP_GENSET; }
            else p = a;
        }
        double co2 = g->co2_per_unit * p;                                         /* :165 */
        double cost = g->genset_cost * p + g->gen_cost_per_unit_co2 * co2;        /* :186, :181, :205 */
        inf[ORC_INFO_REWARD_GENSET] = -1.0 * cost;                                /* :210 */
        reward += inf[ORC_INFO_REWARD_GENSET];
        provided += p;
        inf[ORC_INFO_GENSET_PRODUCTION] = p;
        inf[ORC_INFO_GENSET_CO2] = co2;
    }
    /* ---- controllable: battery (source_and_sink) battery_module.py ---- */
    {
        double lo = -g->max_discharge / g->efficiency; /* min_act :332-334 */
        double hi = g->max_charge * g->efficiency;     /* max_act :336-338 */
        double a = normalized ? space_denormalize(c[0], lo, hi) : c[0];
        c += 1;
        double internal;
        if (a > 0 || a == 0) { /* base_module.py:161-171: a == 0 and is_source -> as_source */
            double mp = fmin(g->max_discharge, g->charge - g->min_capacity) * g->efficiency; /* :283-286 */
            double p; /* min_production is 0 for the battery (base_module.py:604-619), a >= 0 here */
            if (a > mp) { p = mp; err |= ORC_CLIP_BATTERY; }
            else p = a;
            internal = (-1.0 * p) / g->efficiency; /* update :113 -> default_transition_model :275-276 */
            provided += p;
            inf[ORC_INFO_BATTERY_DISCHARGE] = p;
        } else {
            err |= ORC_BATTERY_SINK;
            double e = -1.0 * a;
            double mc = fmin(g->max_charge, g->max_capacity - g->charge) / g->efficiency; /* :288-291 */
            if (e > mc) { e = mc; err |= ORC_CLIP_BATTERY; }
            if (!(e >= 0)) err |= ORC_ERR_NEGATIVE_ABSORB;
            internal = e * g->efficiency; /* :117 -> :277-278 */
            consumed += e;
            inf[ORC_INFO_BATTERY_CHARGE] = e;
        }
        g->charge += internal; /* _update_state :125-130 */
        if (g->charge < g->min_capacity) {
            if (!np_isclose(g->charge, g->min_capacity, 1e-5, 1e-8)) err |= ORC_ERR_BATTERY_MIN_CAP;
            g->charge = g->min_capacity;
        }
        g->soc = g->charge / g->max_capacity; /* :130 */
        inf[ORC_INFO_REWARD_BATTERY] = -1.0 * (fabs(internal) * g->battery_cost_cycle); /* :121, get_cost :147 */
        reward += inf[ORC_INFO_REWARD_BATTERY];
    }
    /* ---- controllable: grid (source_and_sink) grid_module.py ---- */
    if (g->has_grid) {
        const double *row = g->grid_ts + 4 * (size_t)t;
        double a = normalized ? space_denormalize(c[0], -1 * g->max_export, g->max_import) : c[0]; /* :130 */
        c += 1;
        double status = row[3];
        if (a > 0 || a == 0) { /* import */
            double mp = g->max_import * status; /* :314-316 */
            double p;
            if (a > mp) { p = mp; err |= ORC_CLIP_GRID; }
            else p = a;
            double co2 = p * row[2];                                           /* :221-224 */
            inf[ORC_INFO_REWARD_GRID] = -1 * row[0] * p + (-1.0 * g->grid_cost_per_unit_co2 * co2); /* :167-169, :197 */
            reward += inf[ORC_INFO_REWARD_GRID];
            provided += p;
            inf[ORC_INFO_GRID_IMPORT] = p;
            inf[ORC_INFO_GRID_CO2] = co2;
        } else { /* export */
            err |= ORC_GRID_SINK;
            double e = -1.0 * a;
            double mc = g->max_export * status; /* :318-320 */
            if (e > mc) { e = mc; err |= ORC_CLIP_GRID; }
            inf[ORC_INFO_REWARD_GRID] = row[1] * e + (-1.0 * g->grid_cost_per_unit_co2 * 0.0); /* :170-172, :225-226 */
            reward += inf[ORC_INFO_REWARD_GRID];
            consumed += e;
            inf[ORC_INFO_GRID_EXPORT] = e;
        }
        done |= ts_done;
    }

    /* ---- flex modules: pv then unbalanced_energy.  microgrid.py:277-314 ---- */
    double difference = provided - consumed;
    double pv = g->pv_ts[t];
    if (difference > 0) {
        err |= ORC_EXCESS;
        /* pv: not a sink -> step(0.0) -> as_source(0.0) -> provides 0.0, curtailment = pv - 0.0 */
        double pv_used = 0.0;
        inf[ORC_INFO_PV_USED] = pv_used;
        inf[ORC_INFO_CURTAILMENT] = pv - pv_used;
        provided += pv_used;
        reward += 0.0;
        /* unbalanced: max_consumption = inf -> absorbs all the excess */
        double excess = difference;
        inf[ORC_INFO_OVERGENERATION] = excess;
        consumed += excess;
        inf[ORC_INFO_REWARD_UNBALANCED] = -1.0 * (g->overgeneration_cost * excess); /* unbalanced_energy_module.py:28-36,65-68 */
        reward += inf[ORC_INFO_REWARD_UNBALANCED];
    } else {
        double needed = -difference;
        double pv_used = (pv < needed) ? pv : needed; /* microgrid.py:305-310 */
        inf[ORC_INFO_PV_USED] = pv_used;
        inf[ORC_INFO_CURTAILMENT] = pv - pv_used;
        provided += pv_used;
        reward += 0.0;
        needed -= pv_used;
        inf[ORC_INFO_LOSS_LOAD] = needed;
        provided += needed;
        inf[ORC_INFO_REWARD_UNBALANCED] = -1.0 * (g->loss_load_cost * needed);
        reward += inf[ORC_INFO_REWARD_UNBALANCED];
    }
    done |= ts_done; /* pv is a time-series module too */
    if (!np_isclose(provided, consumed, 1e-5, 1e-8)) err |= ORC_ERR_BALANCE; /* microgrid.py:321-323 */

    g->t = t + 1; /* every module: _update_step, base_module.py:292-296 */

    /* MicrogridStep.output -> shaped_reward, microgrid/utils/step.py:38-46 */
    if (g->reward_shaper == ORC_SHAPER_PV_CURTAILMENT)          /* reward_shaping/pv_curtailment_shaper.py:16-18 */
        reward = -1.0 * inf[ORC_INFO_CURTAILMENT];
    else if (g->reward_shaper == ORC_SHAPER_BATTERY_DISCHARGE) { /* reward_shaping/battery_discharge_shaper.py:23-35 */
        /* the reference calls the shaper twice: in balance() before the flex modules (microgrid.py:277; no
           unbalanced_energy entry yet -> loss load 0.0) and for the output; both run the assert at :33 */
        double mid = (inf[ORC_INFO_BATTERY_DISCHARGE] - 0.0) / inf[ORC_INFO_LOAD_MET];
        reward = (inf[ORC_INFO_BATTERY_DISCHARGE] - inf[ORC_INFO_LOSS_LOAD]) / inf[ORC_INFO_LOAD_MET];
        if (!shaper_in_range(mid) || !shaper_in_range(reward)) err |= ORC_ERR_SHAPER_RANGE;
    }

    if (obs) orc_observe(g, order, obs);
    if (reward_out) *reward_out = reward;
    if (done_out) *done_out = done;
    if (info) memcpy(info, inf, sizeof inf);
    if (err_out) *err_out = err;
}

/* ------------------------------------------------------------------------------------------------
 * PriorityListAlgo._populate_action  algos/priority_list/priority_list.py:69-167
 * ---------------------------------------------------------------------------------------------- */
uint32_t orc_priority_control(const OrcGrid *g, const int8_t *plist_module, const int8_t *plist_action, int n_el,
                              double *control) {
    uint32_t err = 0;
    const int t = g->t;
    int gen_off = 0, bat_off = g->has_genset ? 2 : 0, grid_off = bat_off + 1;
    int genset_seen = 0;
    double total_load = 0.0;
    total_load += -1 * g->load_ts[t];               /* _get_load :157-163 */
    double renewable = 0.0 + g->pv_ts[t];           /* _get_renewable :165-166 (np.sum of one item) */
    double remaining = total_load - renewable;      /* :75 */
    for (int i = 0; i < orc_n_act(g); ++i) control[i] = 0.0;
    for (int i = 0; i < n_el; ++i) {
        int mod = plist_module[i];
        int act = plist_action[i];
        if (mod < 0) continue; /* LUT padding */
        if (mod == 0) {
            if (genset_seen) continue; /* :84-87 already hit this module */
            genset_seen = 1;
            control[gen_off] = (double)act;
        }
        double energy;
        if (np_isclose(remaining, 0.0, 1e-5, 1e-4)) { /* :92 np.isclose(remaining, 0.0, atol=1e-4) */
            energy = 0.0;
        } else if (remaining > 0) { /* _produce_from_module :138-155 */
            double mx, mn;
            if (mod == 0) {
                int ns = orc_genset_next_status(g, act);
                mx = ns * g->running_max_production; /* genset_module.py:392-424 */
                mn = ns * g->running_min_production;
            } else if (mod == 1) {
                mx = fmin(g->max_discharge, g->charge - g->min_capacity) * g->efficiency;
                mn = 0.0;
            } else {
                mx = g->max_import * g->grid_ts[4 * (size_t)t + 3];
                mn = 0.0;
            }
            if (mn <= remaining && remaining <= mx) energy = remaining;
            else if (remaining < mn) energy = mn;
            else energy = mx;
        } else { /* _consume_in_module :118-136 */
            if (mod == 0) energy = 0.0; /* genset is not a sink */
            else {
                double mc = (mod == 1) ? fmin(g->max_charge, g->max_capacity - g->charge) / g->efficiency
                                       : g->max_export * g->grid_ts[4 * (size_t)t + 3];
                if (!(mc >= 0)) err |= ORC_ERR_NEGATIVE_ABSORB; /* :124 assert module_max_consumption >= 0 */
                if (-1 * remaining > mc) energy = -1.0 * mc;
                else energy = remaining;
            }
        }
        if (mod == 0) control[gen_off + 1] = energy;
        else if (mod == 1) control[bat_off] = energy;
        else control[grid_off] = energy;
        remaining -= energy; /* :108 */
    }
    return err;
}

/* ------------------------------------------------------------------------------------------------
 * Batched drivers (timing + bulk parity).  Each worker owns a contiguous slice of grids.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    OrcGrid *grids;
    int64_t lo, hi, n;
    const double *actions;
    const int32_t *dactions;
    int32_t max_act, n_steps, normalized, order, obs_stride, plist_len;
    const int8_t *plist_module, *plist_action;
    const int32_t *plist_offset;
    double *rewards;
    uint8_t *dones;
    double *obs_last;
} RolloutJob;

static void *rollout_worker(void *arg) {
    RolloutJob *j = (RolloutJob *)arg;
    double *scratch = (double *)malloc(sizeof(double) * (size_t)(j->obs_stride > 0 ? j->obs_stride : 1));
    double ctrl[4];
    for (int64_t i = j->lo; i < j->hi; ++i)
        if (!j->grids[i].prepared) orc_prepare(&j->grids[i]);
    /* env-major: one grid runs all its steps back to back (its record and series window stay in L1) */
    for (int64_t i = j->lo; i < j->hi; ++i) {
        OrcGrid *g = &j->grids[i];
        for (int32_t s = 0; s < j->n_steps; ++s) {
            double r;
            int32_t d;
            double *obs = (s == j->n_steps - 1 && j->obs_last) ? j->obs_last + (size_t)i * j->obs_stride : scratch;
            const double *c;
            int normalized = j->normalized;
            if (j->dactions) {
                int32_t a = j->dactions[(size_t)s * j->n + i];
                size_t row = (size_t)(j->plist_offset[i] + a) * j->plist_len;
                orc_priority_control(g, j->plist_module + row, j->plist_action + row, j->plist_len, ctrl);
                c = ctrl;
                normalized = 0;
            } else {
                c = j->actions + ((size_t)s * j->n + i) * j->max_act;
            }
            orc_run(g, c, normalized, j->order, obs, &r, &d, NULL, NULL);
            j->rewards[(size_t)s * j->n + i] = r;
            j->dones[(size_t)s * j->n + i] = (uint8_t)d;
        }
    }
    free(scratch);
    return NULL;
}

static void run_jobs(RolloutJob *proto, int32_t n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > proto->n) n_threads = (int32_t)(proto->n > 0 ? proto->n : 1);
    RolloutJob *jobs = (RolloutJob *)malloc(sizeof(RolloutJob) * n_threads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int32_t k = 0; k < n_threads; ++k) {
        jobs[k] = *proto;
        jobs[k].lo = proto->n * k / n_threads;
        jobs[k].hi = proto->n * (k + 1) / n_threads;
        if (n_threads == 1) rollout_worker(&jobs[k]);
        else pthread_create(&th[k], NULL, rollout_worker, &jobs[k]);
    }
    if (n_threads > 1)
        for (int32_t k = 0; k < n_threads; ++k) pthread_join(th[k], NULL);
    free(jobs);
    free(th);
}

void orc_rollout(OrcGrid *grids, int64_t n, const double *actions, int32_t max_act, int32_t n_steps, int normalized,
                 int order, double *rewards, uint8_t *dones, double *obs_last, int32_t obs_stride,
                 int32_t n_threads) {
    RolloutJob j;
    memset(&j, 0, sizeof j);
    j.grids = grids; j.n = n; j.actions = actions; j.max_act = max_act; j.n_steps = n_steps;
    j.normalized = normalized; j.order = order; j.rewards = rewards; j.dones = dones;
    j.obs_last = obs_last; j.obs_stride = obs_stride;
    run_jobs(&j, n_threads);
}

void orc_rollout_discrete(OrcGrid *grids, int64_t n, const int32_t *actions, int32_t n_steps,
                          const int8_t *plist_module, const int8_t *plist_action, const int32_t *plist_offset,
                          int32_t plist_len, int order, double *rewards, uint8_t *dones, double *obs_last,
                          int32_t obs_stride, int32_t n_threads) {
    RolloutJob j;
    memset(&j, 0, sizeof j);
    j.grids = grids; j.n = n; j.dactions = actions; j.n_steps = n_steps; j.order = order;
    j.plist_module = plist_module; j.plist_action = plist_action; j.plist_offset = plist_offset;
    j.plist_len = plist_len; j.rewards = rewards; j.dones = dones; j.obs_last = obs_last;
    j.obs_stride = obs_stride;
    run_jobs(&j, n_threads);
}
