/*
 * mg_oracle.h -- CPU restatement of the reference microgrid step (TEST INFRASTRUCTURE, not product).
 *
 * Scalar IEEE-f64 C, one function per reference module, same operation order as the Python reference
 * (Total-RD/pymgrid @ 7bf3951).  Citations are relative to /root/reference/src/pymgrid/.
 *
 * Pinned: tests/test_oracle_vs_golden.py checks this oracle bit-for-bit against golden vectors recorded
 * from the live reference (tests/golden/make_golden.py) and against the known-answer values of the
 * reference's own unit tests (genset state machine, load/PV balance, time-series windows).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call this.
 */
#ifndef MG_ORACLE_H
#define MG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* info vector layout written by orc_run (one double each) */
enum {
    ORC_INFO_LOAD_MET = 0,      /* load 'absorbed_energy'            load_module.py:86-91        */
    ORC_INFO_PV_USED = 1,       /* pv 'provided_energy'              renewable_module.py:86-93   */
    ORC_INFO_CURTAILMENT = 2,   /* pv 'curtailment'                                              */
    ORC_INFO_LOSS_LOAD = 3,     /* unbalanced 'provided_energy'      unbalanced_energy_module.py */
    ORC_INFO_OVERGENERATION = 4,/* unbalanced 'absorbed_energy'                                  */
    ORC_INFO_GENSET_PRODUCTION = 5,
    ORC_INFO_GENSET_CO2 = 6,
    ORC_INFO_BATTERY_DISCHARGE = 7,
    ORC_INFO_BATTERY_CHARGE = 8,
    ORC_INFO_GRID_IMPORT = 9,
    ORC_INFO_GRID_EXPORT = 10,
    ORC_INFO_GRID_CO2 = 11,
    ORC_INFO_REWARD_GENSET = 12, /* per-module 'reward' log entries, base_module.py:276-290 */
    ORC_INFO_REWARD_BATTERY = 13,
    ORC_INFO_REWARD_GRID = 14,
    ORC_INFO_REWARD_UNBALANCED = 15,
    ORC_N_INFO = 16
};

/* error / event flags (the reference raises exceptions or clips silently at these points) */
enum {
    ORC_ERR_GENSET_GOAL_RANGE = 1u << 0,   /* genset_module.py:147 assert 0 <= goal <= 1                 */
    ORC_ERR_GENSET_AS_SINK    = 1u << 1,   /* genset_module.py:208 assert as_source                      */
    ORC_ERR_BALANCE           = 1u << 2,   /* microgrid.py:321 RuntimeError                              */
    ORC_ERR_BATTERY_MIN_CAP   = 1u << 3,   /* battery_module.py:128 assert isclose                       */
    ORC_ERR_NEGATIVE_ABSORB   = 1u << 4,   /* base_module.py:272 assert absorbed_energy >= 0; priority_list.py:124 */
    ORC_ERR_STEP_PAST_END     = 1u << 5,   /* IndexError on ts[t] when t >= len                          */
    ORC_ERR_SHAPER_RANGE      = 1u << 7,   /* reward_shaping/battery_discharge_shaper.py:33 assert       */
    ORC_CLIP_GENSET           = 1u << 8,   /* raise_errors=True would raise ValueError here              */
    ORC_CLIP_BATTERY          = 1u << 9,   /*   (base_module.py:213-221, 265-268)                        */
    ORC_CLIP_GRID             = 1u << 10,
    ORC_BATTERY_SINK          = 1u << 12,  /* direction bits: which info key the reference wrote                 */
    ORC_GRID_SINK             = 1u << 13,
    ORC_EXCESS                = 1u << 14
};

enum { ORC_ORDER_GYM_SORTED = 0, ORC_ORDER_CONTAINER = 1 };
enum { ORC_SHAPER_NONE = 0, ORC_SHAPER_PV_CURTAILMENT = 1, ORC_SHAPER_BATTERY_DISCHARGE = 2 };

typedef struct OrcGrid {
    /* architecture */
    int32_t has_genset, has_grid;
    int32_t horizon;              /* forecast_horizon H (0 = no forecaster)  */
    int32_t T;                    /* len(time_series)                         */
    int32_t initial_step, final_step;
    /* battery_module.py:66-91 */
    double min_capacity, max_capacity, max_charge, max_discharge, efficiency, battery_cost_cycle;
    /* genset_module.py:61-92 */
    double running_min_production, running_max_production, genset_cost, co2_per_unit, gen_cost_per_unit_co2;
    int32_t start_up_time, wind_down_time, allow_abortion;
    int32_t reward_shaper;        /* ORC_SHAPER_*: Microgrid(reward_shaping_func=...), microgrid/utils/step.py:41-46 */
    /* grid_module.py:70-101 */
    double max_import, max_export, grid_cost_per_unit_co2;
    /* unbalanced_energy_module.py:14-26 */
    double loss_load_cost, overgeneration_cost;
    /* time series as the reference stores them: load negative (base_timeseries_module.py:68-79),
       pv >= 0, grid [T][4] = import_price, export_price, co2_per_kwh, grid_status                    */
    const double *load_ts, *pv_ts, *grid_ts;
    /* mutable state */
    int32_t t;
    int32_t cs, gs, up, dn;      /* genset: current_status, goal_status, steps_until_up, steps_until_down */
    int32_t _pad1;
    double charge;               /* battery _current_charge */
    double soc;                  /* battery _soc: what the constructor was given (init_soc, or init_charge / max_capacity,
                                    battery_module.py:96-106) until the first _update_state recomputes it (:125-130) */
    /* observation bounds, computed once by orc_prepare like the reference does at module construction
       (base_timeseries_module.py:81-88 for load/pv, grid_module.py:125-132 per grid column)            */
    int32_t prepared, _pad2;
    double load_low, load_high, pv_low, pv_high, grid_low[4], grid_high[4];
} OrcGrid;

/* compute the cached observation bounds; must be called once before orc_run / orc_observe */
void orc_prepare(OrcGrid *g);

int orc_obs_dim(const OrcGrid *g);
int orc_n_act(const OrcGrid *g);

/* one Microgrid.run (microgrid/microgrid.py:227-325).
 * control: container order of controllables (genset[goal, energy] if present, battery, grid if present).
 * obs: normalised post-step observation in `order`; may be NULL.  info: ORC_N_INFO doubles or NULL. */
void orc_run(OrcGrid *g, const double *control, int normalized, int order,
             double *obs, double *reward, int32_t *done, double *info, uint32_t *err);

/* normalised observation of the CURRENT state (what reset() returns; base_module.py:65-77) */
void orc_observe(const OrcGrid *g, int order, double *obs);

/* Microgrid.reset: t = initial_step, battery/genset state untouched (microgrid.py:205-225) */
void orc_reset(OrcGrid *g);

/* genset state machine alone (genset_module.py:235-346); returns nothing, mutates cs/gs/up/dn */
void orc_genset_update_status(OrcGrid *g, double goal_status);
int orc_genset_next_status(const OrcGrid *g, int goal_status);

/* PriorityListAlgo._populate_action (algos/priority_list/priority_list.py:69-116).
 * plist: n_el elements, each (module, action) with module 0=genset 1=battery 2=grid.
 * Writes the UNNORMALISED control in container order (same layout as orc_run's control).
 * Returns ORC_ERR_NEGATIVE_ABSORB where the reference's `assert module_max_consumption >= 0` (:124) fires (a battery whose
 * charge sits an ulp above max_capacity is asked to absorb), else 0. */
uint32_t orc_priority_control(const OrcGrid *g, const int8_t *plist_module, const int8_t *plist_action,
                              int n_el, double *control);

/* Batched drivers used only for CPU-baseline timing (bench.py) and bulk parity tests.
 * grids[n]; actions [n_steps][n][max_act] row-major with stride max_act; rewards/dones [n_steps][n];
 * obs_last [n][obs_stride] receives the last step's observation (every step's obs is computed and
 * written to a per-thread scratch row, as the reference does).  n_threads >= 1 (pthreads). */
void orc_rollout(OrcGrid *grids, int64_t n, const double *actions, int32_t max_act, int32_t n_steps,
                 int normalized, int order, double *rewards, uint8_t *dones, double *obs_last,
                 int32_t obs_stride, int32_t n_threads);

void orc_rollout_discrete(OrcGrid *grids, int64_t n, const int32_t *actions, int32_t n_steps,
                          const int8_t *plist_module, const int8_t *plist_action,
                          const int32_t *plist_offset /* per grid: row offset into LUT */,
                          int32_t plist_len, int order, double *rewards, uint8_t *dones,
                          double *obs_last, int32_t obs_stride, int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif
