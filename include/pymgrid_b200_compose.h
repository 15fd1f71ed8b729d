/*
 * pymgrid_b200_compose.h -- C-ABI of the COMPOSED-microgrid step: `Microgrid.run` for ANY module list.
 *
 * pymgrid_b200.h covers the module set of every pymgrid25 / MicrogridGenerator grid (one load, one renewable, one battery,
 * at most one genset and one grid) with kernels specialised for it.  The reference's `Microgrid.run`
 * (src/pymgrid/microgrid/microgrid.py:227-325) dispatches over any list of modules -- several loads and renewables (the
 * reference's own balance tests, tests/microgrid/test_microgrid.py:188-455), several batteries / gensets / grids, no
 * battery, no slack module, one forecast horizon per time-series module.  The entry points below run that general
 * dispatch for a batch of B microgrids that share one COMPOSITION (the module list: kinds, order, horizons) and differ in
 * parameters, series and state.  Same conventions as pymgrid_b200.h: plain C types, raw DEVICE pointers owned by the
 * caller, int return codes (0 ok, MG_E_*), mg_last_error() for the text, asynchronous on `stream`.
 *
 * Per-env event flags are the MG_FLAG_* bits of pymgrid_b200.h where they apply, plus MGC_FLAG_* below.
 */
#ifndef PYMGRID_B200_COMPOSE_H
#define PYMGRID_B200_COMPOSE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGC_ABI_VERSION 1
#define MGC_MAX_MODULES 64     /* modules per microgrid (np.sum's unrolled pairwise order is restated up to 128 addends) */
#define MGC_INFO_SLOTS 5       /* per module: provided_energy, absorbed_energy, co2_production | curtailment, reward, acted as sink */
#define MGC_BALANCE_SLOTS 6    /* fixed provided / absorbed, controllable provided / absorbed, overall provided / absorbed */
#define MGC_CFG_HEADER 2       /* doubles in front of the module blocks of a config record: initial_step, final_step */

/* module kinds (reference classes under src/pymgrid/modules/) and the doubles each takes in a config record */
enum {
    MGC_LOAD = 0,       /* load_module.py        fixed sink:   series_index, low, high                                      */
    MGC_RENEWABLE = 1,  /* renewable_module.py   flex source:  series_index, low, high                                      */
    MGC_BATTERY = 2,    /* battery_module.py     controllable: min_capacity, max_capacity, max_charge, max_discharge,
                                                               efficiency, battery_cost_cycle                               */
    MGC_GENSET = 3,     /* genset_module.py      controllable: running_min_production, running_max_production, genset_cost,
                                                               co2_per_unit, cost_per_unit_co2, start_up_time,
                                                               wind_down_time, allow_abortion                               */
    MGC_GRID = 4,       /* grid_module.py        controllable: series_index, max_import, max_export, cost_per_unit_co2,
                                                               low[4], high[4] (per-column series bounds)                   */
    MGC_UNBALANCED = 5  /* unbalanced_energy_module.py flex:   loss_load_cost, overgeneration_cost                          */
};
#define MGC_N_KINDS 6

/* additional per-env flag bits (beside MG_FLAG_*) */
#define MGC_FLAG_CLIP 0x100u            /* some module clipped the request to its limits (base_module.py:213-224, 265-270)   */
#define MGC_FLAG_CLIP_RAISES 0x800u     /* ... and that module was built with raise_errors=True: ValueError, :79-93          */
#define MGC_FLAG_NOT_A_SINK 0x2u        /* a source-only module was asked to absorb: TypeError, base_module.py:265 (= MG_FLAG_GENSET_AS_SINK) */

/*
 * One module of the composition.  The array is in DISPATCH order -- fixed modules, then controllable, then flex, each in
 * the container's order (module_container.py:355-413) -- which is the order Microgrid.run steps them in and the order
 * their rewards and energies are summed in.  Integers stored in a config record (series_index, start_up_time, ...) are
 * exact in a double.
 */
typedef struct MgcModule {
    int32_t kind;          /* MGC_*                                                                                     */
    int32_t horizon;       /* forecast rows of a time-series module (0 = no forecaster)                                 */
    int32_t act_col;       /* first column of the action row (battery, grid: 1 column; genset: goal, energy), else -1   */
    int32_t obs_off;       /* first element of the module's block in the flat observation row                           */
    int32_t param_off;     /* first double of the module's parameters in a config record                                */
    int32_t fstate_off;    /* battery: current_charge, soc -> 2 doubles of the env's fstate row; else -1                */
    int32_t istate_off;    /* genset: current_status, goal_status, steps_until_up, steps_until_down -> 4 int32; else -1 */
    int32_t listing;       /* position in the container's LISTING order (fixed, flex, controllable) = row of `info`     */
    int32_t raise_errors;  /* the module's raise_errors argument: a clip sets MGC_FLAG_CLIP_RAISES                      */
} MgcModule;

typedef struct MgcLayout {
    int32_t abi_version;
    int32_t n_modules;
    MgcModule modules[MGC_MAX_MODULES];
    int32_t n_act, obs_dim;          /* widths of the action and observation rows                                       */
    int32_t n_fstate, n_istate;      /* widths of the per-env state rows                                                */
    int32_t cfg_stride;              /* doubles per config record                                                       */
    int32_t n_cfg;
    int32_t series_len;              /* T: rows of every series                                                         */
    int32_t n_series;
    int64_t n_envs;
    const double *cfg;               /* [n_cfg][cfg_stride]                                                             */
    const double *series;            /* pool; series i starts at series_off[i], row-major [T][C], C = 4 (grid) or 1;
                                        load series are stored NEGATIVE like the reference (base_timeseries_module.py:68-79) */
    const int64_t *series_off;       /* [n_series] element offsets into `series`                                        */
    const double *series_nrm;        /* the same pool NORMALISED per column, (ts - low) / spread with the bounds of the module
                                        kind that owns the series (base_timeseries_module.py:81-88, grid_module.py:125-132,
                                        utils/space.py:207-218), computed once by the host in f64: observation rows are then
                                        pure gathers (series values lie inside their own bounds, so the forecaster's clip,
                                        forecast/forecaster.py:139-149, never moves them).  NULL -> normalise on the fly. */
    /* per-env state, read AND written */
    int32_t *step;                   /* [n] _current_step (all modules of a microgrid share it)                         */
    double *fstate;                  /* [n][n_fstate]                                                                   */
    int32_t *istate;                 /* [n][n_istate]                                                                   */
    const int32_t *cfg_index;        /* [n] row of `cfg`                                                                */
    /* discrete actions (mgc_run_discrete): the priority lists of the composition (algos/priority_list/priority_list.py:15-67),
       [n_plist][plist_width][2] int16 = (index of a controllable module in `modules`, action number: the goal of a genset,
       0 otherwise); a negative module pads.  NULL / 0 when mgc_run_discrete is not used. */
    const int16_t *plist;
    int32_t n_plist, plist_width;
    /* optional per-env episode windows (microgrid/trajectory/: trajectory_func on reset, microgrid.py:221-225):
       [n] each, read at every step / reset, so the caller may rewrite them between calls; NULL -> the config record's
       initial_step / final_step */
    const int32_t *env_initial_step;
    const int32_t *env_final_step;
    /* optional observation selection (BaseMicrogridEnv(observation_keys=...), envs/base/base.py:109-163, 211-223): HOST array
       [obs_dim], entry j = (index of the module in `modules`) << 16 | element of that module's observation block.  The row
       then holds exactly these elements in this order and the modules' obs_off are ignored.  NULL -> every module block at
       its obs_off. */
    const int32_t *obs_select;
} MgcLayout;

typedef struct MgcIO {
    const double *actions;  /* [n_steps, n, n_act] normalised in [0,1] or unnormalised; may be NULL when n_act == 0      */
    double *obs;            /* [ring, n, obs_dim]: step s writes slot s % ring; NULL to skip                             */
    double *reward;         /* [n_steps, n]                                                                              */
    uint8_t *done;          /* [n_steps, n]                                                                              */
    double *info;           /* [n, n_modules * MGC_INFO_SLOTS + MGC_BALANCE_SLOTS] of the LAST step, or NULL             */
    uint32_t *flags;        /* [n] OR over the steps, or NULL                                                            */
    const uint8_t *mask;    /* mgc_reset only: envs to reset (NULL = all)                                                */
    const int32_t *dactions; /* mgc_run_discrete: [n_steps, n] priority-list index; [n] when dactions_const != 0           */
    int64_t dactions_const;  /* != 0: the same list every step -- rule-based control (algos/rbc/rbc.py:64-93)             */
} MgcIO;

typedef struct MgcHandle MgcHandle;

int mgc_abi_version(void);
int64_t mgc_sizeof(int which);     /* 0 MgcModule, 1 MgcLayout, 2 MgcIO */
int32_t mgc_param_count(int kind); /* doubles a module of this kind takes in a config record */

/* mgc_create -- replaces Microgrid.__init__ for a batch (microgrid.py:100-165): validates the composition (dispatch order,
 * blocks inside their rows and not overlapping, listing a permutation) and uploads one small table to the CURRENT device
 * (obs_dim int32: observation element -> module, offset; with series_nrm also its gather form, 3 more int32 per element)
 * -- the only device memory a handle owns, released by mgc_destroy.  Every other array stays owned by the caller.
 * With series_nrm given, observation rows are written by the staged-gather emitters (mg_compose.cu: per-env window bases and
 * state elements staged in shared memory by the thread that stepped the env, <= 96 KB per 128-env tile; layouts that need
 * more keep the per-element decode).  The environment variable PYMGRID_B200_COMPOSE_GATHER=0, read here, forces the
 * per-element decode: same bytes, kept for A/B measurements and the parity test between the two. */
int mgc_create(const MgcLayout *layout, MgcHandle **out);
int mgc_destroy(MgcHandle *h);

/*
 * mgc_run -- n_steps consecutive Microgrid.run(control, normalized) calls (microgrid.py:227-325) for every env, one
 * kernel launch: fixed modules, controllable modules with the caller's actions, the energy balance, flex modules
 * (microgrid.py:286-314), reward / done / info aggregation (microgrid/utils/step.py) and the post-step normalised
 * observation of every module (base_module.py:157).  n_steps = 1 is one Microgrid.run.
 */
int mgc_run(MgcHandle *h, const MgcIO *io, int32_t n_steps, int32_t ring, int normalized, void *stream);
/*
 * mgc_run_discrete -- DiscreteMicrogridEnv.step(action) (envs/discrete/discrete.py:109-143) / RuleBasedControl.run
 * (algos/rbc/rbc.py:64-93): io->dactions picks a priority list per env and step, its expansion into controls
 * (priority_list.py:69-116) is fused in front of the step.  An index outside the table sets MG_FLAG_BAD_ACTION and leaves
 * the env untouched (reward NaN).
 */
int mgc_run_discrete(MgcHandle *h, const MgcIO *io, int32_t n_steps, int32_t ring, void *stream);
/*
 * mgc_modules_step -- BaseMicrogridModule.step(action, normalized) (modules/base/base_module.py:95-159) for every module of
 * the composition on its own: the reference's operator API without a Microgrid around it.  io->actions: [n, W] with one
 * column per module in dispatch order (load: none; renewable, battery, grid, unbalanced: one; genset: goal, energy).  No
 * energy balance: the flex modules act on THEIR action like everyone else.  reward = sum of the modules' rewards, `info` the
 * per-module slots; one step per call.
 */
int mgc_modules_step(MgcHandle *h, const MgcIO *io, int normalized, void *stream);
/* mgc_reset -- Microgrid.reset (microgrid.py:205-225): step = initial_step for the masked envs; battery and genset
 * state stay; writes every env's observation when io->obs is not NULL. */
int mgc_reset(MgcHandle *h, const MgcIO *io, void *stream);
/* mgc_observe -- the current normalised observation without stepping. */
int mgc_observe(MgcHandle *h, const MgcIO *io, void *stream);
/*
 * mgc_forecast_noise -- GaussianNoiseForecaster (forecast/forecaster.py:220-262) applied to observation rows that mgc_run /
 * mgc_run_discrete (the last step's slot) / mgc_reset / mgc_observe have just written on the same stream: Gaussian noise on
 * every real forecast row of the time-series modules, clipped to the bounds; current values, battery / genset entries and
 * the physics are untouched.  `noise`: DEVICE [n_cfg][2 * obs_dim] -- per config, per row element, the normalised standard
 * deviation (std, times |mean(series)| under relative_noise, divided by the column's spread; 0 = no noise) followed by the
 * increase_uncertainty flags (0 / 1).  The draw is a pure function of (seed, call, env_base + env, the env's step, element):
 * reproducible and independent of the launch shape -- parity with the reference's global numpy generator is distributional.
 */
int mgc_forecast_noise(MgcHandle *h, const double *noise, double *obs, int64_t env_base, uint64_t seed, uint64_t call, void *stream);
int64_t mgc_launch_count(const MgcHandle *h);

#ifdef __cplusplus
}
#endif
#endif /* PYMGRID_B200_COMPOSE_H */
