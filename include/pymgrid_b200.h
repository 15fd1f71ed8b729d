/*
 * pymgrid_b200.h -- C-ABI of the B200-native batched microgrid-step engine.
 *
 * The reference (Total-RD/pymgrid @ 7bf3951, pure Python) has no FFI; its "operator API" for this path is the
 * Python call `Microgrid.run(control, normalized)` (src/pymgrid/microgrid/microgrid.py:227-325) dispatching to
 * `BaseMicrogridModule.step` (src/pymgrid/modules/base/base_module.py:95-159).  This header is the boundary a
 * maintainer would bind from Python (ctypes stub in INTEGRATION.md) to replace that loop for a batch of B
 * independent microgrids.  Each entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - plain C types only: raw DEVICE pointers, sizes, a cudaStream_t passed as void*; no torch types;
 *   - every array is owned by the caller (the Python host keeps them as torch device tensors); a handle owns
 *     only a copy of the layout metadata; nothing is allocated on the step path;
 *   - every call returns 0 on success or a negative MG_E_* code, never throws; mg_last_error() gives the text;
 *   - calls are asynchronous with respect to the host (work is enqueued on `stream`);
 *   - a handle is not re-entrant: one call at a time per handle (the reference object is not thread safe
 *     either); different handles are independent.
 *
 * All floating-point data is IEEE f64, the reference's arithmetic type; integers are int32 / uint8 / uint32.
 */
#ifndef PYMGRID_B200_H
#define PYMGRID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_ABI_VERSION 2
#define MG_MAX_GROUPS 8
#define MG_N_INFO 16
#define MG_PLIST_WIDTH 3
#define MG_N_LOG 24

/* return codes */
enum {
    MG_OK = 0,
    MG_E_INVALID = -1,     /* bad argument / layout            */
    MG_E_CUDA = -2,        /* a CUDA runtime call failed       */
    MG_E_UNSUPPORTED = -3  /* valid but outside the built path */
};

/* flat observation order: the reference flattens with gym.spaces.flatten
 * (envs/base/base.py:211-223); real gym sorts Dict keys alphabetically. */
enum {
    MG_OBS_GYM_SORTED = 0, /* battery, genset, grid, load, pv                                           */
    MG_OBS_CONTAINER = 1,  /* load, pv, genset, battery, grid (module listing order, module_container.py) */
    MG_OBS_GYM_SORTED_PV_FIRST = 2 /* PV, battery, genset, grid, load: MicrogridGenerator grids name the renewable
                                      module 'PV' (convert/convert.py), which sorts before the lower-case names */
};

/* columns of the optional per-env info block (the reference's info dict, microgrid/utils/step.py:22-31) */
enum {
    MG_INFO_LOAD_MET = 0, MG_INFO_PV_USED = 1, MG_INFO_CURTAILMENT = 2, MG_INFO_LOSS_LOAD = 3,
    MG_INFO_OVERGENERATION = 4, MG_INFO_GENSET_PRODUCTION = 5, MG_INFO_GENSET_CO2 = 6,
    MG_INFO_BATTERY_DISCHARGE = 7, MG_INFO_BATTERY_CHARGE = 8, MG_INFO_GRID_IMPORT = 9,
    MG_INFO_GRID_EXPORT = 10, MG_INFO_GRID_CO2 = 11,
    /* per-module rewards (the 'reward' column each module logs, base_module.py:276-290); load and pv are 0.0 */
    MG_INFO_REWARD_GENSET = 12, MG_INFO_REWARD_BATTERY = 13, MG_INFO_REWARD_GRID = 14, MG_INFO_REWARD_UNBALANCED = 15
};

/* per-env event flags: where the reference raises (or, with raise_errors=False, silently clips) */
enum {
    MG_FLAG_GENSET_GOAL_RANGE = 1u << 0, /* AssertionError genset_module.py:147                         */
    MG_FLAG_GENSET_AS_SINK = 1u << 1,    /* TypeError base_module.py:265 (max_consumption is NotImplemented) */
    MG_FLAG_BALANCE = 1u << 2,           /* RuntimeError   microgrid.py:321-323                         */
    MG_FLAG_BATTERY_MIN_CAP = 1u << 3,   /* AssertionError battery_module.py:128                        */
    MG_FLAG_NEGATIVE_ABSORB = 1u << 4,   /* AssertionError base_module.py:272; priority_list.py:124     */
    MG_FLAG_STEP_PAST_END = 1u << 5,     /* IndexError     load_module.py:111 (t >= len(series))        */
    MG_FLAG_BAD_ACTION = 1u << 6,        /* ValueError     envs/discrete/discrete.py:84 (action not in space) */
    MG_FLAG_SHAPER_RANGE = 1u << 7,      /* AssertionError reward_shaping/battery_discharge_shaper.py:33 (value outside [-1, 1]) */
    MG_FLAG_CLIP_GENSET = 1u << 8,       /* ValueError when raise_errors=True, base_module.py:213-221   */
    MG_FLAG_CLIP_BATTERY = 1u << 9,
    MG_FLAG_CLIP_GRID = 1u << 10,
    /* direction bits (not errors): which info key the reference would have written for this step */
    MG_FLAG_BATTERY_SINK = 1u << 12,     /* battery acted as a sink: 'absorbed_energy' (else 'provided_energy')      */
    MG_FLAG_GRID_SINK = 1u << 13,        /* grid exported: 'absorbed_energy'                                          */
    MG_FLAG_EXCESS = 1u << 14            /* microgrid.py:286 excess branch: unbalanced module absorbed (else provided) */
};

/* modules in a priority-list element (algos/priority_list/priority_list_element.py) */
enum { MG_MOD_NONE = -1, MG_MOD_GENSET = 0, MG_MOD_BATTERY = 1, MG_MOD_GRID = 2 };

/*
 * One microgrid parameter set ("config").  Device array, 336-byte stride, indexed by MgGroup.cfg_index.
 * Raw parameters are the reference constructors' arguments; the *_low / *_spread members are the
 * ModuleSpace constants the reference derives once at construction (utils/space.py:183-205), computed by the
 * host in f64 exactly as the reference does.
 */
typedef struct MgConfig {
    /* BatteryModule, modules/battery_module.py:66-91 */
    double bat_min_capacity, bat_max_capacity, bat_max_charge, bat_max_discharge, bat_efficiency, bat_cost_cycle;
    double bat_act_low, bat_act_spread;                      /* :332-338 (note the reference's min/max naming) */
    double bat_soc_low, bat_soc_spread, bat_charge_spread;   /* :323-330; charge low is bat_min_capacity      */
    /* GensetModule, modules/genset_module.py:61-92 */
    double gen_running_min, gen_running_max, gen_cost, gen_co2_per_unit, gen_cost_per_unit_co2;
    double gen_act_spread, gen_up_spread, gen_down_spread;   /* :503-517 */
    /* GridModule, modules/grid_module.py:70-132 */
    double grid_max_import, grid_max_export, grid_cost_per_unit_co2, grid_act_low, grid_act_spread;
    /* UnbalancedEnergyModule, modules/unbalanced_energy_module.py:14-26 */
    double loss_load_cost, overgeneration_cost;
    /* profile-times-scale series (MicrogridGenerator grids, MicrogridGenerator.py:137-148: ts = profile * (size/max)):
       series value = table value * scale.  With series_scaled != 0 the load / pv observation windows are normalised
       on the fly with these bounds ((v - low) / spread, rows past the end = *_fill_nrm) instead of read from the
       pre-normalised tables, which only exist per PROFILE.  scale == 1 and series_scaled == 0 for table-backed grids. */
    double load_scale, pv_scale, load_low, load_spread, pv_low, pv_spread, load_fill_nrm, pv_fill_nrm;
    int32_t gen_start_up_time, gen_wind_down_time, gen_allow_abortion;
    int32_t load_series, pv_series, grid_series;             /* rows of the series tables                     */
    int32_t initial_step, final_step;                        /* base_timeseries_module.py:317-330             */
    int32_t plist_offset, plist_count;                       /* rows of MgLayout.plist owned by this config   */
    int32_t series_scaled;                                   /* see load_scale                                */
    int32_t grid_status_weak;                                /* per-env status bits: 1 if any outage (obs bounds 0..1), 0 if all ones */
    int32_t reward_shaper;                                   /* MG_SHAPER_*: what `reward` carries (microgrid/utils/step.py:41-46) */
    int32_t reserved[3];
} MgConfig;

/* Microgrid(reward_shaping_func=...): the reference's two built-in shapers replace the step reward (the sum of module
 * rewards stays available through the info block's per-module rewards).
 *   PV_CURTAILMENT     reward = -1.0 * curtailment                       reward_shaping/pv_curtailment_shaper.py:16-18
 *   BATTERY_DISCHARGE  reward = (battery discharge - loss load) / load   reward_shaping/battery_discharge_shaper.py:23-35
 *                      (IEEE division; MG_FLAG_SHAPER_RANGE where the reference's assert on [-1, 1] fires, nan included) */
enum { MG_SHAPER_NONE = 0, MG_SHAPER_PV_CURTAILMENT = 1, MG_SHAPER_BATTERY_DISCHARGE = 2 };

/*
 * One priority list = one discrete action (algos/priority_list/priority_list.py:15-67): up to MG_PLIST_WIDTH
 * elements (module, genset-goal), deployed in order by mg_step_discrete.
 */
typedef struct MgPriorityList {
    int8_t module[MG_PLIST_WIDTH];   /* MG_MOD_* ; MG_MOD_NONE pads                                           */
    int8_t action[MG_PLIST_WIDTH];   /* genset: the goal status (0/1); others 0                               */
    int8_t n_elements, _pad;
} MgPriorityList;

/*
 * Architecture group: envs that share (has_genset, has_grid, horizon) and therefore the observation / action
 * row layout.  pymgrid25 has three (genset-only, grid-only, genset+grid).  State arrays are read AND written.
 */
typedef struct MgGroup {
    int32_t has_genset, has_grid, horizon, obs_order;
    int32_t n_act, obs_dim;          /* must equal 1+has_grid+2*has_genset, (1+H)(2+4*has_grid)+2+4*has_genset */
    int64_t n_envs;
    int32_t act_col_genset, act_col_battery, act_col_grid, _pad;  /* columns of the action row               */
    /* per-env state (the reference's serialisable state: base_module.py:852-868, genset_module.py:426-427) */
    int32_t *step;                   /* [n] _current_step                                                     */
    double *charge;                  /* [n] battery _current_charge (soc is derived)                          */
    uint32_t *genset;                /* [n] cs | gs<<8 | steps_until_up<<16 | steps_until_down<<24; NULL if no genset */
    const int32_t *cfg_index;        /* [n] row of MgLayout.cfg                                               */
    const int32_t *env_initial_step; /* [n] per-env trajectory window (microgrid/trajectory/), or NULL -> cfg  */
    const int32_t *env_final_step;   /* [n] or NULL -> cfg                                                    */
    /* optional per-env grid status (weak grids, MicrogridGenerator.py:321-340): bit t of row e = grid_status[t];
       [n][status_words] uint32, rows cover T + max_horizon + 1 bits (bits >= T unused).  NULL -> column 3 of grid_raw */
    const uint32_t *grid_status_bits;
    int64_t status_words;
} MgGroup;

/* MgLayout.flags */
#define MG_LAYOUT_SCALED_SERIES 1    /* some MgConfig has series_scaled != 0: selects the kernels with the per-env series path */
#define MG_LAYOUT_OBS_F32 2          /* every obs / obs_ring buffer is float32 [n, obs_dim] (the f64 observation rounded to
                                        nearest); halves the dominant traffic for consumers that feed fp32 policies.
                                        State, actions, reward and all arithmetic stay f64. */

typedef struct MgLayout {
    int32_t abi_version;
    int32_t n_groups;
    MgGroup groups[MG_MAX_GROUPS];
    int32_t n_cfg;
    int32_t series_len;              /* T: rows of every raw series                                           */
    int32_t max_horizon;             /* normalised tables carry T + max_horizon + 1 rows                      */
    int32_t n_load, n_pv, n_grid;
    const MgConfig *cfg;             /* [n_cfg]                                                               */
    const double *load_raw;          /* [n_load][T]    stored NEGATIVE like the reference (base_timeseries_module.py:68-79) */
    const double *pv_raw;            /* [n_pv][T]                                                             */
    const double *grid_raw;          /* [n_grid][T][4] import_price, export_price, co2_per_kwh, grid_status   */
    /* normalised, end-padded observation tables, FILLED BY mg_create on the device:
       value = (ts - low) / spread (utils/space.py:207-218) with the bounds of base_timeseries_module.py:81-88 /
       grid_module.py:125-132; rows >= T hold the forecaster's fill (high+low)/2 (forecast/forecaster.py:95) */
    double *load_nrm;                /* [n_load][T + max_horizon + 1]                                         */
    double *pv_nrm;                  /* [n_pv][T + max_horizon + 1]                                           */
    double *grid_nrm;                /* [n_grid][T + max_horizon + 1][4]                                      */
    double *bounds;                  /* [(n_load + n_pv + 4*n_grid)][2] low, high of every series column (out) */
    const MgPriorityList *plist;     /* [n_plist] or NULL when mg_step_discrete is not used                   */
    int32_t n_plist;
    int32_t flags;                   /* MG_LAYOUT_*                                                           */
} MgLayout;

/* per-group arguments of one step */
typedef struct MgStepIO {
    const double *actions;   /* [n, n_act] f64 (float32 under MG_OPT_ACTIONS_F32): normalised in [0,1] or unnormalised */
    const int32_t *dactions; /* [n] int32 priority-list index (mg_step_discrete)                              */
    double *obs;             /* [n, obs_dim] normalised post-step observation (float* with MG_LAYOUT_OBS_F32), or NULL to skip */
    double *reward;          /* [n]                                                                           */
    uint8_t *done;           /* [n]                                                                           */
    double *info;            /* [n, MG_N_INFO] or NULL                                                        */
    uint32_t *flags;         /* [n] MG_FLAG_* or NULL                                                         */
    const uint8_t *mask;     /* [n] mg_reset only: envs to reset (NULL = all)                                 */
    double *reward_total;    /* [1] or NULL: += sum of this step's rewards over the group (logging aggregate: warp-shuffle
                                reduction + one atomicAdd per warp; the caller zeroes it; summation order is not fixed) */
} MgStepIO;

/* per-group arguments of a multi-step rollout: leading dimension is the step */
typedef struct MgRolloutIO {
    const double *actions;   /* [n_steps, n, n_act] (mg_rollout); float32 under MG_OPT_ACTIONS_F32             */
    const int32_t *dactions; /* [n_steps, n]        (mg_rollout_discrete); [n] when dactions_const != 0         */
    double *obs_ring;        /* [ring, n, obs_dim]: step s writes slot s % ring; NULL to skip observations    */
    double *reward;          /* [n_steps, n]                                                                  */
    uint8_t *done;           /* [n_steps, n]                                                                  */
    double *reward_sum;      /* [n] sum over the rollout in step order, or NULL                               */
    uint32_t *flags;         /* [n] OR over the rollout, or NULL                                              */
    int64_t dactions_const;  /* != 0: the same priority list every step -- rule-based control (algos/rbc/rbc.py:64-93) */
    double *reward_total;    /* [n_steps] or NULL: [s] += sum over the group of step s's rewards (see MgStepIO)        */
    /* Per-step log of SELECTED envs, written inside the persistent kernel (the reference's Microgrid.get_log, microgrid.py:
       434-475, and ModularLogger, utils/logger.py:18-28; a full log of every env is 169-176 columns per env-step and cannot
       be always-on at batch scale).  log_slot: [n] int32, -1 = not logged, else the env's row r of `log`;
       log: [n_logged, n_steps, MG_N_LOG] f64, columns MG_LOG_* -- what the host needs to rebuild the reference's log row:
       the state BEFORE the step (step counter, battery charge, packed genset status), the genset status AFTER it (the
       reference logs that one, genset_module.py:148-149), reward, done, flags and the MG_N_INFO info columns.  NULL = no log. */
    const int32_t *log_slot;
    double *log;
} MgRolloutIO;

/* columns of one log record (MgRolloutIO.log) */
enum {
    MG_LOG_STEP = 0, MG_LOG_CHARGE = 1, MG_LOG_GENSET_BEFORE = 2, MG_LOG_GENSET_AFTER = 3, MG_LOG_REWARD = 4, MG_LOG_DONE = 5,
    MG_LOG_FLAGS = 6, /* 7 reserved */ MG_LOG_INFO = 8 /* .. 23: MG_INFO_* */
};

typedef struct MgHandle MgHandle;

/* library / build identification (also the symbol the loader probes first) */
int mg_abi_version(void);
/* sizeof of an ABI struct as compiled (0 MgConfig, 1 MgPriorityList, 2 MgGroup, 3 MgLayout, 4 MgStepIO,
 * 5 MgRolloutIO, 6 MgForecastNoise, 7 MgHostRolloutIO): lets a foreign-language binding verify its struct mirror before the first call */
int64_t mg_sizeof(int which);
const char *mg_build_info(void);
const char *mg_last_error(void);

/*
 * mg_create -- replaces Microgrid.__init__ / from_scenario for a batch (microgrid.py:100-128, 958-980):
 * validates the layout, keeps a device copy of the group table, and fills the normalised observation tables
 * and `bounds` on `stream` (the reference computes the same bounds in each module's constructor).
 */
int mg_create(const MgLayout *layout, void *stream, MgHandle **out);
int mg_destroy(MgHandle *h);

/*
 * mg_step -- replaces Microgrid.run(control, normalized) (microgrid.py:227-325) and, with obs != NULL,
 * BaseMicrogridEnv.step (envs/base/base.py:169-209) for every env of every group, ONE fused kernel launch.
 * io: array of n_groups.  normalized: same meaning as the reference argument.
 */
int mg_step(MgHandle *h, const MgStepIO *io, int normalized, void *stream);

/*
 * mg_step_discrete -- replaces DiscreteMicrogridEnv.step(action:int) (envs/discrete/discrete.py:109-143):
 * priority-list expansion (algos/priority_list/priority_list.py:69-116) fused in front of the step.
 */
int mg_step_discrete(MgHandle *h, const MgStepIO *io, void *stream);

/*
 * mg_reset -- replaces Microgrid.reset / BaseMicrogridEnv.reset (microgrid.py:205-225, envs/base/base.py:165-167):
 * step = initial_step for the masked envs (battery charge and genset status are NOT reset, as in the
 * reference), then writes the observation of every env of the group to io[g].obs when it is not NULL.
 */
int mg_reset(MgHandle *h, const MgStepIO *io, void *stream);

/* mg_observe -- current normalised observation without stepping (BaseMicrogridModule.state normalised,
 * base_module.py:65-77, 157) */
int mg_observe(MgHandle *h, const MgStepIO *io, void *stream);

/*
 * mg_rollout / mg_rollout_discrete -- n_steps consecutive mg_step / mg_step_discrete calls in ONE persistent
 * kernel: state stays on chip between steps; replaces the caller's `for t in range(T): microgrid.run(...)`
 * loop (e.g. algos/rbc/rbc.py:87-91).  ring = number of observation slots (>= 1).
 */
int mg_rollout(MgHandle *h, const MgRolloutIO *io, int32_t n_steps, int32_t ring, int normalized, void *stream);
int mg_rollout_discrete(MgHandle *h, const MgRolloutIO *io, int32_t n_steps, int32_t ring, void *stream);

/*
 * mg_rollout_host -- the same loop for callers whose actions and results live in HOST memory, which is where every
 * caller of the reference keeps them (the control dicts of Microgrid.run, microgrid.py:227-251; the numpy action arrays
 * of algos/rbc/rbc.py:87-91 and of the RL notebooks).  The rollout is cut into chunks of `chunk` steps and pipelined on
 * three streams: chunk c+1's actions go host -> device on a copy-in stream while the persistent kernel runs chunk c on
 * `stream` and chunk c-1's reward / done go device -> host on a copy-out stream, so the bus, not the sum of the three,
 * sets the pace.  The handle owns the double-buffered device staging (2 x chunk steps of actions, reward, done; allocated
 * on first use on the current device, released by mg_destroy) -- the only device memory this library ever allocates.
 * Host buffers should be page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) for the copies to overlap;
 * pageable memory gives the same results without the overlap.  The call returns once everything is enqueued; the results
 * are complete when `stream` has drained (it is made to wait for the last copy-out).  Observations go to the caller's
 * DEVICE ring (every chunk restarts at slot 0: step s of a chunk writes slot s % ring).
 */
typedef struct MgHostRolloutIO {
    const double *actions;   /* HOST [n_steps, n, n_act] f64 (discrete == 0); float32 under MG_OPT_ACTIONS_F32 */
    const int32_t *dactions; /* HOST [n_steps, n] int32 priority-list index (discrete != 0)                   */
    double *reward;          /* HOST [n_steps, n]                                                             */
    uint8_t *done;           /* HOST [n_steps, n]                                                             */
    void *obs_ring;          /* DEVICE [ring, n, obs_dim] or NULL to skip observations                        */
    uint32_t *flags;         /* DEVICE [n] OR over the rollout, or NULL                                       */
} MgHostRolloutIO;

int mg_rollout_host(MgHandle *h, const MgHostRolloutIO *io, int32_t n_steps, int32_t chunk, int32_t ring, int discrete,
                    int normalized, void *stream);

/*
 * mg_forecast_noise -- GaussianNoiseForecaster (forecast/forecaster.py:220-262) applied to observation rows that
 * mg_step / mg_step_discrete / mg_reset / mg_observe have just written on the same stream.
 *
 * The reference adds N(0, std_k) to every REAL row k of a module's forecast window (rows past the end of the series are
 * padded afterwards and carry no noise, :120-132), clips to the column's bounds (:139-149) and normalises.  In normalised
 * units that is  obs <- min(max(obs + z * sigma * scale_k, 0), 1)  with sigma = std / (high - low) of the column and
 * scale_k = 1 + log(1 + k) when increase_uncertainty is set (:244-248), 1 otherwise.  The host fills MgForecastNoise per
 * config: relative_noise (std * |mean(series)|, :239-242) is already folded into sigma, and sigma = 0 for a constant
 * column (the clip pins it to its bound) and for modules with the oracle forecaster.  The current values (the first row of
 * every module block), the battery / genset entries and the physics are untouched.
 *
 * z is standard normal (Box-Muller over Philox4x32-10), a pure function of (seed, call, global env id, the env's step,
 * element): reproducible, independent of the launch shape, and independent between envs, steps and calls -- where the
 * reference draws from numpy's global generator.  Parity with the reference is therefore distributional.
 * `env_base[g]` is the global id of group g's first env (group slots are consecutive ids), so that shards of one batch
 * on different GPUs draw different noise.  `obs[g]` may be NULL to skip a group.
 */
typedef struct MgForecastNoise {
    double load_sigma, pv_sigma, grid_sigma[4];              /* normalised units; 0 = leave the column alone */
    int32_t load_increase, pv_increase, grid_increase, _pad; /* increase_uncertainty per module               */
} MgForecastNoise;

int mg_forecast_noise(MgHandle *h, const MgForecastNoise *noise /* DEVICE [n_cfg] */, void *const *obs /* [n_groups] */,
                      const int64_t *env_base /* HOST [n_groups] or NULL -> 0, n_0, n_0 + n_1, ... */,
                      uint64_t seed, uint64_t call, void *stream);
/* The same for rows that were written at an EARLIER step than the env's current one -- the slots of mg_rollout's observation
 * ring: the row of group g's env e observes step min(step_base[g][e] + step_add, T), with step_base the env's step counter
 * before the rollout (DEVICE [n] int32 per group, HOST array of n_groups pointers; a NULL entry skips the group) and step_add
 * the number of steps taken when the slot was written.  With call numbers that continue the per-step sequence, a rollout's
 * ring then holds exactly the noisy rows that mg_step + mg_forecast_noise would have produced step by step. */
int mg_forecast_noise_at(MgHandle *h, const MgForecastNoise *noise, void *const *obs, const int64_t *env_base, uint64_t seed,
                         uint64_t call, const int32_t *const *step_base, int32_t step_add, void *stream);

/* Tuning knobs.  MG_OPT_ROLLOUT_SPECIALISED (default 1): run mg_rollout with the owner / emitter warp-specialised
 * kernel when every group writes observations.  It wins when the envs of a tile advance in lock-step (11.5 vs 12.4
 * us/step at 65 536 envs) and loses when every env is at its own step (39.6 vs 29 us/step): hosts that install per-env
 * trajectory windows turn it off. */
enum { MG_OPT_ROLLOUT_SPECIALISED = 1, MG_OPT_ROLLOUT_RING = 2, MG_OPT_EMIT_IMAGE = 3, MG_OPT_IMAGE_SHAPE = 4, MG_OPT_RAGGED_HINT = 5, MG_OPT_STEP_OVERLAP = 6,
       MG_OPT_ACTIONS_F32 = 7 };
/* MG_OPT_ROLLOUT_RING (default 1): batches with per-env series (MG_LAYOUT_SCALED_SERIES / grid_status_bits) run mg_rollout
 * with every env's normalised load / pv windows held in shared memory (H + 2 slots per env and series; one new value per
 * env, series and step instead of a whole window per row) when all horizons are <= 24.  0 selects the kernel that
 * normalises whole windows per row (the one mg_step uses). */
/* MG_OPT_EMIT_IMAGE: how observation rows leave the SM.  1 = assembled in shared-memory images and stored with TMA bulk
 * stores (cp.async.bulk shared -> global, several rows per store; needs f64 rows, 1 + H <= 32 and a grid block that starts
 * at an even element, else the other emitters run); 0 = per-lane 16-byte stores with run detection (rows of a tile that
 * read the same windows fetch them once); 2 (default) = the library chooses per launch from what was measured: images for
 * per-env series, for envs at unrelated steps (MG_OPT_RAGGED_HINT), for batches of at most two tiles per multiprocessor and
 * for persistent launches over rows with an even forecast horizon; per-lane stores for large table-backed batches in
 * lock-step.  Results never depend on the choice.
 * MG_OPT_IMAGE_SHAPE (default -1 = the library's choice): which instantiated (rows per bulk store, image buffers per
 * emitting warp, rows gathered together) shape the image kernels use -- 0: (4, 2, 2), 1: (2, 2, 2), 2: (4, 2, 4),
 * 3: (4, 4, 2), 4: (4, 4, 4), 5: (8, 2, 4); 3-5 run the per-env-series persistent kernel at three CTAs per SM with 168
 * registers.  A tuning knob.
 * MG_OPT_RAGGED_HINT (default 0): tell the library that the envs of a tile are at unrelated steps (independent resets,
 * per-env episode windows), where no two rows share a window.
 * MG_OPT_STEP_OVERLAP (default 1): chain consecutive mg_step* launches on one stream with programmatic dependent launch
 * (griddepcontrol): a launch starts as soon as every CTA of the previous one has written and fenced the state it hands on,
 * so its latency-bound part (state, inputs, physics) runs under the previous launch's observation stream.  1: its own rows
 * wait for the previous launch to complete (safe for any choice of observation buffers); 2: the row streams overlap too
 * (only launches whose observation buffers differ from those of the two launches before are chained).  Anything else
 * enqueued between two steps (a policy's kernels, copies) simply breaks the chain: ordering is the stream's, as always.
 * MG_OPT_ACTIONS_F32 (default 0): every `actions` pointer given to mg_step / mg_rollout / mg_rollout_host afterwards holds
 * float32 values in the same shape (a policy network's output as it is; half the bytes over PCIe for mg_rollout_host).  Each
 * value is widened to f64 exactly and the step computes what the reference computes for np.float64(action): all arithmetic
 * stays f64.  Rows of four must be 16-byte aligned, rows of two 8-byte. */
int mg_set_option(MgHandle *h, int option, int value);

/*
 * mg_set_reported_soc -- BatteryModule._soc before the first update (battery_module.py:89, 96-106).  The reference keeps
 * the soc a battery was CONSTRUCTED with (init_soc) and only recomputes it as current_charge / max_capacity in
 * _update_state (:125-130); init_soc * max_capacity / max_capacity is not always init_soc in the last bit, so the
 * observation Microgrid.reset() returns before any step can differ from charge / max_capacity by one ulp.
 * `soc`: HOST array of n_groups DEVICE pointers ([n] f64 each; an entry or the whole array may be NULL = derive from the
 * charge).  mg_observe / mg_reset report these values as the battery's soc until the first mg_step* / mg_rollout* call
 * on the handle, which drops them (every battery updates in every step).  The arrays stay owned by the caller and must
 * outlive that first step.  Callers that overwrite the charge array themselves (BatteryModule.current_charge setter,
 * :360-362) pass NULL to drop them.
 */
int mg_set_reported_soc(MgHandle *h, const double *const *soc);

/*
 * mg_set_trajectories -- per-env episode windows (microgrid/trajectory/*.py; Microgrid._set_trajectory, microgrid.py:221-225,
 * which rewrites every module's initial_step / final_step on reset).  `initial_step` / `final_step`: HOST arrays of n_groups
 * DEVICE pointers ([n] int32 each), replacing MgGroup.env_initial_step / env_final_step of every group for the calls that
 * follow; an entry, or a whole array, may be NULL = the configs' own window.  The handle stays the same object: launchers
 * bound earlier, options and the staging of mg_rollout_host are unaffected.  The arrays stay owned by the caller.
 */
int mg_set_trajectories(MgHandle *h, const int32_t *const *initial_step, const int32_t *const *final_step);

/* number of kernel launches this handle has enqueued since creation (bench.py's gpu_launches claim) */
int64_t mg_launch_count(const MgHandle *h);
/* name of the kernel family the last mg_step* / mg_rollout* call on the handle launched (which emitters MG_OPT_EMIT_IMAGE = 2 chose) */
const char *mg_last_kernel(const MgHandle *h);

#ifdef __cplusplus
}
#endif
#endif /* PYMGRID_B200_H */
