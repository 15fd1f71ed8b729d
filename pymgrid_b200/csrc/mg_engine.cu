// mg_engine.cu -- B200 (sm_100a) batched microgrid-step engine: kernels + the C-ABI of include/pymgrid_b200.h.
//
// One fused kernel advances every env of every architecture group one timestep
// (reference: Microgrid.run, src/pymgrid/microgrid/microgrid.py:227-325):
//   phase 1  one thread per env: gather ts[t], genset state machine + clamp, battery transition, grid
//            import/export, energy balance -> pv -> unbalanced, reward, done, state write-back;
//   phase 2  all warps of the CTA emit the normalised observation rows of the tile with coalesced 16-byte
//            stores; the time-series part of a row is a window [t+1, t+1+H] of pre-normalised, end-padded
//            tables (built once by mg_create), the battery/genset part comes from shared memory.
// The same body runs inside a persistent multi-step kernel (mg_rollout) with the env state held in registers.
//
// Compiled with -fmad=false: the reference is plain IEEE f64 Python arithmetic with no fused multiply-add, and
// clip / snap decisions must be bit-identical (tests/test_gpu_parity.py).  The path is HBM-write bound
// (obs rows are >= 95% of the bytes), not FP64 bound, so this costs nothing measurable.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <new>

#include "pymgrid_b200.h"

#ifndef MG_THREADS
#define MG_THREADS 128
#endif
#ifndef MG_TILE
#define MG_TILE 64          // envs per CTA tile
#endif
// (debug builds, -DMG_DEBUG_NOSTORE: the image emitters do everything but issue their bulk stores -- what the kernels cost
//  without the write traffic; results are then meaningless)
#ifdef MG_DEBUG_NOSTORE
#define MG_BULK_STORE_ENABLED 0
#else
#define MG_BULK_STORE_ENABLED 1
#endif
#define MG_STR2(x) #x
#define MG_STR(x) MG_STR2(x)
#define MG_WARPS (MG_THREADS / 32)
#define MG_ROWS_PER_WARP (MG_TILE / MG_WARPS)
#define MG_MAX_IMG 192      // longest observation row (in f64) on the staged path (3 pairs per lane); longer rows use the element-wise path
#ifndef MG_ROLLOUT_WS
#define MG_ROLLOUT_WS 1     // persistent kernel: owner warps / emitter warps pipeline (see mg_rollout_ws_kernel)
#endif
#ifndef MG_MIN_CTAS_HETERO
#define MG_MIN_CTAS_HETERO 7
#endif
#define MG_MIN_CTAS 7       // resident CTAs per SM the register budget must allow (65 536 envs = 1 024 tiles = one wave)
#ifndef MG_TMA_MIN_RUN
#define MG_TMA_MIN_RUN 2
#endif
// persistent kernel for per-env series: sliding load / pv windows live in shared memory, H + 2 slots per env and series
#define MG_RING_MAX 26      // forecast horizons up to 24
#define MG_MIN_CTAS_RING 5  // 40 KB of shared memory per CTA
#define MG_N_IMAGE_SHAPES 6 // (rows per bulk store, buffers) shapes of the image-emitter kernels, see img_kernel_for
// runs of at least this many rows stage their grid window in shared memory with TMA

enum { KIND_BAT = 0, KIND_GEN = 1, KIND_GRID = 2, KIND_LOAD = 3, KIND_PV = 4 };
enum { MODE_STEP = 0, MODE_DISCRETE = 1, MODE_OBSERVE = 2, MODE_RESET = 3 };

struct DevGroup {
    int32_t has_genset, has_grid, horizon, n_act, obs_dim, n_envs;
    int32_t act_col_genset, act_col_battery, act_col_grid;
    int32_t tile_begin;                 // first CTA of this group in the fused launch
    int32_t n_seg;
    int32_t seg_start[5], seg_kind[5];  // observation row layout: segment starts (ascending) and kinds
    int32_t tma_ok;                     // the grid window can be staged with cp.async.bulk (16-byte aligned image slot)
    int32_t state_start, state_genset_first;   // the battery + genset run: first element and order
    int32_t long_path;                         // rows too long for the staged path, or a state run at an odd element
    int32_t img_ok;                            // rows can be assembled by the image emitter (even grid offset, 1 + H <= 32)
    int32_t act_f32;                           // `actions` points at float32 values (MG_OPT_ACTIONS_F32)
    int32_t *step;
    double *charge;
    uint32_t *genset;
    const int32_t *cfg_index, *env_initial, *env_final;
    const uint32_t *status_bits;        // optional per-env grid status bitmask [n][status_words]
    int64_t status_words;
    // step io
    const double *actions;
    const int32_t *dactions;
    double *obs, *reward, *info;
    uint8_t *done;
    uint32_t *flags;
    const uint8_t *mask;
    double *reward_total;               // optional aggregate of the step's rewards over the group (logging)
    const double *soc_reported;         // mg_observe / mg_reset before the first step: the soc each battery was constructed with
    // rollout io
    double *reward_sum;
    int64_t act_step_stride, out_step_stride, obs_slot_stride, dact_step_stride;
    const int32_t *log_slot;            // rollout log of selected envs (MgRolloutIO.log_slot / log)
    double *log;
};

struct LaunchParams {
    int32_t n_groups, total_tiles, T, Tp, mode, normalized, n_steps, ring;
    int32_t pdl, _pad_pdl;      // single-step launches: chain consecutive launches with programmatic dependent launch (MG_OPT_STEP_OVERLAP)
    const MgConfig *cfg;
    const double *load_raw, *pv_raw, *grid_raw;
    const double *load_nrm, *pv_nrm, *grid_nrm;
    const MgPriorityList *plist;
    DevGroup g[MG_MAX_GROUPS];
};

// ------------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_global_v2(double *p, double a, double b) {
    // 16-byte store of row elements: written once, never re-read by this kernel, and not allowed to claim L1 lines the series
    // gathers live on (same-box ABAB against st.global.cs, profiles/r02_store_hint_abab.txt: 11.48 vs 11.66 us/step on the
    // lock-step headline, no difference elsewhere; -DMG_STORE_CS builds the evict-first form)
#ifdef MG_STORE_CS
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
#else
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
#endif
}
// float32 observation rows (MG_LAYOUT_OBS_F32): the f64 value rounded to nearest, 8-byte stores
__device__ __forceinline__ void st_global_v2(float *p, double a, double b) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(__double2float_rn(a)), "f"(__double2float_rn(b)) : "memory");
}

// Role timers (debug builds, -DMG_ROLE_TIMERS): where the warps of the persistent image kernels spend their cycles.
//   [0] owners: inputs + physics + publish   [1] owners: waiting for EMPTY   [2] emitters: waiting for FULL
//   [3] emitters: waiting for an image buffer (bulk_wait_read)   [4] emitters: gather + scatter + fence + issue   [5] warp-steps counted
#ifdef MG_ROLE_TIMERS
__device__ unsigned long long mg_role_cycles[8];
#define MG_T0() const long long t0_ = clock64()
#define MG_TACC(var) do { const long long t1_ = clock64(); var += t1_ - t0_; } while (0)
extern "C" int mg_debug_role_cycles(unsigned long long *out, int reset) {
    cudaMemcpyFromSymbol(out, mg_role_cycles, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(mg_role_cycles, z, sizeof z); }
    return 0;
}
#endif

// ---- TMA (cp.async.bulk, SASS UBLKCP) and mbarrier wrappers -------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MG_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MG_DONE;\n\t"
        "bra MG_WAIT;\n\t"
        "MG_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes and both addresses multiples of 16)
__device__ __forceinline__ void tma_load(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global bulk copy (SASS UBLKCP.G.S) in the issuing thread's bulk async-group; both addresses and the size are
// multiples of 16.  The source must stay untouched until bulk_wait_read says the copy has read it.
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N of this thread's bulk groups may still be reading their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// all of this thread's bulk groups have completed (their global writes included)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// order this thread's shared-memory writes (generic proxy) before later bulk copies (async proxy) that read them
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Aggregate reward of the tile's envs for logging: butterfly reduction with warp shuffles, one atomicAdd per warp.
// Called by whole warps (all 32 lanes); lanes without an env (or with a rejected step, reward NaN) contribute 0.
__device__ __forceinline__ void add_reward_total(double *total, double reward, bool has) {
    double v = (has && reward == reward) ? reward : 0.0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(total, v);
}

// programmatic dependent launch (PDL): wait for the preceding grid's trigger / let the following grid start
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool np_isclose(double a, double b, double rtol, double atol) {
    return fabs(a - b) <= (atol + rtol * fabs(b));
}

// genset_module.py:216-227 _reset_up_down_times
__device__ __forceinline__ void genset_reset_times(int cs, int U, int D, int &up, int &dn) {
    if (cs) { up = 0; dn = D; }
    else { dn = 0; up = U; }
}

// genset_module.py:360-390 next_status
__device__ __forceinline__ int genset_next_status(int cs, int up, int dn, int goal) {
    if (goal) return (cs || up == 0) ? 1 : 0;
    return (!cs || dn == 0) ? 0 : 1;
}

// genset_module.py:235-346 update_status (+ _finish_in_progress_change, _non_instantaneous_update)
__device__ __forceinline__ void genset_update_status(int &cs, int &gs, int &up, int &dn, double goal_status, int U,
                                                     int D, int abort) {
    const int goal = (int)rint(goal_status);  // Python round(): half to even
    if (goal == cs && cs == gs) return;
    const bool instant_up = (U == 0 && goal == 1);
    const bool instant_down = (D == 0 && goal == 0);
    if (goal != gs && (abort || instant_up || instant_down)) gs = goal;
    if (up == 0 && gs == 1) { cs = 1; genset_reset_times(cs, U, D, up, dn); return; }
    if (dn == 0 && gs == 0) { cs = 0; genset_reset_times(cs, U, D, up, dn); return; }
    if (goal == cs && cs != gs && abort) {
        gs = goal;
        genset_reset_times(cs, U, D, up, dn);
    } else if (cs == gs && gs != goal) {
        genset_reset_times(cs, U, D, up, dn);
        gs = goal;
    }
    if (gs != cs) {
        if (gs == 0) dn -= 1;
        else up -= 1;
    }
}

struct EnvRegs {      // per-env state carried in registers (across steps in the rollout kernel)
    int32_t t;
    int32_t cs, gs, up, dn;
    double charge;
};

struct RawRow {       // ts[t] of the env's series (reference: *_module.update reading self._time_series[t])
    double load, pv, imp, exp_, co2, status;
};

// PriorityListAlgo._populate_action, algos/priority_list/priority_list.py:69-167.
// Writes the unnormalised control: gen_goal, gen_energy, bat, grid.  Returns MG_FLAG_NEGATIVE_ABSORB where the reference's
// `assert module_max_consumption >= 0` (:124) fires: a battery whose charge sits an ulp above max_capacity asked to absorb.
__device__ __forceinline__ uint32_t priority_control(const MgPriorityList pl, const MgConfig *__restrict__ c, const DevGroup &G,
                                                 const EnvRegs &s, const RawRow &raw, double &gen_goal,
                                                 double &gen_energy, double &bat, double &grid) {
    gen_goal = 0.0; gen_energy = 0.0; bat = 0.0; grid = 0.0;
    double total_load = 0.0;
    total_load += -1 * raw.load;
    const double renewable = 0.0 + raw.pv;
    double remaining = total_load - renewable;
    bool genset_seen = false;
    uint32_t flags = 0;
#pragma unroll
    for (int i = 0; i < MG_PLIST_WIDTH; ++i) {
        const int mod = pl.module[i];
        const int act = pl.action[i];
        if (i >= pl.n_elements || mod < 0) continue;
        if (mod == MG_MOD_GENSET) {
            if (genset_seen) continue;
            genset_seen = true;
            gen_goal = (double)act;
        }
        double energy;
        if (np_isclose(remaining, 0.0, 1e-5, 1e-4)) {
            energy = 0.0;
        } else if (remaining > 0) {
            double mx, mn;
            if (mod == MG_MOD_GENSET) {
                const int ns = genset_next_status(s.cs, s.up, s.dn, act);
                mx = ns * c->gen_running_max;
                mn = ns * c->gen_running_min;
            } else if (mod == MG_MOD_BATTERY) {
                mx = fmin(c->bat_max_discharge, s.charge - c->bat_min_capacity) * c->bat_efficiency;
                mn = 0.0;
            } else {
                mx = c->grid_max_import * raw.status;
                mn = 0.0;
            }
            if (mn <= remaining && remaining <= mx) energy = remaining;
            else if (remaining < mn) energy = mn;
            else energy = mx;
        } else {
            if (mod == MG_MOD_GENSET) energy = 0.0;
            else {
                const double mc = (mod == MG_MOD_BATTERY)
                                      ? fmin(c->bat_max_charge, c->bat_max_capacity - s.charge) / c->bat_efficiency
                                      : c->grid_max_export * raw.status;
                if (!(mc >= 0)) flags |= MG_FLAG_NEGATIVE_ABSORB;
                if (-1 * remaining > mc) energy = -1.0 * mc;
                else energy = remaining;
            }
        }
        if (mod == MG_MOD_GENSET) gen_energy = energy;
        else if (mod == MG_MOD_BATTERY) bat = energy;
        else grid = energy;
        remaining -= energy;
    }
    return flags;
}

// One Microgrid.run for one env (microgrid.py:227-325; modules as cited inline).  Dispatch order
// load -> genset -> battery -> grid -> (balance) -> pv -> unbalanced_energy; reward and energy sums accumulate in
// exactly that order from 0.0 like MicrogridStep (microgrid/utils/step.py:9-36).
__device__ __forceinline__ bool shaper_in_range(double v) {
    return (-1 <= v && v <= 1) || np_isclose(v, 1.0, 1e-5, 1e-8) || np_isclose(v, 0.0, 1e-5, 1e-8);
}

__device__ __forceinline__ void env_step(const MgConfig *__restrict__ c, const DevGroup &G, EnvRegs &s, const RawRow &raw,
                                         double a_goal, double a_gen, double a_bat, double a_grid, bool normalized,
                                         int final_step, double &reward_out, int &done_out, uint32_t &flags_out,
                                         double *__restrict__ info) {
    uint32_t flags = 0;
    double reward = 0.0, provided = 0.0, consumed = 0.0;
    const int done = (s.t >= final_step - 1);   // base_timeseries_module.py:124-125, before t += 1
    double i_gen = 0.0, i_gen_co2 = 0.0, i_dis = 0.0, i_chg = 0.0, i_imp = 0.0, i_exp = 0.0, i_gco2 = 0.0;
    double r_gen = 0.0, r_bat = 0.0, r_grid = 0.0, r_unb = 0.0;   // per-module rewards (info block only)

    // LoadModule.update, load_module.py:86-91
    const double load = -1 * raw.load;
    consumed += load;
    reward += 0.0;

    if (G.has_genset) {   // GensetModule.step, genset_module.py:100-149, :183-214
        if (!(0 <= a_goal && a_goal <= 1)) flags |= MG_FLAG_GENSET_GOAL_RANGE;
        else genset_update_status(s.cs, s.gs, s.up, s.dn, a_goal, c->gen_start_up_time, c->gen_wind_down_time,
                                  c->gen_allow_abortion);
        const double a = normalized ? (0.0 + c->gen_act_spread * a_gen) : a_gen;
        double p;
        if (a < 0) {
            flags |= MG_FLAG_GENSET_AS_SINK;
            p = 0.0;
        } else {
            const double mx = s.cs * c->gen_running_max;
            const double mn = s.cs * c->gen_running_min;
            if (a > mx) { p = mx; flags |= MG_FLAG_CLIP_GENSET; }      // base_module.py:213-224, upper test first
            else if (a < mn) { p = mn; flags |= MG_FLAG_CLIP_GENSET; }
            else p = a;
        }
        const double co2 = c->gen_co2_per_unit * p;
        const double cost = c->gen_cost * p + c->gen_cost_per_unit_co2 * co2;
        r_gen = -1.0 * cost;
        reward += r_gen;
        provided += p;
        i_gen = p;
        i_gen_co2 = co2;
    }
    {   // BatteryModule, battery_module.py:108-130, :244-291
        const double a = normalized ? (c->bat_act_low + c->bat_act_spread * a_bat) : a_bat;
        double internal;
        if (a > 0 || a == 0) {
            const double mp = fmin(c->bat_max_discharge, s.charge - c->bat_min_capacity) * c->bat_efficiency;
            double p;
            if (a > mp) { p = mp; flags |= MG_FLAG_CLIP_BATTERY; }
            else p = a;
            internal = (-1.0 * p) / c->bat_efficiency;
            provided += p;
            i_dis = p;
        } else {
            flags |= MG_FLAG_BATTERY_SINK;
            double e = -1.0 * a;
            const double mc = fmin(c->bat_max_charge, c->bat_max_capacity - s.charge) / c->bat_efficiency;
            if (e > mc) { e = mc; flags |= MG_FLAG_CLIP_BATTERY; }
            if (!(e >= 0)) flags |= MG_FLAG_NEGATIVE_ABSORB;
            internal = e * c->bat_efficiency;
            consumed += e;
            i_chg = e;
        }
        s.charge += internal;
        if (s.charge < c->bat_min_capacity) {
            if (!np_isclose(s.charge, c->bat_min_capacity, 1e-5, 1e-8)) flags |= MG_FLAG_BATTERY_MIN_CAP;
            s.charge = c->bat_min_capacity;
        }
        r_bat = -1.0 * (fabs(internal) * c->bat_cost_cycle);
        reward += r_bat;
    }
    if (G.has_grid) {   // GridModule, grid_module.py:134-228, :314-320
        const double a = normalized ? (c->grid_act_low + c->grid_act_spread * a_grid) : a_grid;
        if (a > 0 || a == 0) {
            const double mp = c->grid_max_import * raw.status;
            double p;
            if (a > mp) { p = mp; flags |= MG_FLAG_CLIP_GRID; }
            else p = a;
            const double co2 = p * raw.co2;
            r_grid = -1 * raw.imp * p + (-1.0 * c->grid_cost_per_unit_co2 * co2);
            reward += r_grid;
            provided += p;
            i_imp = p;
            i_gco2 = co2;
        } else {
            flags |= MG_FLAG_GRID_SINK;
            double e = -1.0 * a;
            const double mc = c->grid_max_export * raw.status;
            if (e > mc) { e = mc; flags |= MG_FLAG_CLIP_GRID; }
            r_grid = raw.exp_ * e + (-1.0 * c->grid_cost_per_unit_co2 * 0.0);
            reward += r_grid;
            consumed += e;
            i_exp = e;
        }
    }
    // flex modules: pv then unbalanced_energy, microgrid.py:277-314
    const double difference = provided - consumed;
    double pv_used, loss = 0.0, overgen = 0.0;
    if (difference > 0) {
        flags |= MG_FLAG_EXCESS;
        pv_used = 0.0;
        provided += pv_used;
        reward += 0.0;
        overgen = difference;
        consumed += overgen;
        r_unb = -1.0 * (c->overgeneration_cost * overgen);
        reward += r_unb;
    } else {
        double needed = -difference;
        pv_used = (raw.pv < needed) ? raw.pv : needed;
        provided += pv_used;
        reward += 0.0;
        needed -= pv_used;
        loss = needed;
        provided += needed;
        r_unb = -1.0 * (c->loss_load_cost * needed);
        reward += r_unb;
    }
    if (!np_isclose(provided, consumed, 1e-5, 1e-8)) flags |= MG_FLAG_BALANCE;
    s.t += 1;
    // MicrogridStep.shaped_reward, microgrid/utils/step.py:41-46: a built-in shaper replaces the reward
    const int shaper = c->reward_shaper;
    if (shaper == MG_SHAPER_PV_CURTAILMENT) reward = -1.0 * (raw.pv - pv_used);
    else if (shaper == MG_SHAPER_BATTERY_DISCHARGE) {
        // evaluated twice by the reference: in the mid-step balance() before the flex modules (microgrid.py:277, no loss
        // load yet) and at the end; either assert raises (battery_discharge_shaper.py:33)
        const double mid = (i_dis - 0.0) / load;
        reward = (i_dis - loss) / load;
        if (!shaper_in_range(mid) || !shaper_in_range(reward)) flags |= MG_FLAG_SHAPER_RANGE;
    }
    reward_out = reward;
    done_out = done;
    flags_out = flags;
    if (info) {
        info[MG_INFO_LOAD_MET] = load;
        info[MG_INFO_PV_USED] = pv_used;
        info[MG_INFO_CURTAILMENT] = raw.pv - pv_used;
        info[MG_INFO_LOSS_LOAD] = loss;
        info[MG_INFO_OVERGENERATION] = overgen;
        info[MG_INFO_GENSET_PRODUCTION] = i_gen;
        info[MG_INFO_GENSET_CO2] = i_gen_co2;
        info[MG_INFO_BATTERY_DISCHARGE] = i_dis;
        info[MG_INFO_BATTERY_CHARGE] = i_chg;
        info[MG_INFO_GRID_IMPORT] = i_imp;
        info[MG_INFO_GRID_EXPORT] = i_exp;
        info[MG_INFO_GRID_CO2] = i_gco2;
        info[MG_INFO_REWARD_GENSET] = r_gen;
        info[MG_INFO_REWARD_BATTERY] = r_bat;
        info[MG_INFO_REWARD_GRID] = r_grid;
        info[MG_INFO_REWARD_UNBALANCED] = r_unb;
    }
}

// shared-memory record of one env of the tile, published by the env's owner thread after the physics: where its
// observation windows start in the *_nrm tables, and its battery / genset observation.
struct __align__(16) TileEnv {
    int32_t off_grid, off_load, off_pv;
    int32_t special;   // -1: plain table-backed row; >= 0: row needs per-env work (scaled series / status bits), value = t_obs
    double state[6];   // battery (soc, charge) and genset (cs, gs, up, dn) observation in ROW order
};

// per-env series parameters of a MicrogridGenerator-style grid, published next to TileEnv (kHetero kernels only) so
// that the emission lanes need no dependent global loads to normalise the env's own load / pv window
struct __align__(16) HeteroEnv {
    double load_scale, load_low, load_spread, load_fill, pv_scale, pv_low, pv_spread, pv_fill;
    int32_t load_base, pv_base, scaled, weak;   // element offsets of the profiles in the raw tables; flags
};

struct TileShared {
    alignas(128) double img[MG_WARPS][MG_MAX_IMG];   // per-warp TMA staging of a run's grid window (4 * (1 + H) <= MG_MAX_IMG f64)
    TileEnv env[2][MG_TILE];                         // double buffered across the steps of the persistent kernel
    alignas(8) uint64_t bar[MG_WARPS];               // per-warp completion of the TMA window load
};

struct TileSharedHetero {
    HeteroEnv het[2][MG_TILE];
};
struct Empty {
    int unused;
};
template <bool kHetero>
struct HeteroStorage;
template <>
struct HeteroStorage<true> {
    typedef TileSharedHetero type;
    __device__ static __forceinline__ HeteroEnv *rows(type &s, int buf) { return s.het[buf]; }
};
template <>
struct HeteroStorage<false> {
    typedef Empty type;
    __device__ static __forceinline__ HeteroEnv *rows(type &, int) { return nullptr; }
};

template <bool kHetero>
__device__ __forceinline__ void publish_env(TileEnv &te, HeteroEnv *het, const MgConfig *__restrict__ c, const DevGroup &G,
                                            const EnvRegs &s, int T, int Tp) {
    // rows >= T of the tables hold the forecaster's fill value, so a window that runs past the end of the series
    // needs no branch (forecast/forecaster.py:120-137)
    const int t_obs = min(s.t, T);
    te.off_load = c->load_series * Tp + t_obs;
    te.off_pv = c->pv_series * Tp + t_obs;
    te.off_grid = G.has_grid ? (c->grid_series * Tp + t_obs) * 4 : 0;
    te.special = (kHetero && (c->series_scaled || (G.has_grid && G.status_bits))) ? t_obs : -1;
    if (kHetero && het) {
        het->load_scale = c->load_scale; het->load_low = c->load_low; het->load_spread = c->load_spread; het->load_fill = c->load_fill_nrm;
        het->pv_scale = c->pv_scale; het->pv_low = c->pv_low; het->pv_spread = c->pv_spread; het->pv_fill = c->pv_fill_nrm;
        het->load_base = c->load_series * T; het->pv_base = c->pv_series * T;
        het->scaled = c->series_scaled; het->weak = c->grid_status_weak;
    }
    // battery_module.py:323-330, genset_module.py:503-509, utils/space.py:207-218
    const double soc = s.charge / c->bat_max_capacity;
    const double b0 = (soc - c->bat_soc_low) / c->bat_soc_spread;
    const double b1 = (s.charge - c->bat_min_capacity) / c->bat_charge_spread;
    if (!G.has_genset) {
        te.state[0] = b0; te.state[1] = b1;
        return;
    }
    const double g0 = ((double)s.cs - 0.0) / 1.0;
    const double g1 = ((double)s.gs - 0.0) / 1.0;
    const double g2 = ((double)s.up - 0.0) / c->gen_up_spread;
    const double g3 = ((double)s.dn - 0.0) / c->gen_down_spread;
    if (G.state_genset_first) {   // container order: genset, battery
        te.state[0] = g0; te.state[1] = g1; te.state[2] = g2; te.state[3] = g3; te.state[4] = b0; te.state[5] = b1;
    } else {                      // gym order: battery, genset
        te.state[0] = b0; te.state[1] = b1; te.state[2] = g0; te.state[3] = g1; te.state[4] = g2; te.state[5] = g3;
    }
}

// element j of an observation row -> (kind, offset inside the module's obs vector)
__device__ __forceinline__ void decode_element(const DevGroup &G, int j, int &kind, int &off) {
    int sgi = 0;
#pragma unroll
    for (int q = 1; q < 5; ++q)
        if (q < G.n_seg && j >= G.seg_start[q]) sgi = q;
    kind = G.seg_kind[sgi];
    off = j - G.seg_start[sgi];
}

// Where each module's block starts inside an observation row (the row layout is static per group).
struct RowStarts {
    int load, pv, grid, state;
};
__device__ __forceinline__ RowStarts row_starts(const DevGroup &G) {
    RowStarts st;
    st.load = st.pv = st.grid = 0;
    st.state = G.state_start;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        if (q < G.n_seg) {
            if (G.seg_kind[q] == KIND_LOAD) st.load = G.seg_start[q];
            else if (G.seg_kind[q] == KIND_PV) st.pv = G.seg_start[q];
            else if (G.seg_kind[q] == KIND_GRID) st.grid = G.seg_start[q];
        }
    }
    return st;
}

// ---- sliding windows of the persistent kernel for per-env series (kRing) --------------------------------------------
// A MicrogridGenerator-style env has its own load / pv series (profile * scale), so no two rows share a window and the
// step kernel normalises all 2 (1 + H) window values of a row on the fly (emit_row_hetero: two f64 divisions per lane
// and row).  Inside a rollout the window of step s+1 is the window of step s shifted by one: the persistent kernel keeps
// every env's normalised windows in shared memory -- ring of R = H + 2 slots per env and series, absolute series index
// idx lives in slot idx mod R -- and the env's owner thread appends ONE new value per series and step.  R = H + 2, not
// H + 1: the slot the owner overwrites for step s+1 then belongs to index t_s - 1, which the warps still streaming the
// rows of step s ([t_s, t_s + H]) do not read.
struct RingShared {
    double win[2][MG_TILE][MG_RING_MAX];   // [0] load, [1] pv
};
template <bool kRing>
struct RingStorage;
template <>
struct RingStorage<true> {
    typedef RingShared type;
    __device__ static __forceinline__ RingShared *get(type &s) { return &s; }
};
template <>
struct RingStorage<false> {
    typedef Empty type;
    __device__ static __forceinline__ RingShared *get(type &) { return nullptr; }
};

__device__ __forceinline__ bool env_is_special(const MgConfig *__restrict__ c, const DevGroup &G) {
    return c->series_scaled || (G.has_grid && G.status_bits);
}

// normalised load / pv observation value at absolute series index idx (idx <= T + H): the env's own series where it is
// profile * scale -- (profile * scale - low) / spread, or the normalised forecaster fill past the end, exactly as
// emit_row_hetero computes it -- else the pre-normalised, end-padded table
__device__ __forceinline__ void series_obs_value(const LaunchParams &P, const MgConfig *__restrict__ c, int idx, double &ld, double &pv) {
    if (c->series_scaled) {
        const bool in = idx < P.T;
        const double rp = in ? __ldg(P.pv_raw + (size_t)c->pv_series * P.T + idx) : 0.0;
        const double rl = in ? __ldg(P.load_raw + (size_t)c->load_series * P.T + idx) : 0.0;
        const double np_ = (rp * c->pv_scale - c->pv_low) / c->pv_spread;
        const double nl = (rl * c->load_scale - c->load_low) / c->load_spread;
        pv = in ? np_ : c->pv_fill_nrm;
        ld = in ? nl : c->load_fill_nrm;
    } else {
        pv = __ldg(P.pv_nrm + (size_t)c->pv_series * P.Tp + idx);
        ld = __ldg(P.load_nrm + (size_t)c->load_series * P.Tp + idx);
    }
}

// series_obs_value in two halves, so that a caller can issue the loads long before it needs the values
struct SeriesRaw {
    double a, b;      // scaled series: raw pv / load profile values; table-backed: the normalised pv / load values
    bool in;
};
__device__ __forceinline__ SeriesRaw series_obs_fetch(const LaunchParams &P, const MgConfig *__restrict__ c, int idx) {
    SeriesRaw r;
    if (c->series_scaled) {
        r.in = idx < P.T;
        r.a = r.in ? __ldg(P.pv_raw + (size_t)c->pv_series * P.T + idx) : 0.0;
        r.b = r.in ? __ldg(P.load_raw + (size_t)c->load_series * P.T + idx) : 0.0;
    } else {
        r.in = true;
        r.a = __ldg(P.pv_nrm + (size_t)c->pv_series * P.Tp + idx);
        r.b = __ldg(P.load_nrm + (size_t)c->load_series * P.Tp + idx);
    }
    return r;
}
__device__ __forceinline__ void series_obs_finish(const MgConfig *__restrict__ c, const SeriesRaw &r, double &ld, double &pv) {
    if (c->series_scaled) {
        const double np_ = (r.a * c->pv_scale - c->pv_low) / c->pv_spread;
        const double nl = (r.b * c->load_scale - c->load_low) / c->load_spread;
        pv = r.in ? np_ : c->pv_fill_nrm;
        ld = r.in ? nl : c->load_fill_nrm;
    } else {
        pv = r.a;
        ld = r.b;
    }
}
// status_window in two halves
__device__ __forceinline__ uint2 status_words_fetch(const DevGroup &G, int e, int t_obs) {
    const uint32_t *bits = G.status_bits + (size_t)e * G.status_words;
    const int i = t_obs >> 5;
    uint2 w;
    w.x = __ldg(bits + i);
    w.y = (i + 1 < G.status_words) ? __ldg(bits + i + 1) : 0u;
    return w;
}

// the 32 grid-status bits of env e starting at step t_obs (bit k = status at t_obs + k)
__device__ __forceinline__ uint32_t status_window(const DevGroup &G, int e, int t_obs) {
    const uint32_t *bits = G.status_bits + (size_t)e * G.status_words;
    const int i = t_obs >> 5;
    const uint32_t lo = __ldg(bits + i);
    const uint32_t hi = (i + 1 < G.status_words) ? __ldg(bits + i + 1) : 0u;
    return __funnelshift_r(lo, hi, t_obs & 31);
}

// TileEnv of a ring row: off_grid = grid window offset (columns 0-2), special = t_obs, and -- the table offsets of the
// load / pv windows being meaningless for it -- off_load = the status window word, off_pv = ring base | weak << 16
__device__ __forceinline__ void publish_ring(TileEnv &te, const MgConfig *__restrict__ c, const DevGroup &G, int e, int base) {
    const int t_obs = te.special;
    te.off_load = (G.has_grid && G.status_bits) ? (int32_t)status_window(G, e, t_obs) : 0;
    te.off_pv = base | (c->grid_status_weak ? (1 << 16) : 0);
}

// One observation row of a ring env.  Warp-cooperative and block-wise like emit_row_hetero -- every lane of a block runs
// the same code, so nothing diverges -- but with nothing left to compute: lane k copies window element k of the pv and the
// load ring, lanes copy grid elements l, l + 32, ... from the table (a lane always sees the same grid column because 32
// is a multiple of 4; the status column comes from the env's window word), lanes 0..5 the battery / genset values; the
// row is assembled in the warp's shared-memory image and streamed out with 16-byte stores.
// (A first version gathered each lane's six row elements straight into registers: 260 warp-instructions per row, almost
//  all of them integer / branch work for the per-element kind dispatch -- profiles/r01_ncu_generator_ring_v1.txt.)
// (row-invariant quantities are hoisted into RingRowCtx once per step; every loop has a compile-time trip count -- the ring
//  path implies 1 + H <= 25 window elements, 4 (1 + H) <= 100 grid elements and at most MG_MAX_IMG / 2 pairs per row)
struct RingRowCtx {
    int pairs, rows, R, T, pv0, load0, grid0, state0, n_state;
    bool has_grid, own_status;
    const double *grid_nrm;
};
__device__ __forceinline__ RingRowCtx ring_row_ctx(const LaunchParams &P, const DevGroup &G, const RowStarts &rs) {
    RingRowCtx x;
    x.pairs = G.obs_dim >> 1; x.rows = 1 + G.horizon; x.R = G.horizon + 2; x.T = P.T;
    x.pv0 = rs.pv; x.load0 = rs.load; x.grid0 = rs.grid; x.state0 = rs.state; x.n_state = 2 + 4 * G.has_genset;
    x.has_grid = G.has_grid != 0; x.own_status = G.status_bits != nullptr;
    x.grid_nrm = P.grid_nrm;
    return x;
}
template <typename TO>
__device__ __forceinline__ void emit_row_ring(const RingRowCtx &x, const TileEnv &te, const double *__restrict__ ring_load,
                                              const double *__restrict__ ring_pv, double *__restrict__ img, TO *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int t_obs = te.special, base = te.off_pv & 0xffff;
    const bool weak = (te.off_pv >> 16) != 0;
    const uint32_t w = (uint32_t)te.off_load;
    __syncwarp();   // the previous row has been read out of the image
    if (lane < x.rows) {
        int slot = base + lane;
        if (slot >= x.R) slot -= x.R;
        img[x.pv0 + lane] = ring_pv[slot];
        img[x.load0 + lane] = ring_load[slot];
    }
    if (x.has_grid) {
        const bool status_lane = x.own_status && (lane & 3) == 3;
        const double *__restrict__ gsrc = x.grid_nrm + te.off_grid;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g = lane + 32 * j;
            if (g < 4 * x.rows) {
                const int kk = g >> 2;
                // bounds of the status column: (0, 1) on a weak grid, (1, 1) -> spread 1 otherwise (utils/space.py:204-205)
                const double own = weak ? (t_obs + kk < x.T ? (((w >> kk) & 1u) ? 1.0 : 0.0) : 0.5) : 0.0;
                const double tab = __ldg(gsrc + g);
                img[x.grid0 + g] = status_lane ? own : tab;
            }
        }
    }
    if (lane < x.n_state) img[x.state0 + lane] = te.state[lane];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < MG_MAX_IMG / 64; ++j) {
        const int p = lane + 32 * j;
        if (p < x.pairs) {
            const double2 v2 = *reinterpret_cast<const double2 *>(img + 2 * p);
            st_global_v2(out + 2 * p, v2.x, v2.y);
        }
    }
}

// One observation row of an env with per-env series (MicrogridGenerator grids: load / pv = profile * scale normalised
// on the fly, grid status from the env's own bit row).  Warp-cooperative, one block of the row at a time so that every
// lane of a block runs the same code: lane l normalises window element l (+32) of the pv block, then of the load block
// (one f64 division site each), copies grid elements l, l+32, ... (a lane always sees the same grid column because 32 is
// a multiple of 4: the lanes with (l & 3) == 3 read the env's status bits instead of the table), lanes 0..5 copy the
// battery / genset pairs; the row is assembled in the warp's shared-memory image and streamed out with 16-byte stores.
template <typename TO>
__device__ __forceinline__ void emit_row_hetero(const LaunchParams &P, const DevGroup &G, const TileEnv &te, const HeteroEnv &hv,
                                                const RowStarts &rs, double *__restrict__ img, int e, TO *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int D = G.obs_dim, rows = 1 + G.horizon, t_obs = te.special;
    __syncwarp();   // the previous row has been read out of the image
    for (int k = lane; k < rows; k += 32) {   // pv and load windows
        const int idx = t_obs + k;
        double pv, ld;
        if (hv.scaled) {   // (profile * scale - low) / spread, or the normalised forecaster fill past the end
            const bool in = idx < P.T;
            const double rp = in ? __ldg(P.pv_raw + hv.pv_base + idx) : 0.0;
            const double rl = in ? __ldg(P.load_raw + hv.load_base + idx) : 0.0;
            const double np_ = (rp * hv.pv_scale - hv.pv_low) / hv.pv_spread;
            const double nl = (rl * hv.load_scale - hv.load_low) / hv.load_spread;
            pv = in ? np_ : hv.pv_fill;
            ld = in ? nl : hv.load_fill;
        } else {
            pv = __ldg(P.pv_nrm + te.off_pv + k);
            ld = __ldg(P.load_nrm + te.off_load + k);
        }
        img[rs.pv + k] = pv;
        img[rs.load + k] = ld;
    }
    if (G.has_grid) {
        const bool status_lane = G.status_bits != nullptr && (lane & 3) == 3;
        const uint32_t *bits = G.status_bits + (size_t)e * G.status_words;
        for (int g = lane; g < 4 * rows; g += 32) {
            double v;
            if (status_lane) {
                const int idx = t_obs + (g >> 2);
                const double bit = (double)((__ldg(bits + (idx >> 5)) >> (idx & 31)) & 1u);
                // bounds of the status column: (0, 1) on a weak grid, (1, 1) -> spread 1 otherwise (utils/space.py:204-205)
                v = hv.weak ? (idx < P.T ? bit : 0.5) : 0.0;
            } else {
                v = __ldg(P.grid_nrm + te.off_grid + g);
            }
            img[rs.grid + g] = v;
        }
    }
    if (lane < 2 + 4 * G.has_genset) img[rs.state + lane] = te.state[lane];
    __syncwarp();
    for (int p = lane; p < (D >> 1); p += 32) {
        const double2 v2 = *reinterpret_cast<const double2 *>(img + 2 * p);
        st_global_v2(out + 2 * p, v2.x, v2.y);
    }
}

// Time-series part of the observation rows of one tile.  Each warp owns MG_ROWS_PER_WARP consecutive rows, i.e. one
// contiguous chunk of the [n, obs_dim] output; lane l owns the 16-byte pairs l, l+32, ... of every row.
// Consecutive rows whose windows coincide (same series, same step -- every row of the tile in the lock-step case)
// form a run: the lane fetches its pairs of the [t, t+H] windows ONCE per run into registers -- for runs of
// MG_TMA_MIN_RUN rows or more the grid window (2/3 of the row) is first staged in shared memory by one TMA bulk load,
// shorter runs read the tables directly -- and the run is then written with nothing but coalesced 16-byte stores
// (3 per row per lane for a 150-element row).  Rows at unrelated steps degrade gracefully to one fetch per row.
// The 1-3 lanes that own the battery / genset pairs take them from the env's shared-memory record instead, so every
// byte of a row -- and of the contiguous 19 KB chunk of 16 rows -- is written by one warp in consecutive instructions
// (a separate writer for those 48 bytes costs ~20% of the store bandwidth: partial-sector merging in L2).
template <int SLOTS, bool kHetero, typename TO, bool kRing = false>
__device__ __forceinline__ void warp_emit_rows_t(const LaunchParams &P, const DevGroup &G, TileShared &S, const HeteroEnv *het,
                                                 int ebuf, TO *__restrict__ obs_tile, int n_rows, int e_base, uint32_t &phase,
                                                 int r_begin, const RingShared *rings = nullptr) {
    // rows [r_begin, r_begin + MG_ROWS_PER_WARP) of the tile; staging slot and mbarrier are the calling warp's own
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r_end = min(r_begin + MG_ROWS_PER_WARP, n_rows);
    const int D = G.obs_dim, pairs = D >> 1;
    const int sp0 = G.state_start >> 1, sp1 = sp0 + 1 + 2 * G.has_genset;
    const TileEnv *env = S.env[ebuf];
    double *img = S.img[warp];
    // static layout of this lane's pairs: kind << 16 | offset inside the module's window
    int code[SLOTS][2];
    bool act[SLOTS], st_lane[SLOTS];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
        const int p = lane + 32 * k;
        act[k] = p < pairs;
        st_lane[k] = p >= sp0 && p < sp1;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int kind, off;
            decode_element(G, 2 * p + h, kind, off);
            code[k][h] = (kind << 16) | off;
        }
    }
    const RowStarts rs = row_starts(G);   // used by rows with per-env series (kHetero kernels only)
    RingRowCtx rx;
    if (kRing) rx = ring_row_ctx(P, G, rs);
    const bool tma_grid = G.has_grid && G.tma_ok;
    const uint32_t grid_bytes = (uint32_t)(4 * (1 + G.horizon) * sizeof(double));
    // run boundaries of this warp's rows in one vote: bit l set <=> row l starts a new run
    uint32_t starts;
    {
        const int rr = r_begin + lane;
        bool brk = false;
        if (lane > 0 && lane < MG_ROWS_PER_WARP && rr < r_end) {
            const TileEnv a = env[rr], b = env[rr - 1];
            brk = a.off_grid != b.off_grid || a.off_load != b.off_load || a.off_pv != b.off_pv ||
                  (kHetero && (a.special >= 0 || b.special >= 0));
        }
        starts = __ballot_sync(0xffffffffu, brk) | (r_end > r_begin ? (1u << (r_end - r_begin)) : 0u);
    }
    int r = r_begin;
    while (r < r_end) {
        const TileEnv sig = env[r];
        const int n = __ffs(starts >> (r - r_begin + 1));   // distance to the next run start (or to the end marker)
        const bool stage = tma_grid && n >= MG_TMA_MIN_RUN;
        if (stage) {
            __syncwarp();   // every lane has finished reading the previous run's staged window
            if (lane == 0) {   // the grid window is one contiguous, 32-byte aligned run of the normalised table
                mbar_expect_tx(&S.bar[warp], grid_bytes);
                tma_load(img, P.grid_nrm + sig.off_grid, grid_bytes, &S.bar[warp]);
            }
        }
        if (kHetero && sig.special >= 0) {   // a row with per-env series: always a run of one
            if (kRing) emit_row_ring<TO>(rx, env[r], rings->win[0][r], rings->win[1][r], img, obs_tile + (size_t)r * D);
            else emit_row_hetero(P, G, env[r], het[r], rs, img, e_base + r, obs_tile + (size_t)r * D);
            r += 1;
            continue;
        }
        double v[SLOTS][2];
        {
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int kind = code[k][h] >> 16, off = code[k][h] & 0xffff;
                    v[k][h] = 0.0;
                    if (act[k] && !st_lane[k]) {
                        if (kind == KIND_LOAD) v[k][h] = __ldg(P.load_nrm + sig.off_load + off);
                        else if (kind == KIND_PV) v[k][h] = __ldg(P.pv_nrm + sig.off_pv + off);
                        else if (kind == KIND_GRID && !stage) v[k][h] = __ldg(P.grid_nrm + sig.off_grid + off);
                    }
                }
            }
        }
        if (stage) {
            mbar_wait(&S.bar[warp], phase);
            phase ^= 1;
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int kind = code[k][h] >> 16, off = code[k][h] & 0xffff;
                    if (act[k] && kind == KIND_GRID) v[k][h] = img[off];
                }
            }
        }
        TO *out = obs_tile + (size_t)r * D + 2 * lane;
        for (int row = 0; row < n; ++row, out += D) {
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) {
                if (st_lane[k]) {
                    const double2 sv = *reinterpret_cast<const double2 *>(env[r + row].state + 2 * (lane + 32 * k - sp0));
                    st_global_v2(out + 64 * k, sv.x, sv.y);
                } else if (act[k]) {
                    st_global_v2(out + 64 * k, v[k][0], v[k][1]);
                }
            }
        }
        r += n;
    }
}

template <bool kHetero, typename TO, bool kRing = false>
__device__ __forceinline__ void warp_emit_rows(const LaunchParams &P, const DevGroup &G, TileShared &S, const HeteroEnv *het,
                                               int ebuf, TO *__restrict__ obs_tile, int n_rows, int e_base, uint32_t &phase,
                                               int r_begin, const RingShared *rings = nullptr) {
    const int pairs = G.obs_dim >> 1;
    if (pairs <= 32) warp_emit_rows_t<1, kHetero, TO, kRing>(P, G, S, het, ebuf, obs_tile, n_rows, e_base, phase, r_begin, rings);
    else warp_emit_rows_t<3, kHetero, TO, kRing>(P, G, S, het, ebuf, obs_tile, n_rows, e_base, phase, r_begin, rings);
}

// rows longer than MG_MAX_IMG: element-wise path straight from the tables (no staging)
template <typename TO>
__device__ __forceinline__ void warp_emit_rows_long(const LaunchParams &P, const DevGroup &G, TileShared &S, int ebuf,
                                                    TO *__restrict__ obs_tile, int n_rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r_end = min((warp + 1) * MG_ROWS_PER_WARP, n_rows);
    const int D = G.obs_dim, pairs = D >> 1;
    for (int r = warp * MG_ROWS_PER_WARP; r < r_end; ++r) {
        const TileEnv te = S.env[ebuf][r];
        for (int p = lane; p < pairs; p += 32) {
            double v[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int kind, off;
                decode_element(G, 2 * p + h, kind, off);
                if (kind == KIND_LOAD) v[h] = __ldg(P.load_nrm + te.off_load + off);
                else if (kind == KIND_PV) v[h] = __ldg(P.pv_nrm + te.off_pv + off);
                else if (kind == KIND_GRID) v[h] = __ldg(P.grid_nrm + te.off_grid + off);
                else v[h] = S.env[ebuf][r].state[2 * p + h - G.state_start];
            }
            st_global_v2(obs_tile + (size_t)r * D + 2 * p, v[0], v[1]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Image emitter: rows that share no window with their neighbours (per-env series, envs at unrelated steps).
// A warp assembles RC consecutive rows in a shared-memory image -- grid window by 16-byte read-only loads from the
// normalised table (L1 / L2 resident), load / pv window from the env's shared-memory ring, the tables or normalised on the
// fly, battery / genset values from the tile record -- and hands the RC * obs_dim * 8 contiguous output bytes to the TMA
// unit as ONE bulk store (cp.async.bulk shared -> global).  All loads of a chunk are issued before its first
// shared-memory store, so their latencies overlap across rows; the bulk store drains in the background while the next
// chunk is gathered into the other image buffer (NB buffers per warp, recycled through the bulk-group queue).  No lane
// ever issues a global store: the LSU only sees the gathers.
// Needs obs_dim even, 1 + H <= 32, an even grid offset (16-byte aligned pairs in the image) and f64 rows.
// ------------------------------------------------------------------------------------------------------------------
// shared-memory accesses by 32-bit address (the emitter's address arithmetic stays in 32 bits, one add per access)
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_v2f64(uint32_t a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ int4 lds_v4s32(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

struct ImgCtx {     // per group and lane, hoisted out of every loop
    int T, ringR, row_bytes;
    int lk;                              // window element this lane copies (clamped to a valid one for lanes past the window)
    const double *pv_src, *load_src;     // table pointers already offset by this lane's window element
    const double2 *g0_src, *g1_src;      // grid table pointers already offset by this lane's two pairs (clamped)
    uint32_t a_pv, a_load, a_state, a_g0, a_g1;   // byte offsets of this lane's destinations inside a row image
    uint32_t s_state;                    // byte offset of this lane's state element inside a TileEnv record
    int k0, k1;                          // forecast row of this lane's two grid pairs (status column patch)
    bool win_lane, state_lane, g0, g1, patch_lane, own_status, has_grid;
};
__device__ __forceinline__ ImgCtx img_ctx(const LaunchParams &P, const DevGroup &G) {
    const RowStarts rs = row_starts(G);
    const int lane = threadIdx.x & 31, rows = 1 + G.horizon, n_state = 2 + 4 * G.has_genset;
    ImgCtx x;
    x.T = P.T; x.ringR = G.horizon + 2; x.row_bytes = G.obs_dim * (int)sizeof(double);
    x.has_grid = G.has_grid != 0;
    x.win_lane = lane < rows; x.state_lane = lane < n_state;
    x.lk = min(lane, rows - 1);
    x.g0 = x.has_grid && lane < 2 * rows; x.g1 = x.has_grid && lane + 32 < 2 * rows;
    x.pv_src = P.pv_nrm + x.lk; x.load_src = P.load_nrm + x.lk;
    x.g0_src = reinterpret_cast<const double2 *>(P.grid_nrm) + (x.g0 ? lane : 0);
    x.g1_src = reinterpret_cast<const double2 *>(P.grid_nrm) + (x.g1 ? lane + 32 : 0);
    x.a_pv = 8u * (rs.pv + lane); x.a_load = 8u * (rs.load + lane); x.a_state = 8u * (rs.state + lane);
    x.a_g0 = 8u * (rs.grid + 2 * lane); x.a_g1 = 8u * (rs.grid + 2 * (lane + 32));
    x.s_state = (uint32_t)offsetof(TileEnv, state) + 8u * min(lane, n_state - 1);
    x.k0 = lane >> 1; x.k1 = (lane + 32) >> 1;
    x.own_status = G.status_bits != nullptr;
    // odd grid pairs hold (co2, status); 32 is even, so both of a lane's pairs agree
    x.patch_lane = x.own_status && (lane & 1) != 0;
    return x;
}

struct ImgRow {     // what one lane holds of one row between the gather and the scatter
    double2 g0, g1;
    double pv, load, state;
};
struct ImgShared {  // the parts of the previous row that following rows of the same run reuse (table values, unpatched)
    double2 g0, g1;
    double pv, load;
};

// Gather this lane's share of row r (tile record at shared address env_a + 64 r).  Every load is unconditional (clamped
// addresses) and nothing diverges between lanes; need_grid / need_win are warp-uniform: false where the row's grid window /
// load + pv windows equal the previous row's (same table offsets), which `sh` still holds -- the whole tile in lock-step.
template <bool kHetero, bool kRing>
__device__ __forceinline__ ImgRow img_gather(const LaunchParams &P, const DevGroup &G, const ImgCtx &x, uint32_t env_a,
                                             const HeteroEnv *__restrict__ het, uint32_t ring_a, int r, int e_base, bool need_grid,
                                             bool need_win, ImgShared &sh) {
    static_assert(sizeof(TileEnv) == 64, "the emitter addresses tile records by hand");
    ImgRow v;
    const uint32_t rec = env_a + 64u * (uint32_t)r;
    int4 hd = make_int4(0, 0, 0, -1);
    if (kHetero || need_grid || need_win) hd = lds_v4s32(rec);      // off_grid, off_load, off_pv, special
    const int off_grid = hd.x, off_load = hd.y, off_pv = hd.z, special = kHetero ? hd.w : -1;
    if (x.has_grid) {   // group-uniform
        if (need_grid) {
            sh.g0 = __ldg(x.g0_src + (off_grid >> 1));
            sh.g1 = __ldg(x.g1_src + (off_grid >> 1));
        }
        v.g0 = sh.g0;
        v.g1 = sh.g1;
        if (kHetero && x.own_status && special >= 0) {   // warp-uniform
            // The env's own status column.  Bounds (0, 1) on a weak grid, (1, 1) -> spread 1 otherwise (utils/space.py:204-205):
            // a grid without outages reports 0.0, which is what the shared table holds, so only weak grids patch.
            uint32_t w;
            bool weak;
            if (kRing) { w = (uint32_t)off_load; weak = (off_pv >> 16) != 0; }
            else { weak = het[r].weak != 0; w = weak ? status_window(G, e_base + r, special) : 0u; }
            if (weak) {
                double s0 = (double)((w >> x.k0) & 1u), s1 = (double)((w >> (x.k1 & 31)) & 1u);
                if (special + 32 > x.T) {   // the window runs past the end of the series: the forecaster's fill
                    s0 = special + x.k0 < x.T ? s0 : 0.5;
                    s1 = special + x.k1 < x.T ? s1 : 0.5;
                }
                v.g0.y = x.patch_lane ? s0 : v.g0.y;
                v.g1.y = x.patch_lane ? s1 : v.g1.y;
            }
        }
    } else {
        v.g0 = v.g1 = make_double2(0.0, 0.0);
    }
    if (kRing && special >= 0) {   // warp-uniform
        int slot = (off_pv & 0xffff) + x.lk;
        slot -= slot >= x.ringR ? x.ringR : 0;
        const uint32_t a = ring_a + 8u * (uint32_t)(r * MG_RING_MAX + slot);
        v.load = lds_f64(a);                                     // RingShared::win[0][r][slot]
        v.pv = lds_f64(a + 8u * MG_TILE * MG_RING_MAX);          // RingShared::win[1][r][slot]
    } else if (kHetero && !kRing && special >= 0 && het[r].scaled) {
        // (profile * scale - low) / spread, or the normalised forecaster fill past the end (as emit_row_hetero)
        const HeteroEnv &hv = het[r];
        const int idx = special + x.lk;
        const bool in = idx < x.T;
        const double rp = in ? __ldg(P.pv_raw + hv.pv_base + idx) : 0.0;
        const double rl = in ? __ldg(P.load_raw + hv.load_base + idx) : 0.0;
        const double np_ = (rp * hv.pv_scale - hv.pv_low) / hv.pv_spread;
        const double nl = (rl * hv.load_scale - hv.load_low) / hv.load_spread;
        v.pv = in ? np_ : hv.pv_fill;
        v.load = in ? nl : hv.load_fill;
    } else {
        if (need_win) {
            sh.pv = __ldg(x.pv_src + off_pv);
            sh.load = __ldg(x.load_src + off_load);
        }
        v.pv = sh.pv;
        v.load = sh.load;
    }
    v.state = lds_f64(rec + x.s_state);
    return v;
}

// write this lane's share of a row into the image at shared address row_a: five predicated stores
__device__ __forceinline__ void img_scatter(const ImgCtx &x, uint32_t row_a, const ImgRow &v) {
    if (x.g0) sts_v2f64(row_a + x.a_g0, v.g0);
    if (x.g1) sts_v2f64(row_a + x.a_g1, v.g1);
    if (x.win_lane) sts_f64(row_a + x.a_pv, v.pv);
    if (x.win_lane) sts_f64(row_a + x.a_load, v.load);
    if (x.state_lane) sts_f64(row_a + x.a_state, v.state);
}

// env_a / ring_a / img_a: shared-memory addresses of the tile records of this step, of the RingShared block and of the
// calling warp's NB image buffers (warp-uniform values, so that the bulk store's operands stay in uniform registers).
// r_count <= 32 rows per call.
template <int RC, int NB, int GB, bool kHetero, bool kRing>
__device__ __forceinline__ void warp_emit_rows_img(const LaunchParams &P, const DevGroup &G, const ImgCtx &x, uint32_t env_a,
                                                   const HeteroEnv *__restrict__ het, uint32_t ring_a, uint32_t img_a, int &buf,
                                                   double *__restrict__ obs_tile, int n_rows, int r_begin, int r_count, int e_base) {
    const int lane = threadIdx.x & 31;
    const int r_end = min(r_begin + r_count, n_rows);
    const uint32_t buf_bytes = (uint32_t)(RC * x.row_bytes);
    // Per-env series batches: run boundaries of this call's rows in two votes -- bit i set <=> row r_begin + i has a grid
    // window / load + pv windows of its own (different table offsets than the row before it, or per-env windows); the other
    // rows reuse `sh` (envs are laid out by grid table, so most rows of a tile continue the previous row's grid window).
    // Table-backed batches come here when their envs are at unrelated steps (in lock-step the run emitters of
    // warp_emit_rows_t are the faster ones): every row gathers its own windows, with no run bookkeeping in the way.
    uint32_t new_grid = 0xffffffffu, new_win = 0xffffffffu;
    if (kHetero) {
        const int rr = min(r_begin + lane, max(r_end - 1, r_begin));
        const int4 me = lds_v4s32(env_a + 64u * (uint32_t)rr);
        const int4 pr = lds_v4s32(env_a + 64u * (uint32_t)max(rr - 1, r_begin));
        const bool first = lane == 0;
        new_grid = __ballot_sync(0xffffffffu, first || me.x != pr.x);
        new_win = __ballot_sync(0xffffffffu, first || me.y != pr.y || me.z != pr.z || me.w >= 0 || pr.w >= 0);
    }
    ImgShared sh;
    sh.g0 = sh.g1 = make_double2(0.0, 0.0);
    sh.pv = sh.load = 0.0;
#pragma unroll 1
    for (int r = r_begin; r < r_end; r += RC) {
        const uint32_t im = img_a + (uint32_t)buf * buf_bytes;
        const uint32_t ng = kHetero ? new_grid >> (r - r_begin) : 0xffffffffu, nw = kHetero ? new_win >> (r - r_begin) : 0xffffffffu;
#ifdef MG_ROLE_TIMERS
        const long long tw0 = clock64();
#endif
        if (lane == 0) bulk_wait_read<NB - 1>();    // the bulk store that last read this buffer is done with it
        __syncwarp();
#ifdef MG_ROLE_TIMERS
        if (lane == 0) atomicAdd(&mg_role_cycles[3], (unsigned long long)(clock64() - tw0));
#endif
        int n = RC;
        if (r + RC <= r_end) {                      // full chunk, GB rows at a time: their gathers are all in flight before the first image store
            static_assert(RC % GB == 0, "gather batches tile the chunk");
#pragma unroll
            for (int b = 0; b < RC; b += GB) {
                ImgRow v[GB];
#pragma unroll
                for (int q = 0; q < GB; ++q)
                    v[q] = img_gather<kHetero, kRing>(P, G, x, env_a, het, ring_a, r + b + q, e_base, (ng >> (b + q)) & 1u, (nw >> (b + q)) & 1u, sh);
#pragma unroll
                for (int q = 0; q < GB; ++q) img_scatter(x, im + (uint32_t)((b + q) * x.row_bytes), v[q]);
            }
        } else {                                    // the last rows of a partial tile, one by one
            n = r_end - r;
#pragma unroll 1
            for (int q = 0; q < n; ++q)
                img_scatter(x, im + (uint32_t)(q * x.row_bytes),
                            img_gather<kHetero, kRing>(P, G, x, env_a, het, ring_a, r + q, e_base, (ng >> q) & 1u, (nw >> q) & 1u, sh));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            if (MG_BULK_STORE_ENABLED)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(obs_tile + (size_t)r * (x.row_bytes >> 3)), "r"(im),
                             "r"((uint32_t)(n * x.row_bytes)) : "memory");
            bulk_commit();
        }
        buf = buf + 1 == NB ? 0 : buf + 1;
    }
}

template <bool kHetero>
__device__ __forceinline__ RawRow gather_raw(const LaunchParams &P, const DevGroup &G, const MgConfig *__restrict__ c, int e, int t) {
    RawRow r;
    r.load = __ldg(P.load_raw + (size_t)c->load_series * P.T + t);
    r.pv = __ldg(P.pv_raw + (size_t)c->pv_series * P.T + t);
    if (kHetero) {   // series value = profile value * scale (MicrogridGenerator grids)
        r.load *= c->load_scale;
        r.pv *= c->pv_scale;
    }
    if (G.has_grid) {
        const double2 *g2 = reinterpret_cast<const double2 *>(P.grid_raw + ((size_t)c->grid_series * P.T + t) * 4);
        const double2 a = __ldg(g2), b = __ldg(g2 + 1);
        r.imp = a.x; r.exp_ = a.y; r.co2 = b.x; r.status = b.y;
        if (kHetero && G.status_bits)
            r.status = (double)((__ldg(G.status_bits + (size_t)e * G.status_words + (t >> 5)) >> (t & 31)) & 1u);
    } else {
        r.imp = r.exp_ = r.co2 = 0.0;
        r.status = 1.0;
    }
    return r;
}

__device__ __forceinline__ void unpack_genset(uint32_t w, EnvRegs &s) {
    s.cs = w & 0xff; s.gs = (w >> 8) & 0xff; s.up = (w >> 16) & 0xff; s.dn = (w >> 24) & 0xff;
}
__device__ __forceinline__ uint32_t pack_genset(const EnvRegs &s) {
    return (uint32_t)s.cs | ((uint32_t)s.gs << 8) | ((uint32_t)s.up << 16) | ((uint32_t)s.dn << 24);
}

// Rollout log of selected envs (MgRolloutIO.log): log_open writes the state BEFORE the step and returns the record (or
// nullptr: env not logged), whose info block owner_step fills; log_close adds what is known after the step.
__device__ __forceinline__ double *log_open(const DevGroup &G, int slot, int step, int n_steps, const EnvRegs &s) {
    if (slot < 0) return nullptr;
    double *rec = G.log + ((size_t)slot * n_steps + step) * MG_N_LOG;
    rec[MG_LOG_STEP] = (double)s.t;
    rec[MG_LOG_CHARGE] = s.charge;
    rec[MG_LOG_GENSET_BEFORE] = (double)pack_genset(s);
    return rec;
}
__device__ __forceinline__ void log_close(double *rec, const EnvRegs &s, double reward, int done, uint32_t flags, bool valid) {
    if (!rec) return;
    rec[MG_LOG_GENSET_AFTER] = (double)pack_genset(s);
    rec[MG_LOG_REWARD] = reward;
    rec[MG_LOG_DONE] = (double)done;
    rec[MG_LOG_FLAGS] = (double)flags;
    rec[7] = 0.0;
    if (!valid)      // a rejected step (past the end, bad action) logs nothing in the reference: the info block stays empty
        for (int q = 0; q < MG_N_INFO; ++q) rec[MG_LOG_INFO + q] = 0.0;
}

__device__ __forceinline__ int find_group(const LaunchParams &P, int tile) {
    int g = 0;
#pragma unroll
    for (int q = 1; q < MG_MAX_GROUPS; ++q)
        if (q < P.n_groups && tile >= P.g[q].tile_begin) g = q;
    return g;
}

struct ActionRegs {
    double goal, gen, bat, grid;
};

// read one env's action row into the four logical controls (one or two vector loads where rows allow it); T = double, or
// float when the caller's actions are float32 (MG_OPT_ACTIONS_F32: widened exactly, the arithmetic stays f64)
template <typename T> struct ActVec;
template <> struct ActVec<double> { typedef double2 v2; };
template <> struct ActVec<float> { typedef float2 v2; };
template <typename T>
__device__ __forceinline__ ActionRegs read_action_t(const DevGroup &G, const T *__restrict__ row) {
    typedef typename ActVec<T>::v2 V2;
    ActionRegs a;
    a.goal = a.gen = a.grid = 0.0;
    if (G.n_act == 2) {          // battery + grid in either order: one vector load
        const V2 v = __ldg(reinterpret_cast<const V2 *>(row));
        a.bat = G.act_col_battery == 0 ? (double)v.x : (double)v.y;
        a.grid = G.act_col_grid == 0 ? (double)v.x : (double)v.y;
    } else if (G.n_act == 4) {   // two vector loads
        const V2 v0 = __ldg(reinterpret_cast<const V2 *>(row));
        const V2 v1 = __ldg(reinterpret_cast<const V2 *>(row) + 1);
        const double x0 = v0.x, y0 = v0.y, x1 = v1.x, y1 = v1.y;
        const int cg = G.act_col_genset, cb = G.act_col_battery, cr = G.act_col_grid;
        a.goal = cg == 0 ? x0 : cg == 1 ? y0 : x1;
        a.gen = cg == 0 ? y0 : cg == 1 ? x1 : y1;
        a.bat = cb == 0 ? x0 : cb == 1 ? y0 : cb == 2 ? x1 : y1;
        a.grid = cr == 0 ? x0 : cr == 1 ? y0 : cr == 2 ? x1 : y1;
    } else {
        a.bat = (double)__ldg(row + G.act_col_battery);
        if (G.has_genset) { a.goal = (double)__ldg(row + G.act_col_genset); a.gen = (double)__ldg(row + G.act_col_genset + 1); }
        if (G.has_grid) a.grid = (double)__ldg(row + G.act_col_grid);
    }
    return a;
}
__device__ __forceinline__ ActionRegs read_action(const DevGroup &G, size_t elem) {      // elem: index of the row's first element
    if (G.act_f32) return read_action_t<float>(G, reinterpret_cast<const float *>(G.actions) + elem);
    return read_action_t<double>(G, G.actions + elem);
}

// Inputs of one env-step, fetched by the owner thread BEFORE the tile's observation rows are streamed out so that
// their latency overlaps the stores.  `valid` is false where the reference raises before touching any state
// (IndexError past the end of the series, ValueError for an action outside the discrete space).
struct StepInputs {
    bool valid;
    uint32_t invalid_flag;
    int dact;
    ActionRegs act;
    RawRow raw;
};

template <bool kHetero>
__device__ __forceinline__ StepInputs fetch_inputs(const LaunchParams &P, const DevGroup &G, const MgConfig *__restrict__ c,
                                                   int e, int step, int t) {
    StepInputs in;
    in.valid = t < P.T;
    in.invalid_flag = in.valid ? 0u : (uint32_t)MG_FLAG_STEP_PAST_END;
    in.dact = 0;
    in.act.goal = in.act.gen = in.act.bat = in.act.grid = 0.0;
    if (P.mode == MODE_DISCRETE) {
        in.dact = __ldg(G.dactions + (size_t)step * G.dact_step_stride + e);
        if (in.valid && (in.dact < 0 || in.dact >= c->plist_count)) {
            in.valid = false;
            in.invalid_flag = MG_FLAG_BAD_ACTION;
        }
    } else {
        in.act = read_action(G, (size_t)step * G.act_step_stride + (size_t)e * G.n_act);
    }
    if (in.valid) in.raw = gather_raw<kHetero>(P, G, c, e, t);
    else in.raw.load = in.raw.pv = in.raw.imp = in.raw.exp_ = in.raw.co2 = in.raw.status = 0.0;
    return in;
}

// the physics of one env-step from pre-fetched inputs
__device__ __forceinline__ void owner_step(const LaunchParams &P, const DevGroup &G, const MgConfig *__restrict__ c, EnvRegs &s,
                                           const StepInputs &in, int final_step, double *__restrict__ info, double &reward,
                                           int &done, uint32_t &flags) {
    if (!in.valid) {
        reward = __longlong_as_double(0x7ff8000000000000LL);
        done = (in.invalid_flag == MG_FLAG_STEP_PAST_END) ? 1 : 0;
        flags = in.invalid_flag;
        return;
    }
    ActionRegs a = in.act;
    bool normalized = P.normalized != 0;
    uint32_t list_flags = 0;
    if (P.mode == MODE_DISCRETE) {
        normalized = false;
        list_flags = priority_control(P.plist[c->plist_offset + in.dact], c, G, s, in.raw, a.goal, a.gen, a.bat, a.grid);
    }
    env_step(c, G, s, in.raw, a.goal, a.gen, a.bat, a.grid, normalized, final_step, reward, done, flags, info);
    flags |= list_flags;
}

// ------------------------------------------------------------------------------------------------------------------
// fused single-step kernel: MODE_STEP / MODE_DISCRETE / MODE_OBSERVE / MODE_RESET
//   owners (one thread per env): inputs, physics, state / reward / done write-back, publish the env's record;
//   all warps: stream the tile's observation rows.
// The latency-bound part comes first and the stores last: stores are fire-and-forget, so a CTA retires as soon as
// its rows are issued and the drain overlaps the next launch (measured: 18.4 us/step against 22.9 the other way round).
// ------------------------------------------------------------------------------------------------------------------
// kHetero = true adds the per-env series paths (profile * scale, status bits); table-backed batches run the lean kernel
template <bool kHetero, typename TO>
__global__ void __launch_bounds__(MG_THREADS, kHetero ? MG_MIN_CTAS_HETERO : MG_MIN_CTAS) mg_step_kernel(const __grid_constant__ LaunchParams P) {
    __shared__ TileShared S;
    __shared__ typename HeteroStorage<kHetero>::type SH;
    HeteroEnv *het0 = HeteroStorage<kHetero>::rows(SH, 0);
    const int gi = find_group(P, blockIdx.x);
    const DevGroup &G = P.g[gi];
    const int e0 = (blockIdx.x - G.tile_begin) * MG_TILE;
    const int n_rows = min(MG_TILE, G.n_envs - e0);
    const int tid = threadIdx.x;
    const int e = e0 + tid;
    if ((tid & 31) == 0 && G.obs) mbar_init(&S.bar[tid >> 5], 1);
    // Step overlap (P.pdl, MG_OPT_STEP_OVERLAP): this grid may have started while the previous step launch is still streaming
    // its observation rows -- the hardware only starts it once EVERY CTA of that launch has passed its trigger below, i.e.
    // has written (and fenced) the state, reward and done this launch reads and overwrites.  The state is read past L1
    // (ld.global.cg): an SM may still hold lines an earlier launch cached.
    double my_reward = 0.0;
    bool stepped = false;
    if (tid < n_rows) {
        const MgConfig *__restrict__ c = P.cfg + __ldg(G.cfg_index + e);
        EnvRegs s;
        s.t = __ldcg(G.step + e);
        s.charge = __ldcg(G.charge + e);
        s.cs = s.gs = s.up = s.dn = 0;
        if (G.has_genset) unpack_genset(__ldcg(G.genset + e), s);
        if (P.mode == MODE_STEP || P.mode == MODE_DISCRETE) {
            const StepInputs in = fetch_inputs<kHetero>(P, G, c, e, 0, s.t);
            const int final_step = G.env_final ? __ldg(G.env_final + e) : c->final_step;
            double reward;
            int done;
            uint32_t flags;
            owner_step(P, G, c, s, in, final_step, G.info ? G.info + (size_t)e * MG_N_INFO : nullptr, reward, done, flags);
            if (in.valid) {
                G.step[e] = s.t;
                G.charge[e] = s.charge;
                if (G.has_genset) G.genset[e] = pack_genset(s);
            }
            G.reward[e] = reward;
            G.done[e] = (uint8_t)done;
            if (G.flags) G.flags[e] = flags;
            my_reward = reward;
            stepped = true;
        } else if (P.mode == MODE_RESET) {
            if (!G.mask || G.mask[e]) {   // Microgrid.reset: only the step counter moves (microgrid.py:205-225)
                s.t = G.env_initial ? __ldg(G.env_initial + e) : c->initial_step;
                G.step[e] = s.t;
            }
        }
        if (G.obs) {
            publish_env<kHetero>(S.env[0][tid], het0 ? het0 + tid : nullptr, c, G, s, P.T, P.Tp);
            if (P.mode >= MODE_OBSERVE && G.soc_reported) {
                // BatteryModule keeps the soc it was constructed with until its first _update_state (battery_module.py:89,
                // 125-130); init_soc * max_capacity / max_capacity is not always init_soc
                const double soc = __ldg(G.soc_reported + e);
                S.env[0][tid].state[(G.has_genset && G.state_genset_first) ? 4 : 0] = (soc - c->bat_soc_low) / c->bat_soc_spread;
            }
        }
    }
    if (G.reward_total && tid < MG_TILE) add_reward_total(G.reward_total, my_reward, stepped);   // warps 0..1, uniformly
    if (G.obs) __syncthreads();   // CTA-uniform: the tile records are published
    // Every write a following step depends on (state, reward, done, flags, info) is issued: once they are visible device-wide
    // (the fence; only the owner warps have anything pending) the trigger lets the next launch's CTAs start their
    // latency-bound part on SMs as they free up while this grid is still streaming observation rows.  The host only chains
    // launches that write different observation buffers (launch_step).
    if (P.pdl) {
        if (tid < MG_TILE) __threadfence();
        pdl_launch_dependents();
        // mode 1: this launch's rows wait for the launch it overlapped with to complete -- its physics ran under that
        // launch's row stream, the row streams themselves never overlap, so any choice of observation buffers is safe
        if (P.pdl == 1) pdl_wait();
    }
    if (G.obs) {
        uint32_t phase = 0;
        TO *obs_tile = reinterpret_cast<TO *>(G.obs) + (size_t)e0 * G.obs_dim;
        if (!G.long_path) warp_emit_rows<kHetero, TO>(P, G, S, het0, 0, obs_tile, n_rows, e0, phase, (tid >> 5) * MG_ROWS_PER_WARP);
        else warp_emit_rows_long<TO>(P, G, S, 0, obs_tile, n_rows);
    }
    // mode 2: row streams of consecutive launches overlap too (the host chains only launches whose observation buffers differ
    // from those of the two launches before); complete in launch order all the same
    if (P.pdl == 2) pdl_wait();
}

// ------------------------------------------------------------------------------------------------------------------
// persistent multi-step kernel: every CTA owns its tile for all n_steps; env state stays in registers
// ------------------------------------------------------------------------------------------------------------------
// kLog: record the per-step log of selected envs (MgRolloutIO.log).  Only this kernel family logs: the info block keeps ~16
// more doubles alive through the physics, which the register budgets of the role-split kernels have no room for -- a
// launch that asks for a log runs here.
template <bool kHetero, typename TO, bool kRing = false, bool kLog = false>
__global__ void __launch_bounds__(MG_THREADS, kRing ? MG_MIN_CTAS_RING : kHetero ? MG_MIN_CTAS_HETERO : MG_MIN_CTAS) mg_rollout_kernel(const __grid_constant__ LaunchParams P) {
    static_assert(!kRing || kHetero, "the ring path is a variant of the per-env series kernels");
    __shared__ TileShared S;
    __shared__ typename HeteroStorage<kHetero && !kRing>::type SH;    // (the ring variant needs no per-env series records)
    __shared__ typename RingStorage<kRing>::type SR;
    RingShared *rings = RingStorage<kRing>::get(SR);
    const int gi = find_group(P, blockIdx.x);
    const DevGroup &G = P.g[gi];
    const int e0 = (blockIdx.x - G.tile_begin) * MG_TILE;
    const int n_rows = min(MG_TILE, G.n_envs - e0);
    const int tid = threadIdx.x;
    const bool owner = tid < n_rows;
    const int e = e0 + tid;
    const MgConfig *__restrict__ c = P.cfg;
    EnvRegs s;
    s.t = 0; s.charge = 0.0; s.cs = s.gs = s.up = s.dn = 0;
    int final_step = 0;
    double rsum = 0.0;
    uint32_t fsum = 0;
    uint32_t phase = 0;
    if ((tid & 31) == 0 && G.obs) mbar_init(&S.bar[tid >> 5], 1);
    int log_slot = -1;
    if (owner) {
        c = P.cfg + __ldg(G.cfg_index + e);
        s.t = G.step[e];
        s.charge = G.charge[e];
        if (G.has_genset) unpack_genset(G.genset[e], s);
        final_step = G.env_final ? __ldg(G.env_final + e) : c->final_step;
        if (kLog && G.log_slot && G.log) log_slot = __ldg(G.log_slot + e);
    }
    // ring variant: fill every special env's windows for its current step [t, t + H], one env per warp pass, lane k = window
    // element k (the only place where a whole window is normalised); the owner then appends one value per step
    const int ring_R = G.horizon + 2;
    int ring_base = 0;
    bool ring_env = false;
    if (kRing && G.obs) {
        const int lane = tid & 31, w0 = (tid >> 5) * MG_ROWS_PER_WARP;
        for (int r = w0; r < min(w0 + MG_ROWS_PER_WARP, n_rows); ++r) {
            const MgConfig *__restrict__ cr = P.cfg + __ldg(G.cfg_index + e0 + r);
            if (!env_is_special(cr, G)) continue;
            const int t_r = min(G.step[e0 + r], P.T);
            for (int k = lane; k <= G.horizon; k += 32) {
                double ld, pv;
                series_obs_value(P, cr, t_r + k, ld, pv);
                const int slot = (t_r + k) % ring_R;
                rings->win[0][r][slot] = ld;
                rings->win[1][r][slot] = pv;
            }
        }
        if (owner) {
            ring_env = env_is_special(c, G);
            ring_base = min(s.t, P.T) % ring_R;
        }
        // (visibility: the first reader passes the __syncthreads of step 0 first)
    }
    // (requesting step s+1's inputs before streaming step s's rows was measured SLOWER -- 13.6 vs 12.2 us/step: the
    //  prefetched registers spill under the 72-register budget that keeps all 1 024 tiles resident)
    for (int step = 0; step < P.n_steps; ++step) {
        const int ebuf = step & 1;
        double my_reward = 0.0;
        if (owner) {
            const StepInputs in = fetch_inputs<kHetero>(P, G, c, e, step, s.t);
            double reward;
            int done;
            uint32_t flags;
            double *const log_rec = kLog ? log_open(G, log_slot, step, P.n_steps, s) : nullptr;
            owner_step(P, G, c, s, in, final_step, (kLog && log_rec) ? log_rec + MG_LOG_INFO : nullptr, reward, done, flags);
            if (kLog) log_close(log_rec, s, reward, done, flags, in.valid);
            G.reward[(size_t)step * G.out_step_stride + e] = reward;
            G.done[(size_t)step * G.out_step_stride + e] = (uint8_t)done;
            rsum += reward;
            fsum |= flags;
            my_reward = reward;
            if (G.obs) {
                publish_env<kHetero>(S.env[ebuf][tid], (kHetero && !kRing) ? HeteroStorage<kHetero && !kRing>::rows(SH, ebuf) + tid : nullptr, c, G, s, P.T, P.Tp);
                if (kRing && ring_env) {
                    if (in.valid) {   // the step counter moved from t to t + 1 <= T: the window gains index t + 1 + H
                        ring_base = ring_base + 1 == ring_R ? 0 : ring_base + 1;
                        double ld, pv;
                        series_obs_value(P, c, s.t + G.horizon, ld, pv);
                        int slot = ring_base + G.horizon;
                        if (slot >= ring_R) slot -= ring_R;
                        rings->win[0][tid][slot] = ld;
                        rings->win[1][tid][slot] = pv;
                    }
                    publish_ring(S.env[ebuf][tid], c, G, e, ring_base);
                }
            }
        }
        if (G.reward_total && tid < MG_TILE) add_reward_total(G.reward_total + step, my_reward, owner);
        if (G.obs) {
            // one barrier per step: the records of step s live in env[s & 1]; a warp can only reach the barrier of
            // step s+1 after it has finished reading env[s & 1], so the owners may overwrite it at step s+2
            __syncthreads();
            TO *obs_tile = reinterpret_cast<TO *>(G.obs) + (size_t)(step % P.ring) * G.obs_slot_stride + (size_t)e0 * G.obs_dim;
            if (!G.long_path) warp_emit_rows<kHetero, TO, kRing>(P, G, S, HeteroStorage<kHetero && !kRing>::rows(SH, ebuf), ebuf, obs_tile, n_rows, e0, phase, (tid >> 5) * MG_ROWS_PER_WARP, rings);
            else warp_emit_rows_long<TO>(P, G, S, ebuf, obs_tile, n_rows);
        }
    }
    if (owner) {
        G.step[e] = s.t;
        G.charge[e] = s.charge;
        if (G.has_genset) G.genset[e] = pack_genset(s);
        if (G.reward_sum) G.reward_sum[e] = rsum;
        if (G.flags) G.flags[e] = fsum;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// warp-specialised persistent kernel (used when observations are written and rows take the staged path):
//   warps 0-1  owners : physics of step s, publish the tile record into env[s & 1], signal FULL[s & 1]
//   warps 2-3  emitters: wait FULL[s & 1], stream the 64 rows of step s (32 rows per warp), signal EMPTY[s & 1]
// The co-resident CTAs of an SM run in lock-step, so without this split every CTA is in its latency-bound physics at the
// same time and nothing is storing; here the owners compute step s+1 while the emitters stream step s.
// Named barriers (bar.sync / bar.arrive, 128 threads each): a barrier completes when the 64 threads of one role have
// arrived and the 64 of the other role have synced.
// ------------------------------------------------------------------------------------------------------------------
enum { BAR_FULL0 = 1, BAR_FULL1 = 2, BAR_EMPTY0 = 3, BAR_EMPTY1 = 4 };
// immediate barrier ids (a register id would make ptxas reserve all 16 hardware barriers of the CTA)
__device__ __forceinline__ void named_bar_sync(int base, int buf) {
    if (base == BAR_FULL0) {
        if (buf) asm volatile("bar.sync 2, 128;" ::: "memory");
        else asm volatile("bar.sync 1, 128;" ::: "memory");
    } else {
        if (buf) asm volatile("bar.sync 4, 128;" ::: "memory");
        else asm volatile("bar.sync 3, 128;" ::: "memory");
    }
}
__device__ __forceinline__ void named_bar_arrive(int base, int buf) {
    if (base == BAR_FULL0) {
        if (buf) asm volatile("bar.arrive 2, 128;" ::: "memory");
        else asm volatile("bar.arrive 1, 128;" ::: "memory");
    } else {
        if (buf) asm volatile("bar.arrive 4, 128;" ::: "memory");
        else asm volatile("bar.arrive 3, 128;" ::: "memory");
    }
}

template <bool kHetero, typename TO>
__global__ void __launch_bounds__(MG_THREADS, kHetero ? MG_MIN_CTAS_HETERO : MG_MIN_CTAS) mg_rollout_ws_kernel(const __grid_constant__ LaunchParams P) {
    static_assert(MG_THREADS == 128 && MG_TILE == 64, "role split assumes 2 owner warps + 2 emitter warps");
    __shared__ TileShared S;
    __shared__ typename HeteroStorage<kHetero>::type SH;
    const int gi = find_group(P, blockIdx.x);
    const DevGroup &G = P.g[gi];
    const int e0 = (blockIdx.x - G.tile_begin) * MG_TILE;
    const int n_rows = min(MG_TILE, G.n_envs - e0);
    const int tid = threadIdx.x;
    if ((tid & 31) == 0) mbar_init(&S.bar[tid >> 5], 1);
    __syncthreads();
    if (tid < MG_TILE) {   // ---------------- owners ----------------
        const bool owner = tid < n_rows;
        const int e = e0 + tid;
        const MgConfig *__restrict__ c = P.cfg;
        EnvRegs s;
        s.t = 0; s.charge = 0.0; s.cs = s.gs = s.up = s.dn = 0;
        int final_step = 0;
        double rsum = 0.0;
        uint32_t fsum = 0;
        if (owner) {
            c = P.cfg + __ldg(G.cfg_index + e);
            s.t = G.step[e];
            s.charge = G.charge[e];
            if (G.has_genset) unpack_genset(G.genset[e], s);
            final_step = G.env_final ? __ldg(G.env_final + e) : c->final_step;
        }
        for (int step = 0; step < P.n_steps; ++step) {
            const int ebuf = step & 1;
            double my_reward = 0.0;
            int my_done = 0;
            StepInputs in;
            in.valid = false;
            if (owner) in = fetch_inputs<kHetero>(P, G, c, e, step, s.t);
            if (step >= 2) named_bar_sync(BAR_EMPTY0, ebuf);   // the emitters are done with env[ebuf] of step - 2
            if (owner) {
                double reward;
                int done;
                uint32_t flags;
                owner_step(P, G, c, s, in, final_step, nullptr, reward, done, flags);
                rsum += reward;
                fsum |= flags;
                my_reward = reward;
                my_done = done;
                publish_env<kHetero>(S.env[ebuf][tid], kHetero ? HeteroStorage<kHetero>::rows(SH, ebuf) + tid : nullptr, c, G, s, P.T, P.Tp);
            }
            // Hand the tile record over: when the barrier completes, this thread's prior shared-memory writes are performed
            // for the threads that sync on it (PTX barrier semantics), so no fence.  The step's global results are stored
            // AFTER the hand-over: a fence or barrier between them and the arrive would wait for their acknowledgement, which
            // takes microseconds while the observation stream saturates the memory system (it set the pace of the kernel).
            named_bar_arrive(BAR_FULL0, ebuf);
            if (owner) {
                G.reward[(size_t)step * G.out_step_stride + e] = my_reward;
                G.done[(size_t)step * G.out_step_stride + e] = (uint8_t)my_done;
            }
            if (G.reward_total) add_reward_total(G.reward_total + step, my_reward, owner);
        }
        if (owner) {
            G.step[e] = s.t;
            G.charge[e] = s.charge;
            if (G.has_genset) G.genset[e] = pack_genset(s);
            if (G.reward_sum) G.reward_sum[e] = rsum;
            if (G.flags) G.flags[e] = fsum;
        }
    } else {               // ---------------- emitters ----------------
        uint32_t phase = 0;
        const int half = (tid >> 5) - 2;   // 0 or 1: rows [32 half, 32 half + 32)
        for (int step = 0; step < P.n_steps; ++step) {
            const int ebuf = step & 1;
            named_bar_sync(BAR_FULL0, ebuf);
            TO *obs_tile = reinterpret_cast<TO *>(G.obs) + (size_t)(step % P.ring) * G.obs_slot_stride + (size_t)e0 * G.obs_dim;
            const HeteroEnv *het = HeteroStorage<kHetero>::rows(SH, ebuf);
#pragma unroll 1
            for (int q = 0; q < 2; ++q)
                warp_emit_rows<kHetero, TO>(P, G, S, het, ebuf, obs_tile, n_rows, e0, phase, 32 * half + MG_ROWS_PER_WARP * q);
            // (every read of env[ebuf] / het[ebuf] has returned: the stores that carry the values were issued; no fence, which
            //  would wait for those stores to be acknowledged)
            named_bar_arrive(BAR_EMPTY0, ebuf);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// persistent kernel with the image emitter (warp_emit_rows_img): rows leave the SM as TMA bulk stores of RC rows.
//   kWS = false  every warp: owners (threads < 64) run the physics, CTA barrier, each warp emits its 16 rows
//   kWS = true   warps 0-1 own the envs and run one step ahead, warps 2-3 emit 32 rows each (named barriers as in
//                mg_rollout_ws_kernel); only the two emitting warps have images
// kRing keeps the sliding load / pv windows of envs with their own series in shared memory (see RingShared).
// Dynamic shared memory: (kWS ? 2 : MG_WARPS) * NB * RC * obs_dim doubles (the images).
// ------------------------------------------------------------------------------------------------------------------
#ifndef MG_IMG_MIN_CTAS
#define MG_IMG_MIN_CTAS 3
#endif
extern __shared__ __align__(128) unsigned char mg_dyn_smem[];
#define MG_MAX_SMS 256
__device__ uint32_t mg_sm_cta_count[MG_MAX_SMS];   // CTAs of the role-split kernels seen by each SM so far (only the parity is used)


struct ImgTileShared {
    TileEnv env[2][MG_TILE];   // double buffered across steps
};

// single-step kernel with the image emitter: mg_step_kernel's phases, rows leave as bulk stores of RC rows
template <int RC, int NB, int GB, bool kHetero>
__global__ void __launch_bounds__(MG_THREADS, 5) mg_step_img_kernel(const __grid_constant__ LaunchParams P) {
    __shared__ TileEnv S_env[MG_TILE];
    __shared__ typename HeteroStorage<kHetero>::type SH;
    HeteroEnv *het0 = HeteroStorage<kHetero>::rows(SH, 0);
    const int gi = find_group(P, blockIdx.x);
    const DevGroup &G = P.g[gi];
    const int e0 = (blockIdx.x - G.tile_begin) * MG_TILE;
    const int n_rows = min(MG_TILE, G.n_envs - e0);
    const int tid = threadIdx.x;
    const int e = e0 + tid;
    double my_reward = 0.0;
    bool stepped = false;
    if (tid < n_rows) {
        const MgConfig *__restrict__ c = P.cfg + __ldg(G.cfg_index + e);
        EnvRegs s;
        s.t = __ldcg(G.step + e);            // (read past L1: see mg_step_kernel on step overlap)
        s.charge = __ldcg(G.charge + e);
        s.cs = s.gs = s.up = s.dn = 0;
        if (G.has_genset) unpack_genset(__ldcg(G.genset + e), s);
        if (P.mode == MODE_STEP || P.mode == MODE_DISCRETE) {
            const StepInputs in = fetch_inputs<kHetero>(P, G, c, e, 0, s.t);
            const int final_step = G.env_final ? __ldg(G.env_final + e) : c->final_step;
            double reward;
            int done;
            uint32_t flags;
            owner_step(P, G, c, s, in, final_step, G.info ? G.info + (size_t)e * MG_N_INFO : nullptr, reward, done, flags);
            if (in.valid) {
                G.step[e] = s.t;
                G.charge[e] = s.charge;
                if (G.has_genset) G.genset[e] = pack_genset(s);
            }
            G.reward[e] = reward;
            G.done[e] = (uint8_t)done;
            if (G.flags) G.flags[e] = flags;
            my_reward = reward;
            stepped = true;
        } else if (P.mode == MODE_RESET) {
            if (!G.mask || G.mask[e]) {   // Microgrid.reset: only the step counter moves (microgrid.py:205-225)
                s.t = G.env_initial ? __ldg(G.env_initial + e) : c->initial_step;
                G.step[e] = s.t;
            }
        }
        if (G.obs) {
            publish_env<kHetero>(S_env[tid], het0 ? het0 + tid : nullptr, c, G, s, P.T, P.Tp);
            if (P.mode >= MODE_OBSERVE && G.soc_reported) {
                // BatteryModule keeps the soc it was constructed with until its first _update_state (battery_module.py:89, 125-130)
                const double soc = __ldg(G.soc_reported + e);
                S_env[tid].state[(G.has_genset && G.state_genset_first) ? 4 : 0] = (soc - c->bat_soc_low) / c->bat_soc_spread;
            }
        }
    }
    if (G.reward_total && tid < MG_TILE) add_reward_total(G.reward_total, my_reward, stepped);   // warps 0..1, uniformly
    if (G.obs) __syncthreads();   // CTA-uniform: the tile records are published
    if (P.pdl) {   // step overlap, as in mg_step_kernel: hand the state on, then (mode 1) let the previous launch finish its rows
        if (tid < MG_TILE) __threadfence();
        pdl_launch_dependents();
        if (P.pdl == 1) pdl_wait();
    }
    if (G.obs) {
        const ImgCtx x = img_ctx(P, G);
        int buf = 0;
        const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // (provably warp-uniform: bulk-store operands in uniform registers)
        const uint32_t img_a = smem_u32(mg_dyn_smem) + (uint32_t)(warp * NB * RC * x.row_bytes);
        warp_emit_rows_img<RC, NB, GB, kHetero, false>(P, G, x, smem_u32(S_env), het0, 0u, img_a, buf, G.obs + (size_t)e0 * G.obs_dim, n_rows,
                                                   warp * MG_ROWS_PER_WARP, MG_ROWS_PER_WARP, e0);
        if ((tid & 31) == 0) bulk_wait_read<0>();   // shared memory must outlive the bulk stores' reads
    }
    if (P.pdl == 2) pdl_wait();
}

template <int RC, int NB, int GB, bool kHetero, bool kRing, bool kWS, int MINB = 0>
__global__ void __launch_bounds__(MG_THREADS, MINB ? MINB : !kWS ? MG_IMG_MIN_CTAS : !kHetero ? MG_MIN_CTAS : kRing ? 4 : 5) mg_rollout_img_kernel(const __grid_constant__ LaunchParams P) {
    static_assert(!kRing || kHetero, "the ring path is a variant of the per-env series kernels");
    static_assert(MG_THREADS == 128 && MG_TILE == 64, "role split assumes 2 owner warps + 2 emitter warps");
    __shared__ ImgTileShared S;
    __shared__ typename HeteroStorage<kHetero && !kRing>::type SH;
    __shared__ typename RingStorage<kRing>::type SR;
    RingShared *rings = RingStorage<kRing>::get(SR);
    const int gi = find_group(P, blockIdx.x);
    const DevGroup &G = P.g[gi];
    const int e0 = (blockIdx.x - G.tile_begin) * MG_TILE;
    const int n_rows = min(MG_TILE, G.n_envs - e0);
    // Role-relative thread id: owners are tid < 64.  With the role split, every other CTA of an SM gives the owner role to
    // warps 2-3: a warp's scheduler is its index mod 4, so without the flip every emitting warp of the SM would sit on the same
    // two of the four schedulers.
    // (The CTAs of one SM are counted through a per-SM counter: block indices alone do not alternate on an SM -- the
    //  hardware deals consecutive indices to consecutive SMs, and the SM count is even.)
    __shared__ uint32_t role_flip;
    if (kWS) {
        if (threadIdx.x == 0) {
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            role_flip = atomicAdd(&mg_sm_cta_count[smid & (MG_MAX_SMS - 1)], 1u) & 1u;
        }
        __syncthreads();
    }
    const int tid = kWS ? (int)(threadIdx.x ^ (role_flip << 6)) : (int)threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // (provably warp-uniform: bulk-store operands in uniform registers)
    const ImgCtx x = img_ctx(P, G);
    const uint32_t ring_a = kRing ? smem_u32(rings) : 0u;
    // ring variant: fill every special env's windows for its current step [t, t + H], one env per warp pass, lane k = window
    // element k (the only place where a whole window is normalised); the owner then appends one value per step
    const int ring_R = G.horizon + 2;
    if (kRing) {
        const int lane = tid & 31, w0 = warp * MG_ROWS_PER_WARP;
        for (int r = w0; r < min(w0 + MG_ROWS_PER_WARP, n_rows); ++r) {
            const MgConfig *__restrict__ cr = P.cfg + __ldg(G.cfg_index + e0 + r);
            if (!env_is_special(cr, G)) continue;
            const int t_r = min(G.step[e0 + r], P.T);
            for (int k = lane; k <= G.horizon; k += 32) {
                double ld, pv;
                series_obs_value(P, cr, t_r + k, ld, pv);
                const int slot = (t_r + k) % ring_R;
                rings->win[0][r][slot] = ld;
                rings->win[1][r][slot] = pv;
            }
        }
    }
    __syncthreads();
    if (!kWS || tid < MG_TILE) {   // ---------------- owners (and, without the role split, emitters too) ----------------
        const bool owner = tid < n_rows;
        const int e = e0 + tid;
        // Per-env parameter records (one MgConfig per env: MicrogridGenerator grids) are copied out of the 336-byte-stride
        // array once, into a per-thread copy that the compiler keeps in registers (or, where they run out, in this thread's
        // local memory: coalesced across the warp, L1 resident) -- instead of 32 different cache lines per field and warp
        // from L2, every step.  Table-backed batches share a few records per tile and read them through L1.
        MgConfig cl;
        const MgConfig *__restrict__ const c = kHetero ? &cl : P.cfg + (owner ? __ldg(G.cfg_index + e) : 0);
        if (kHetero) {
            const double2 *__restrict__ src = reinterpret_cast<const double2 *>(P.cfg + (owner ? __ldg(G.cfg_index + e) : 0));
            double2 *dst = reinterpret_cast<double2 *>(&cl);
#pragma unroll
            for (int q = 0; q < (int)(sizeof(MgConfig) / sizeof(double2)); ++q) dst[q] = __ldg(src + q);
        }
        EnvRegs s;
        s.t = 0; s.charge = 0.0; s.cs = s.gs = s.up = s.dn = 0;
        int final_step = 0, ring_base = 0, buf = 0;
        bool ring_env = false;
        double rsum = 0.0;
        uint32_t fsum = 0;
        if (owner) {
            s.t = G.step[e];
            s.charge = G.charge[e];
            if (G.has_genset) unpack_genset(G.genset[e], s);
            final_step = G.env_final ? __ldg(G.env_final + e) : c->final_step;
            if (kRing) {
                ring_env = env_is_special(c, G);
                ring_base = min(s.t, P.T) % ring_R;
            }
        }
        const uint32_t img_a = smem_u32(mg_dyn_smem) + (uint32_t)(warp * NB * RC * x.row_bytes);   // (not used by owners of the split)
#ifdef MG_ROLE_TIMERS
        long long tm_phys = 0, tm_empty = 0;
#endif
        // The inputs of a step depend only on its index and on the env's step counter, which moves by one per valid step:
        // they are requested one step ahead (at the top of step s for step s + 1), so that their memory latency -- several
        // dependent trips to L2 / HBM that take microseconds while the stores saturate the memory system -- overlaps the
        // physics of step s.  A counter that did not move as predicted (rejected discrete action) re-fetches.
        StepInputs in_next;
        in_next.valid = false;
        if (kHetero && owner) in_next = fetch_inputs<kHetero>(P, G, c, e, 0, s.t);
        for (int step = 0; step < P.n_steps; ++step) {
            const int ebuf = step & 1;
            double my_reward = 0.0;
            int my_done = 0;
            // (table-backed batches run under a 72-register budget -- 7 CTAs per SM keep all 1 024 tiles of the 65 536-env
            //  batch resident -- where holding a second set of inputs spills: they fetch at the top of the step)
            if (!kHetero && owner) in_next = fetch_inputs<kHetero>(P, G, c, e, step, s.t);
            const StepInputs in = in_next;
#ifdef MG_ROLE_TIMERS
            const long long tp0 = clock64();
#endif
            const int t_pred = s.t + (in.valid ? 1 : 0);
            SeriesRaw ring_raw;
            uint2 status_w = make_uint2(0u, 0u);
            ring_raw.a = ring_raw.b = 0.0; ring_raw.in = false;
            if (owner) {
                if (kHetero && step + 1 < P.n_steps) in_next = fetch_inputs<kHetero>(P, G, c, e, step + 1, t_pred);
                if (kRing && ring_env) {    // this step's ring append and status window, requested before the physics
                    if (in.valid) ring_raw = series_obs_fetch(P, c, t_pred + G.horizon);
                    if (G.has_grid && G.status_bits) status_w = status_words_fetch(G, e, min(t_pred, P.T));
                }
            }
#ifdef MG_ROLE_TIMERS
            const long long te0 = clock64();
#endif
            if (kWS && step >= 2) named_bar_sync(BAR_EMPTY0, ebuf);   // the emitters are done with env[ebuf] of step - 2
#ifdef MG_ROLE_TIMERS
            const long long te1 = clock64();
            tm_empty += te1 - te0;
#endif
            if (owner) {
                double reward;
                int done;
                uint32_t flags;
                owner_step(P, G, c, s, in, final_step, nullptr, reward, done, flags);
                rsum += reward;
                fsum |= flags;
                my_reward = reward;
                my_done = done;
                if (kHetero && s.t != t_pred && step + 1 < P.n_steps) in_next = fetch_inputs<kHetero>(P, G, c, e, step + 1, s.t);
                publish_env<kHetero>(S.env[ebuf][tid], (kHetero && !kRing) ? HeteroStorage<kHetero && !kRing>::rows(SH, ebuf) + tid : nullptr, c, G, s, P.T, P.Tp);
                if (kRing && ring_env) {
                    if (in.valid) {   // the step counter moved from t to t + 1 <= T: the window gains index t + 1 + H
                        ring_base = ring_base + 1 == ring_R ? 0 : ring_base + 1;
                        double ld, pv;
                        series_obs_finish(c, ring_raw, ld, pv);
                        int slot = ring_base + G.horizon;
                        if (slot >= ring_R) slot -= ring_R;
                        rings->win[0][tid][slot] = ld;
                        rings->win[1][tid][slot] = pv;
                    }
                    // TileEnv of a ring row: see publish_ring (the status window word comes from the words requested above)
                    TileEnv &te = S.env[ebuf][tid];
                    te.off_load = (G.has_grid && G.status_bits) ? (int32_t)__funnelshift_r(status_w.x, status_w.y, te.special & 31) : 0;
                    te.off_pv = ring_base | (c->grid_status_weak ? (1 << 16) : 0);
                }
            }
#ifdef MG_ROLE_TIMERS
            tm_phys += (clock64() - tp0) - (te1 - te0);
#endif
            if (kWS) {
                // hand-over first, global results after it (see mg_rollout_ws_kernel)
                named_bar_arrive(BAR_FULL0, ebuf);
                if (owner) {
                    G.reward[(size_t)step * G.out_step_stride + e] = my_reward;
                    G.done[(size_t)step * G.out_step_stride + e] = (uint8_t)my_done;
                }
                if (G.reward_total && tid < MG_TILE) add_reward_total(G.reward_total + step, my_reward, owner);
            } else {
                if (owner) {
                    G.reward[(size_t)step * G.out_step_stride + e] = my_reward;
                    G.done[(size_t)step * G.out_step_stride + e] = (uint8_t)my_done;
                }
                if (G.reward_total && tid < MG_TILE) add_reward_total(G.reward_total + step, my_reward, owner);
                // one barrier per step: the records of step s live in env[s & 1]; a warp can only reach the barrier of
                // step s+1 after it has finished reading env[s & 1], so the owners may overwrite it at step s+2
                __syncthreads();
                double *obs_tile = G.obs + (size_t)(step % P.ring) * G.obs_slot_stride + (size_t)e0 * G.obs_dim;
                warp_emit_rows_img<RC, NB, GB, kHetero, kRing>(P, G, x, smem_u32(S.env[ebuf]), HeteroStorage<kHetero && !kRing>::rows(SH, ebuf), ring_a,
                                                             img_a, buf, obs_tile, n_rows, warp * MG_ROWS_PER_WARP, MG_ROWS_PER_WARP, e0);
                if (P.ring == 1 && (tid & 31) == 0) bulk_wait_all();   // the next step rewrites the same rows: keep the order
            }
        }
        if (owner) {
            G.step[e] = s.t;
            G.charge[e] = s.charge;
            if (G.has_genset) G.genset[e] = pack_genset(s);
            if (G.reward_sum) G.reward_sum[e] = rsum;
            if (G.flags) G.flags[e] = fsum;
        }
        if (!kWS && (tid & 31) == 0) bulk_wait_read<0>();   // shared memory must outlive the last bulk stores' reads
#ifdef MG_ROLE_TIMERS
        if (kWS && (tid & 31) == 0) {
            atomicAdd(&mg_role_cycles[0], (unsigned long long)tm_phys);
            atomicAdd(&mg_role_cycles[1], (unsigned long long)tm_empty);
            atomicAdd(&mg_role_cycles[5], (unsigned long long)P.n_steps);
        }
#endif
    } else {               // ---------------- emitters (kWS) ----------------
        const int half = warp - 2;   // 0 or 1: rows [32 half, 32 half + 32)
        const uint32_t img_a = smem_u32(mg_dyn_smem) + (uint32_t)(half * NB * RC * x.row_bytes);
        int buf = 0;
#ifdef MG_ROLE_TIMERS
        long long tm_full = 0, tm_emit = 0;
#endif
        for (int step = 0; step < P.n_steps; ++step) {
            const int ebuf = step & 1;
#ifdef MG_ROLE_TIMERS
            const long long tf0 = clock64();
#endif
            named_bar_sync(BAR_FULL0, ebuf);
#ifdef MG_ROLE_TIMERS
            const long long tf1 = clock64();
            tm_full += tf1 - tf0;
#endif
            double *obs_tile = G.obs + (size_t)(step % P.ring) * G.obs_slot_stride + (size_t)e0 * G.obs_dim;
            warp_emit_rows_img<RC, NB, GB, kHetero, kRing>(P, G, x, smem_u32(S.env[ebuf]), HeteroStorage<kHetero && !kRing>::rows(SH, ebuf), ring_a, img_a,
                                                         buf, obs_tile, n_rows, 32 * half, 32, e0);
            if (P.ring == 1 && (tid & 31) == 0) bulk_wait_all();
            // (every read of env[ebuf] / het[ebuf] / the rings has returned: the values are in the images)
#ifdef MG_ROLE_TIMERS
            tm_emit += clock64() - tf1;
#endif
            named_bar_arrive(BAR_EMPTY0, ebuf);
        }
#ifdef MG_ROLE_TIMERS
        if ((tid & 31) == 0) {
            atomicAdd(&mg_role_cycles[2], (unsigned long long)tm_full);
            atomicAdd(&mg_role_cycles[4], (unsigned long long)tm_emit);
        }
#endif
        if ((tid & 31) == 0) bulk_wait_read<0>();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// table construction (mg_create): bounds + normalised, end-padded observation tables
//   load / pv : low = min(ts), high = max(ts), pulled to include 0 (base_timeseries_module.py:81-88)
//   grid cols : per-column min / max                               (grid_module.py:125-132)
//   value     : (ts - low) / spread, spread 0 -> 1                 (utils/space.py:204-218)
//   rows >= T : the forecaster's fill (high + low) / 2, clipped    (forecast/forecaster.py:95, 120-149)
// one CTA per series column
// ------------------------------------------------------------------------------------------------------------------
struct TableParams {
    int32_t T, Tp, n_load, n_pv, n_grid;
    const double *load_raw, *pv_raw, *grid_raw;
    double *load_nrm, *pv_nrm, *grid_nrm, *bounds;
};

__global__ void __launch_bounds__(256) mg_build_tables_kernel(const TableParams P) {
    __shared__ double s_min[256], s_max[256];
    const int col = blockIdx.x;
    const double *src;
    double *dst;
    int stride;
    bool pull_zero;
    if (col < P.n_load) {
        src = P.load_raw + (size_t)col * P.T; dst = P.load_nrm + (size_t)col * P.Tp; stride = 1; pull_zero = true;
    } else if (col < P.n_load + P.n_pv) {
        const int k = col - P.n_load;
        src = P.pv_raw + (size_t)k * P.T; dst = P.pv_nrm + (size_t)k * P.Tp; stride = 1; pull_zero = true;
    } else {
        const int k = col - P.n_load - P.n_pv;
        src = P.grid_raw + (size_t)(k >> 2) * P.T * 4 + (k & 3);
        dst = P.grid_nrm + (size_t)(k >> 2) * P.Tp * 4 + (k & 3);
        stride = 4; pull_zero = false;
    }
    double mn = src[0], mx = src[0];
    for (int r = threadIdx.x; r < P.T; r += blockDim.x) {
        const double v = src[(size_t)r * stride];
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    s_min[threadIdx.x] = mn;
    s_max[threadIdx.x] = mx;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) {
            s_min[threadIdx.x] = fmin(s_min[threadIdx.x], s_min[threadIdx.x + w]);
            s_max[threadIdx.x] = fmax(s_max[threadIdx.x], s_max[threadIdx.x + w]);
        }
        __syncthreads();
    }
    double low = s_min[0], high = s_max[0];
    if (pull_zero) {
        if (low > 0) low = 0;
        else if (high < 0) high = 0;
    }
    double spread = high - low;
    if (spread == 0.0) spread = 1.0;
    double fill = (high + low) / 2;
    if (fill < low) fill = low;
    if (fill > high) fill = high;
    for (int r = threadIdx.x; r < P.Tp; r += blockDim.x) {
        const double v = r < P.T ? src[(size_t)r * stride] : fill;
        dst[(size_t)r * stride] = (v - low) / spread;
    }
    if (threadIdx.x == 0 && P.bounds) {
        P.bounds[2 * col] = low;
        P.bounds[2 * col + 1] = high;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// GaussianNoiseForecaster as a post-pass over freshly written observation rows (include/pymgrid_b200.h, mg_forecast_noise)
// reference: forecast/forecaster.py:220-262 (noise), :120-149 (pad rows carry no noise, clip to the bounds)
// One warp per env row; lane l draws for element pairs l, l + 32, ... of the row's forecast entries.
// ------------------------------------------------------------------------------------------------------------------
struct NoiseGroup {
    int32_t n_envs, obs_dim, horizon, has_grid, load_start, pv_start, grid_start, first_block;
    int64_t env_base;
    const int32_t *step, *cfg_index;
    const int32_t *step_base;           // mg_forecast_noise_at: the env's step counter before a rollout, or NULL
    void *obs;
};
struct NoiseParams {
    int32_t n_groups, T;
    int32_t step_add, _pad0;            // with step_base: the row observes step min(step_base[e] + step_add, T)
    uint32_t k0, k1, c3, _pad;
    const MgForecastNoise *noise;
    NoiseGroup g[MG_MAX_GROUPS];
};

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), counter c, key (k0, k1)
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        c[0] = hi1 ^ c[1] ^ k0;
        c[1] = lo1;
        c[2] = hi0 ^ c[3] ^ k1;
        c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

#define MG_NOISE_ROWS 16      // rows per CTA of the noise post-pass: (row, element pair) items are dealt to the 128 threads in turn

template <typename TO>
__global__ void __launch_bounds__(128) mg_forecast_noise_kernel(const __grid_constant__ NoiseParams P) {
    int gi = 0;
#pragma unroll
    for (int q = 1; q < MG_MAX_GROUPS; ++q)
        if (q < P.n_groups && (int)blockIdx.x >= P.g[q].first_block) gi = q;
    const NoiseGroup &G = P.g[gi];
    const int H = G.horizon;
    // increase_uncertainty: the scale of forecast row k, 1 + log(1 + k) (forecaster.py:244-248), depends on k alone -- one
    // evaluation per CTA instead of one f64 logarithm per element (the same expression, so the same bits)
    __shared__ double s_scale[128];
    __shared__ int s_t[MG_NOISE_ROWS], s_real[MG_NOISE_ROWS];
    const int tid = threadIdx.x;
    const int e0 = ((int)blockIdx.x - G.first_block) * MG_NOISE_ROWS;
    if (tid < H) s_scale[tid] = 1.0 + log(1.0 + (double)tid);
    if (tid < MG_NOISE_ROWS) {
        int t = 0, n_real = 0;
        if (e0 + tid < G.n_envs) {
            // the step the row observes (post-step state): the env's counter, or -- for a slot of a rollout's observation ring --
            // what the counter was when the slot was written (it moves by one per step and stops at the end of the series)
            t = G.step_base ? min(G.step_base[e0 + tid] + P.step_add, P.T) : G.step[e0 + tid];
            n_real = P.T - (t + 1);           // forecast rows that exist in the series: t + 1 + k < T
            n_real = n_real < 0 ? 0 : (n_real > H ? H : n_real);      // past the end everything is padding (base_timeseries_module.py:113-116)
        }
        s_t[tid] = t;
        s_real[tid] = n_real;
    }
    __syncthreads();
    const int n_fc = H * (2 + 4 * G.has_grid);
    const int n_pairs = (n_fc + 1) >> 1;
    // one Philox block = two normals = two neighbouring forecast elements of one row; a row's pairs rarely fill whole warps
    // (69 at H = 23), so the (row, pair) items of the CTA's rows are flattened over its threads
    for (int i = tid; i < MG_NOISE_ROWS * n_pairs; i += 128) {
        const int r = i / n_pairs, p = i - r * n_pairs;
        const int n_real = s_real[r];
        if (n_real == 0) continue;
        const int e = e0 + r, t = s_t[r];
        const MgForecastNoise *__restrict__ nz = P.noise + G.cfg_index[e];
        TO *__restrict__ row = reinterpret_cast<TO *>(G.obs) + (size_t)e * G.obs_dim;
        const uint64_t env = (uint64_t)(G.env_base + e);
        uint32_t c[4] = {(uint32_t)env, ((uint32_t)(env >> 32) & 0xffffu) | ((uint32_t)p << 16), (uint32_t)t, P.c3};
        philox4x32_10(c, P.k0, P.k1);
        // two 53-bit uniforms, u1 in (0, 1], u2 in [0, 1); Box-Muller
        const double u1 = ((double)(c[0] >> 5) * 67108864.0 + (double)(c[1] >> 6) + 1.0) * (1.0 / 9007199254740992.0);
        const double u2 = ((double)(c[2] >> 5) * 67108864.0 + (double)(c[3] >> 6)) * (1.0 / 9007199254740992.0);
        const double radius = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int f = 2 * p + h;
            if (f >= n_fc) break;
            int k, off, inc;
            double sigma;
            if (f < H) { k = f; off = G.load_start + 1 + k; sigma = nz->load_sigma; inc = nz->load_increase; }
            else if (f < 2 * H) { k = f - H; off = G.pv_start + 1 + k; sigma = nz->pv_sigma; inc = nz->pv_increase; }
            else {
                const int q = f - 2 * H;
                k = q >> 2; off = G.grid_start + 4 + q; sigma = nz->grid_sigma[q & 3]; inc = nz->grid_increase;
            }
            if (k >= n_real || sigma == 0.0) continue;
            if (inc) sigma = sigma * (k < 128 ? s_scale[k] : 1.0 + log(1.0 + (double)k));
            double v = (double)row[off] + (radius * (h ? sn : cs)) * sigma;
            v = fmin(fmax(v, 0.0), 1.0);                                // the clip to [low, high] in normalised units
            row[off] = (TO)v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side: the C-ABI
// ------------------------------------------------------------------------------------------------------------------
struct MgHandle {
    MgLayout layout;
    LaunchParams base;      // groups + tables, io fields cleared
    int64_t launches;
    // programmatic dependent launch bookkeeping: observation buffers and stream of the previous step launch
    const double *last_obs[MG_MAX_GROUPS];
    void *last_stream;
    bool last_was_step;
    bool hetero;            // per-env series (profile * scale) or per-env grid status present: run the kHetero kernels
    bool obs_f32;           // observation buffers are float32 (MG_LAYOUT_OBS_F32)
    bool rollout_specialised;   // MG_OPT_ROLLOUT_SPECIALISED
    bool rollout_ring;          // MG_OPT_ROLLOUT_RING
    int emit_image;             // MG_OPT_EMIT_IMAGE: 0 LSU row emitters, 1 image + TMA bulk stores, 2 choose per launch (default)
    bool ragged_hint;           // MG_OPT_RAGGED_HINT: the envs of a tile are (probably) at unrelated steps
    int actions_f32;            // MG_OPT_ACTIONS_F32: every `actions` pointer the handle is given holds float32
    int step_overlap;           // MG_OPT_STEP_OVERLAP: consecutive mg_step launches overlap (programmatic dependent launch): 0, 1, 2
    const double *prev_obs[MG_MAX_GROUPS];   // observation buffers of the launch before the last one
    int n_sms;                  // multiprocessors of the device the handle was created on
    const char *last_kernel;    // mg_last_kernel
    int image_shape;            // MG_OPT_IMAGE_SHAPE: index into the instantiated (rows per bulk store, buffers) shapes
    struct HostStage *stage;    // mg_rollout_host: streams, events and device staging (lazy)
    const double *soc_reported[MG_MAX_GROUPS];   // mg_set_reported_soc; dropped by the first step / rollout
};

// Device staging of mg_rollout_host: two slots of `chunk` steps each (actions in, reward + done out) for every group,
// two copy streams and the events that order them against the caller's stream.
struct HostStage {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t uploaded[2] = {nullptr, nullptr}, computed[2] = {nullptr, nullptr}, drained[2] = {nullptr, nullptr};
    cudaEvent_t fence = nullptr;            // start of a call on the caller's stream / end of the previous call
    bool fence_recorded = false;
    char *in[2] = {nullptr, nullptr}, *out[2] = {nullptr, nullptr};
    size_t in_cap = 0, out_cap = 0;         // bytes per slot
};

static void stage_free(HostStage *st) {
    if (!st) return;
    for (int k = 0; k < 2; ++k) {
        if (st->in[k]) cudaFree(st->in[k]);
        if (st->out[k]) cudaFree(st->out[k]);
        if (st->uploaded[k]) cudaEventDestroy(st->uploaded[k]);
        if (st->computed[k]) cudaEventDestroy(st->computed[k]);
        if (st->drained[k]) cudaEventDestroy(st->drained[k]);
    }
    if (st->fence) cudaEventDestroy(st->fence);
    if (st->s_in) cudaStreamDestroy(st->s_in);
    if (st->s_out) cudaStreamDestroy(st->s_out);
    delete st;
}

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof g_err, fmt, detail);
    return code;
}

// the other translation units of this library (mg_compose.cu) report through the same text
int mg_set_error(int code, const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}

static int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    return MG_E_CUDA;
}

extern "C" int mg_abi_version(void) { return MG_ABI_VERSION; }

extern "C" int64_t mg_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(MgConfig);
        case 1: return sizeof(MgPriorityList);
        case 2: return sizeof(MgGroup);
        case 3: return sizeof(MgLayout);
        case 4: return sizeof(MgStepIO);
        case 5: return sizeof(MgRolloutIO);
        case 6: return sizeof(MgForecastNoise);
        case 7: return sizeof(MgHostRolloutIO);
        default: return -1;
    }
}

extern "C" const char *mg_build_info(void) {
    return "pymgrid_b200 engine, sm_100a, f64, -fmad=false, tile=" MG_STR(MG_TILE) " threads=" MG_STR(MG_THREADS);
}

extern "C" const char *mg_last_error(void) { return g_err; }

static void layout_segments(const MgGroup &g, DevGroup &d) {
    const int rows = 1 + g.horizon;
    int n = 0;
    int kinds[5], lens[5];
    if (g.obs_order == MG_OBS_GYM_SORTED || g.obs_order == MG_OBS_GYM_SORTED_PV_FIRST) {
        if (g.obs_order == MG_OBS_GYM_SORTED_PV_FIRST) { kinds[n] = KIND_PV; lens[n++] = rows; }
        kinds[n] = KIND_BAT; lens[n++] = 2;
        if (g.has_genset) { kinds[n] = KIND_GEN; lens[n++] = 4; }
        if (g.has_grid) { kinds[n] = KIND_GRID; lens[n++] = 4 * rows; }
        kinds[n] = KIND_LOAD; lens[n++] = rows;
        if (g.obs_order == MG_OBS_GYM_SORTED) { kinds[n] = KIND_PV; lens[n++] = rows; }
    } else {
        kinds[n] = KIND_LOAD; lens[n++] = rows;
        kinds[n] = KIND_PV; lens[n++] = rows;
        if (g.has_genset) { kinds[n] = KIND_GEN; lens[n++] = 4; }
        kinds[n] = KIND_BAT; lens[n++] = 2;
        if (g.has_grid) { kinds[n] = KIND_GRID; lens[n++] = 4 * rows; }
    }
    int start = 0;
    for (int k = 0; k < 5; ++k) {
        d.seg_start[k] = k < n ? start : 0x7fffffff;
        d.seg_kind[k] = k < n ? kinds[k] : KIND_PV;
        if (k < n) start += lens[k];
    }
    d.n_seg = n;
    d.tma_ok = 1;
    for (int k = 0; k < n; ++k) {
        if (kinds[k] == KIND_GRID && (d.seg_start[k] & 1)) d.tma_ok = 0;   // bulk copies need 16-byte aligned addresses
        if ((kinds[k] == KIND_BAT || kinds[k] == KIND_GEN) && (k == 0 || (kinds[k - 1] != KIND_BAT && kinds[k - 1] != KIND_GEN))) {
            d.state_start = d.seg_start[k];
            d.state_genset_first = kinds[k] == KIND_GEN;
        }
    }
}

extern "C" int mg_create(const MgLayout *L, void *stream, MgHandle **out) {
    if (!L || !out) return fail(MG_E_INVALID, "mg_create: null argument");
    if (L->abi_version != MG_ABI_VERSION) return fail(MG_E_INVALID, "mg_create: abi_version mismatch");
    if (L->n_groups < 1 || L->n_groups > MG_MAX_GROUPS) return fail(MG_E_INVALID, "mg_create: n_groups out of range");
    if (L->series_len < 1 || L->max_horizon < 0 || L->n_cfg < 1 || L->n_load < 1 || L->n_pv < 1 || L->n_grid < 0)
        return fail(MG_E_INVALID, "mg_create: bad table sizes");
    if (!L->cfg || !L->load_raw || !L->pv_raw || !L->load_nrm || !L->pv_nrm || (L->n_grid > 0 && (!L->grid_raw || !L->grid_nrm)))
        return fail(MG_E_INVALID, "mg_create: null table pointer");
    const int64_t Tp = (int64_t)L->series_len + L->max_horizon + 1;
    if ((int64_t)L->n_load * Tp >= (1ll << 31) || (int64_t)L->n_pv * Tp >= (1ll << 31) || (int64_t)L->n_grid * Tp * 4 >= (1ll << 31))
        return fail(MG_E_UNSUPPORTED, "mg_create: series tables exceed 2^31 elements");
    MgHandle *h = new (std::nothrow) MgHandle();
    if (!h) return fail(MG_E_INVALID, "mg_create: out of host memory");
    h->layout = *L;
    h->launches = 0;
    h->last_was_step = false;
    h->hetero = (L->flags & MG_LAYOUT_SCALED_SERIES) != 0;
    h->obs_f32 = (L->flags & MG_LAYOUT_OBS_F32) != 0;
    h->rollout_specialised = true;
    h->rollout_ring = true;
    h->emit_image = 2;
    h->image_shape = -1;
    h->ragged_hint = false;
    h->step_overlap = 1;
    h->actions_f32 = 0;
    for (int g = 0; g < MG_MAX_GROUPS; ++g) h->prev_obs[g] = nullptr;
    h->last_kernel = "";
    h->n_sms = 148;
    {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) h->n_sms = n;
    }
    h->stage = nullptr;
    h->last_stream = nullptr;
    for (int g = 0; g < MG_MAX_GROUPS; ++g) h->last_obs[g] = nullptr;
    LaunchParams &B = h->base;
    memset(&B, 0, sizeof B);
    B.n_groups = L->n_groups;
    B.T = L->series_len;
    B.Tp = (int32_t)Tp;
    B.cfg = L->cfg;
    B.load_raw = L->load_raw; B.pv_raw = L->pv_raw; B.grid_raw = L->grid_raw;
    B.load_nrm = L->load_nrm; B.pv_nrm = L->pv_nrm; B.grid_nrm = L->grid_nrm;
    B.plist = L->plist;
    int tiles = 0;
    for (int gidx = 0; gidx < L->n_groups; ++gidx) {
        const MgGroup &g = L->groups[gidx];
        DevGroup &d = B.g[gidx];
        if (g.n_envs < 1 || g.n_envs >= (1ll << 31) - MG_TILE) { delete h; return fail(MG_E_INVALID, "mg_create: group n_envs out of range"); }
        if (g.horizon < 0 || g.horizon > L->max_horizon) { delete h; return fail(MG_E_INVALID, "mg_create: group horizon exceeds max_horizon"); }
        const int rows = 1 + g.horizon;
        const int n_act = 1 + (g.has_grid != 0) + 2 * (g.has_genset != 0);
        const int obs_dim = rows * (2 + 4 * (g.has_grid != 0)) + 2 + 4 * (g.has_genset != 0);
        if (g.n_act != n_act || g.obs_dim != obs_dim) { delete h; return fail(MG_E_INVALID, "mg_create: n_act / obs_dim do not match the architecture"); }
        if (!g.step || !g.charge || !g.cfg_index || (g.has_genset && !g.genset)) { delete h; return fail(MG_E_INVALID, "mg_create: null state pointer"); }
        if (g.has_grid && L->n_grid < 1) { delete h; return fail(MG_E_INVALID, "mg_create: grid group without grid series"); }
        if (g.obs_order < MG_OBS_GYM_SORTED || g.obs_order > MG_OBS_GYM_SORTED_PV_FIRST) { delete h; return fail(MG_E_INVALID, "mg_create: bad obs_order"); }
        d.has_genset = g.has_genset != 0; d.has_grid = g.has_grid != 0; d.horizon = g.horizon;
        d.n_act = n_act; d.obs_dim = obs_dim; d.n_envs = (int32_t)g.n_envs;
        d.act_col_genset = g.act_col_genset; d.act_col_battery = g.act_col_battery; d.act_col_grid = g.act_col_grid;
        d.tile_begin = tiles;
        tiles += (int)((g.n_envs + MG_TILE - 1) / MG_TILE);
        layout_segments(g, d);
        if (obs_dim > MG_MAX_IMG) d.tma_ok = 0;
        d.long_path = (obs_dim > MG_MAX_IMG) || (d.state_start & 1);
        d.img_ok = rows <= 32;
        for (int q = 0; q < d.n_seg; ++q)
            if (d.seg_kind[q] == KIND_GRID && (d.seg_start[q] & 1)) d.img_ok = 0;
        if (d.long_path && g.grid_status_bits) { delete h; return fail(MG_E_UNSUPPORTED, "mg_create: per-env grid status needs the staged row path (obs_dim <= 192, even forecast rows)"); }
        d.step = g.step; d.charge = g.charge; d.genset = g.genset; d.cfg_index = g.cfg_index;
        d.env_initial = g.env_initial_step; d.env_final = g.env_final_step;
        d.status_bits = g.grid_status_bits; d.status_words = g.status_words;
        if (g.grid_status_bits) h->hetero = true;
        if (g.grid_status_bits && g.status_words * 32 < Tp) { delete h; return fail(MG_E_INVALID, "mg_create: grid_status_bits rows are shorter than T + max_horizon + 1 bits"); }
    }
    B.total_tiles = tiles;
    // normalised tables + bounds
    TableParams tp;
    tp.T = L->series_len; tp.Tp = (int32_t)Tp; tp.n_load = L->n_load; tp.n_pv = L->n_pv; tp.n_grid = L->n_grid;
    tp.load_raw = L->load_raw; tp.pv_raw = L->pv_raw; tp.grid_raw = L->grid_raw;
    tp.load_nrm = L->load_nrm; tp.pv_nrm = L->pv_nrm; tp.grid_nrm = L->grid_nrm; tp.bounds = L->bounds;
    const int cols = L->n_load + L->n_pv + 4 * L->n_grid;
    mg_build_tables_kernel<<<cols, 256, 0, (cudaStream_t)stream>>>(tp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "mg_create: table kernel launch"); }
    h->launches += 1;
    *out = h;
    return MG_OK;
}

extern "C" int mg_destroy(MgHandle *h) {
    if (h && h->stage) {
        stage_free(h->stage);
        cudaGetLastError();     // a context that is already gone at interpreter exit is not an error worth keeping
    }
    delete h;
    return MG_OK;
}

extern "C" int64_t mg_launch_count(const MgHandle *h) { return h ? h->launches : 0; }
extern "C" const char *mg_last_kernel(const MgHandle *h) { return h ? h->last_kernel : ""; }

extern "C" int mg_set_reported_soc(MgHandle *h, const double *const *soc) {
    if (!h) return fail(MG_E_INVALID, "mg_set_reported_soc: null handle");
    for (int g = 0; g < MG_MAX_GROUPS; ++g) h->soc_reported[g] = (soc && g < h->base.n_groups) ? soc[g] : nullptr;
    return MG_OK;
}

extern "C" int mg_set_trajectories(MgHandle *h, const int32_t *const *initial_step, const int32_t *const *final_step) {
    if (!h) return fail(MG_E_INVALID, "mg_set_trajectories: null handle");
    for (int g = 0; g < h->base.n_groups; ++g) {
        const int32_t *lo = initial_step ? initial_step[g] : nullptr, *hi = final_step ? final_step[g] : nullptr;
        h->base.g[g].env_initial = lo;
        h->base.g[g].env_final = hi;
        h->layout.groups[g].env_initial_step = lo;
        h->layout.groups[g].env_final_step = hi;
    }
    return MG_OK;
}

static int launch_noise(MgHandle *h, const MgForecastNoise *noise, void *const *obs, const int64_t *env_base, uint64_t seed,
                        uint64_t call, const int32_t *const *step_base, int32_t step_add, void *stream);

extern "C" int mg_forecast_noise(MgHandle *h, const MgForecastNoise *noise, void *const *obs, const int64_t *env_base,
                                 uint64_t seed, uint64_t call, void *stream) {
    return launch_noise(h, noise, obs, env_base, seed, call, nullptr, 0, stream);
}

extern "C" int mg_forecast_noise_at(MgHandle *h, const MgForecastNoise *noise, void *const *obs, const int64_t *env_base,
                                    uint64_t seed, uint64_t call, const int32_t *const *step_base, int32_t step_add, void *stream) {
    if (!step_base) return fail(MG_E_INVALID, "mg_forecast_noise_at: null step_base");
    return launch_noise(h, noise, obs, env_base, seed, call, step_base, step_add, stream);
}

static int launch_noise(MgHandle *h, const MgForecastNoise *noise, void *const *obs, const int64_t *env_base, uint64_t seed,
                        uint64_t call, const int32_t *const *step_base, int32_t step_add, void *stream) {
    if (!h || !noise || !obs) return fail(MG_E_INVALID, "mg_forecast_noise: null argument");
    NoiseParams P;
    memset(&P, 0, sizeof P);
    P.step_add = step_add;
    P.n_groups = h->base.n_groups;
    P.T = h->base.T;
    P.k0 = (uint32_t)seed;
    P.k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(call >> 32);
    P.c3 = (uint32_t)call;
    P.noise = noise;
    int blocks = 0;
    int64_t base = 0;
    for (int g = 0; g < P.n_groups; ++g) {
        const DevGroup &d = h->base.g[g];
        NoiseGroup &n = P.g[g];
        n.n_envs = obs[g] ? d.n_envs : 0;
        n.obs_dim = d.obs_dim; n.horizon = d.horizon; n.has_grid = d.has_grid;
        if (d.horizon * (2 + 4 * d.has_grid) > 2 * 65535) return fail(MG_E_UNSUPPORTED, "mg_forecast_noise: horizon too long");
        for (int q = 0; q < d.n_seg; ++q) {
            if (d.seg_kind[q] == KIND_LOAD) n.load_start = d.seg_start[q];
            else if (d.seg_kind[q] == KIND_PV) n.pv_start = d.seg_start[q];
            else if (d.seg_kind[q] == KIND_GRID) n.grid_start = d.seg_start[q];
        }
        n.first_block = blocks;
        n.env_base = env_base ? env_base[g] : base;
        n.step = d.step; n.cfg_index = d.cfg_index; n.obs = obs[g];
        n.step_base = step_base ? step_base[g] : nullptr;
        if (step_base && !n.step_base) n.n_envs = 0;
        blocks += (n.n_envs + MG_NOISE_ROWS - 1) / MG_NOISE_ROWS;
        base += d.n_envs;
    }
    if (blocks == 0) return MG_OK;
    if (h->obs_f32) mg_forecast_noise_kernel<float><<<blocks, 128, 0, (cudaStream_t)stream>>>(P);
    else mg_forecast_noise_kernel<double><<<blocks, 128, 0, (cudaStream_t)stream>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "mg_forecast_noise: launch");
    h->launches += 1;
    h->last_was_step = false;
    return MG_OK;
}


extern "C" int mg_set_option(MgHandle *h, int option, int value) {
    if (!h) return fail(MG_E_INVALID, "mg_set_option: null handle");
    if (option == MG_OPT_ROLLOUT_SPECIALISED) {
        h->rollout_specialised = value != 0;
        return MG_OK;
    }
    if (option == MG_OPT_ROLLOUT_RING) {
        h->rollout_ring = value != 0;
        return MG_OK;
    }
    if (option == MG_OPT_EMIT_IMAGE) {
        if (value < 0 || value > 2) return fail(MG_E_INVALID, "mg_set_option: MG_OPT_EMIT_IMAGE takes 0, 1 or 2");
        h->emit_image = value;
        return MG_OK;
    }
    if (option == MG_OPT_IMAGE_SHAPE) {
        if (value < -1 || value >= MG_N_IMAGE_SHAPES) return fail(MG_E_INVALID, "mg_set_option: image shape out of range");
        h->image_shape = value;
        return MG_OK;
    }
    if (option == MG_OPT_RAGGED_HINT) {
        h->ragged_hint = value != 0;
        return MG_OK;
    }
    if (option == MG_OPT_ACTIONS_F32) {
        h->actions_f32 = value != 0;
        return MG_OK;
    }
    if (option == MG_OPT_STEP_OVERLAP) {
        if (value < 0 || value > 2) return fail(MG_E_INVALID, "mg_set_option: MG_OPT_STEP_OVERLAP takes 0, 1 or 2");
        h->step_overlap = value;
        return MG_OK;
    }
    return fail(MG_E_INVALID, "mg_set_option: unknown option");
}

// ---- image-emitter kernels: (rows per bulk store, buffers per warp) shapes instantiated for mg_set_option(MG_OPT_IMAGE_SHAPE)
typedef void (*ImgKernel)(const LaunchParams);
static void img_kernel_for(bool hetero, bool ring, bool ws, int shape, ImgKernel *k, int *rc, int *nb) {
    // (rows per bulk store, image buffers per emitting warp, rows gathered together)
    static const int shapes[MG_N_IMAGE_SHAPES][2] = {{4, 2}, {2, 2}, {4, 2}, {4, 4}, {4, 4}, {8, 2}};
    *rc = shapes[shape][0];
    *nb = shapes[shape][1];
#define MG_IMG_PICK(RC, NB, GB, MINB)                                                                            \
    do {                                                                                                         \
        if (ring) *k = ws ? mg_rollout_img_kernel<RC, NB, GB, true, true, true, MINB> : mg_rollout_img_kernel<RC, NB, GB, true, true, false>;   \
        else if (hetero) *k = ws ? mg_rollout_img_kernel<RC, NB, GB, true, false, true> : mg_rollout_img_kernel<RC, NB, GB, true, false, false>; \
        else *k = ws ? mg_rollout_img_kernel<RC, NB, GB, false, false, true> : mg_rollout_img_kernel<RC, NB, GB, false, false, false>;    \
    } while (0)
    switch (shape) {
        case 1: MG_IMG_PICK(2, 2, 2, 0); break;
        case 2: MG_IMG_PICK(4, 2, 4, 0); break;
        // shapes 3-5, ring kernels: three CTAs per SM instead of four -- 168 registers (the per-env parameter record stays in
        // registers) and 38 KB of images per CTA; 31.7 us/step against 36.7 for shape 0 at 131 072 MicrogridGenerator grids
        case 3: MG_IMG_PICK(4, 4, 2, 3); break;
        case 4: MG_IMG_PICK(4, 4, 4, 3); break;
        case 5: MG_IMG_PICK(8, 2, 4, 3); break;
        default: MG_IMG_PICK(4, 2, 2, 0); break;
    }
#undef MG_IMG_PICK
}

static void img_step_kernel_for(bool hetero, int shape, ImgKernel *k, int *rc, int *nb) {
    ImgKernel unused;
    img_kernel_for(false, false, false, shape, &unused, rc, nb);
    switch (shape) {
        case 1: *k = hetero ? mg_step_img_kernel<2, 2, 2, true> : mg_step_img_kernel<2, 2, 2, false>; break;
        case 2: *k = hetero ? mg_step_img_kernel<4, 2, 4, true> : mg_step_img_kernel<4, 2, 4, false>; break;
        case 3: case 4: case 5: *k = hetero ? mg_step_img_kernel<4, 4, 2, true> : mg_step_img_kernel<4, 4, 2, false>; break;
        default: *k = hetero ? mg_step_img_kernel<4, 2, 2, true> : mg_step_img_kernel<4, 2, 2, false>; break;
    }
}

// opt a kernel into `bytes` of dynamic shared memory on the current device (once per kernel, device and size)
static int ensure_dynamic_smem(const void *func, size_t bytes) {
    struct Entry { const void *func; int device; size_t bytes; };
    static Entry seen[64];
    static int n_seen = 0;
    static std::mutex lock;
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    std::lock_guard<std::mutex> guard(lock);
    for (int i = 0; i < n_seen; ++i)
        if (seen[i].func == func && seen[i].device == device) {
            if (seen[i].bytes >= bytes) return MG_OK;
            e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute (dynamic shared memory)");
            seen[i].bytes = bytes;
            return MG_OK;
        }
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute (dynamic shared memory)");
    if (n_seen < 64) seen[n_seen++] = Entry{func, device, bytes};
    return MG_OK;
}

// Which emitters a launch uses when the caller left the choice to the library (MG_OPT_EMIT_IMAGE = 2), from what was
// measured on B200 (profiles/r02_emitters.md): the image + TMA bulk-store emitters wherever rows share no window with their
// neighbours (per-env series: 31.8 vs 76 us/step; envs at unrelated steps, MG_OPT_RAGGED_HINT: 15.3 vs 34-40), on batches of
// at most two tiles per SM (2.7-4.1 vs 6.8) and, for persistent launches, on rows with an even forecast horizon (14.1 vs
// 15.2; deeper image queues there and for per-env series); the per-lane store emitters with their run detection for large
// table-backed batches in lock-step -- level with the image kernels in tools/tune_emitters.py (11.44 both), 0.5 us/step
// ahead inside bench.py's longer launches, and 3-4 us/step ahead for single steps (14.6 vs 18.9 with step overlap).  Single
// steps use two-row bulk stores (shape 1: 21.2 vs 23.0 us/step at unrelated steps, 78.5 vs 106.8 for per-env series).
struct EmitChoice {
    bool image, split;
    int shape;
};
static EmitChoice choose_emitters(const MgHandle *h, const LaunchParams &P, bool persistent) {
    EmitChoice c;
    c.split = h->rollout_specialised;
    bool odd_rows = false;
    for (int g = 0; g < P.n_groups; ++g)
        if (P.g[g].has_grid && (P.g[g].horizon & 1) == 0) odd_rows = true;
    c.shape = h->image_shape >= 0 ? h->image_shape : persistent ? ((h->hetero || odd_rows) ? 4 : 0) : 1;
    if (h->emit_image != 2) c.image = h->emit_image == 1;
    else c.image = h->hetero || h->ragged_hint || P.total_tiles <= 2 * h->n_sms || (persistent && odd_rows);
    return c;
}

static int launch_step(MgHandle *h, const MgStepIO *io, int mode, int normalized, void *stream) {
    if (!h || !io) return fail(MG_E_INVALID, "step: null argument");
    LaunchParams P = h->base;
    P.mode = mode;
    P.normalized = normalized;
    P.pdl = h->step_overlap;
    for (int g = 0; g < P.n_groups; ++g) {
        DevGroup &d = P.g[g];
        d.actions = io[g].actions; d.dactions = io[g].dactions; d.obs = io[g].obs; d.reward = io[g].reward;
        d.done = io[g].done; d.info = io[g].info; d.flags = io[g].flags; d.mask = io[g].mask;
        d.reward_total = io[g].reward_total;
        d.log_slot = nullptr; d.log = nullptr;
        d.soc_reported = (mode == MODE_OBSERVE || mode == MODE_RESET) ? h->soc_reported[g] : nullptr;
        if (mode == MODE_STEP && !d.actions) return fail(MG_E_INVALID, "mg_step: null actions");
        if (mode == MODE_DISCRETE && (!d.dactions || !P.plist)) return fail(MG_E_INVALID, "mg_step_discrete: null actions or priority lists");
        if ((mode == MODE_STEP || mode == MODE_DISCRETE) && (!d.reward || !d.done)) return fail(MG_E_INVALID, "step: null reward / done");
        if ((mode == MODE_OBSERVE) && !d.obs) return fail(MG_E_INVALID, "mg_observe: null obs");
        if (d.obs && (((uintptr_t)d.obs) & 15)) return fail(MG_E_INVALID, "step: obs must be 16-byte aligned");
        d.act_f32 = h->actions_f32;
        if (d.actions && (d.n_act == 2 || d.n_act == 4) && (((uintptr_t)d.actions) & (d.act_f32 ? 4 * d.n_act - 1 : 15)))
            return fail(MG_E_INVALID, "step: actions must be 16-byte aligned (float32 rows of two: 8-byte)");
    }
    // Overlap with the previous step launch (PDL) only when that launch cannot still be writing the observation
    // buffers this one writes: the previous kernel releases its dependents before it streams its rows.
    // Chain this launch to the previous one (it may start once every CTA of that launch has passed its trigger) when both
    // are single-step launches on the same stream; in mode 2 the row streams overlap as well, so the observation buffers
    // must differ from those of the two launches before this one.
    bool overlap = h->step_overlap != 0 && h->last_was_step && h->last_stream == stream;
    for (int g = 0; g < P.n_groups && overlap && h->step_overlap == 2; ++g)
        for (int q = 0; q < P.n_groups; ++q)
            if (P.g[g].obs && (P.g[g].obs == h->last_obs[q] || P.g[g].obs == h->prev_obs[q])) overlap = false;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(P.total_tiles);
    cfg.blockDim = dim3(MG_THREADS);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1 : 0;
    // image emitter: every group that writes observations has an image-compatible f64 row layout
    const EmitChoice choice = choose_emitters(h, P, false);
    bool image = choice.image && !h->obs_f32, any_obs = false;
    int max_dim = 0;
    for (int g = 0; g < P.n_groups; ++g) {
        if (!P.g[g].obs) continue;
        any_obs = true;
        if (!P.g[g].img_ok) image = false;
        if (P.g[g].obs_dim > max_dim) max_dim = P.g[g].obs_dim;
    }
    cudaError_t e;
    if (image && any_obs) {
        ImgKernel k;
        int rc, nb;
        img_step_kernel_for(h->hetero, choice.shape, &k, &rc, &nb);
        cfg.dynamicSmemBytes = (size_t)MG_WARPS * nb * rc * max_dim * sizeof(double);
        const int rcode = ensure_dynamic_smem((const void *)k, cfg.dynamicSmemBytes);
        if (rcode != MG_OK) return rcode;
        e = cudaLaunchKernelEx(&cfg, k, P);
        h->last_kernel = "mg_step_img_kernel";
    } else if (h->obs_f32) e = h->hetero ? cudaLaunchKernelEx(&cfg, mg_step_kernel<true, float>, P) : cudaLaunchKernelEx(&cfg, mg_step_kernel<false, float>, P);
    else e = h->hetero ? cudaLaunchKernelEx(&cfg, mg_step_kernel<true, double>, P) : cudaLaunchKernelEx(&cfg, mg_step_kernel<false, double>, P);
    if (e != cudaSuccess) return cuda_fail(e, "step kernel launch");
    if (!(image && any_obs)) h->last_kernel = "mg_step_kernel";
    if (mode == MODE_STEP || mode == MODE_DISCRETE) memset(h->soc_reported, 0, sizeof h->soc_reported);   // every battery updates
    h->launches += 1;
    h->last_was_step = P.pdl != 0;
    h->last_stream = stream;
    for (int g = 0; g < MG_MAX_GROUPS; ++g) {
        h->prev_obs[g] = h->last_obs[g];
        h->last_obs[g] = g < P.n_groups ? P.g[g].obs : nullptr;
    }
    return MG_OK;
}

extern "C" int mg_step(MgHandle *h, const MgStepIO *io, int normalized, void *stream) {
    return launch_step(h, io, MODE_STEP, normalized != 0, stream);
}
extern "C" int mg_step_discrete(MgHandle *h, const MgStepIO *io, void *stream) {
    return launch_step(h, io, MODE_DISCRETE, 0, stream);
}
extern "C" int mg_reset(MgHandle *h, const MgStepIO *io, void *stream) { return launch_step(h, io, MODE_RESET, 0, stream); }
extern "C" int mg_observe(MgHandle *h, const MgStepIO *io, void *stream) { return launch_step(h, io, MODE_OBSERVE, 0, stream); }

static int launch_rollout(MgHandle *h, const MgRolloutIO *io, int32_t n_steps, int32_t ring, int mode, int normalized,
                          void *stream) {
    if (!h || !io) return fail(MG_E_INVALID, "rollout: null argument");
    if (n_steps < 1 || ring < 1) return fail(MG_E_INVALID, "rollout: n_steps and ring must be >= 1");
    LaunchParams P = h->base;
    P.mode = mode;
    P.normalized = normalized;
    P.n_steps = n_steps;
    P.ring = ring;
    for (int g = 0; g < P.n_groups; ++g) {
        DevGroup &d = P.g[g];
        d.actions = io[g].actions; d.dactions = io[g].dactions; d.obs = io[g].obs_ring; d.reward = io[g].reward;
        d.done = io[g].done; d.reward_sum = io[g].reward_sum; d.flags = io[g].flags; d.info = nullptr; d.mask = nullptr;
        d.reward_total = io[g].reward_total;
        d.log_slot = io[g].log_slot; d.log = io[g].log;
        if ((d.log_slot != nullptr) != (d.log != nullptr)) return fail(MG_E_INVALID, "rollout: log_slot and log go together");
        d.act_step_stride = (int64_t)d.n_envs * d.n_act;
        d.out_step_stride = d.n_envs;
        d.dact_step_stride = io[g].dactions_const ? 0 : d.n_envs;
        d.obs_slot_stride = (int64_t)d.n_envs * d.obs_dim;
        if (mode == MODE_STEP && !d.actions) return fail(MG_E_INVALID, "mg_rollout: null actions");
        if (mode == MODE_DISCRETE && (!d.dactions || !P.plist)) return fail(MG_E_INVALID, "mg_rollout_discrete: null actions or priority lists");
        if (!d.reward || !d.done) return fail(MG_E_INVALID, "rollout: null reward / done");
        if (d.obs && (((uintptr_t)d.obs) & 15)) return fail(MG_E_INVALID, "rollout: obs_ring must be 16-byte aligned");
        d.act_f32 = h->actions_f32;
        if (d.actions && (d.n_act == 2 || d.n_act == 4)) {
            const uintptr_t mask = d.act_f32 ? 4 * d.n_act - 1 : 15;
            if ((((uintptr_t)d.actions) & mask) || ((d.act_step_stride * (d.act_f32 ? 4 : 8)) & mask))
                return fail(MG_E_INVALID, "rollout: action rows must stay 16-byte aligned across steps (float32 rows of two: 8-byte)");
        }
    }
    // per-env series rows are expensive to assemble: four emitting warps beat two there (98 vs 120 us/step measured)
    bool ws = MG_ROLLOUT_WS != 0 && !h->hetero && h->rollout_specialised;
    for (int g = 0; g < P.n_groups; ++g)
        if (!P.g[g].obs || P.g[g].long_path) ws = false;
    // per-env series: keep the sliding windows in shared memory when every group's ring fits (H <= 24) and rows are staged
    bool use_ring = h->hetero && h->rollout_ring;
    for (int g = 0; g < P.n_groups; ++g)
        if (!P.g[g].obs || P.g[g].horizon + 2 > MG_RING_MAX) use_ring = false;
    // image emitter (rows leave as TMA bulk stores): every group writes f64 observations with an image-compatible layout
    const EmitChoice choice = choose_emitters(h, P, true);
    bool image = choice.image && !h->obs_f32;
    int max_dim = 0;
    for (int g = 0; g < P.n_groups; ++g) {
        if (!P.g[g].obs || !P.g[g].img_ok) image = false;
        if (P.g[g].obs_dim > max_dim) max_dim = P.g[g].obs_dim;
    }
    bool any_log = false;
    for (int g = 0; g < P.n_groups; ++g)
        if (P.g[g].log) any_log = true;
    if (any_log) image = ws = false;          // the logging kernels are instantiations of mg_rollout_kernel
    if (!image)
        for (int g = 0; g < P.n_groups; ++g)
            if (P.g[g].long_path) use_ring = false;
    if (image) {
        const bool ws_img = choice.split;
        ImgKernel k;
        int rc, nb;
        img_kernel_for(h->hetero, use_ring, ws_img, choice.shape, &k, &rc, &nb);
        const size_t dyn = (size_t)(ws_img ? 2 : MG_WARPS) * nb * rc * max_dim * sizeof(double);
        const int rcode = ensure_dynamic_smem((const void *)k, dyn);
        if (rcode != MG_OK) return rcode;
        k<<<P.total_tiles, MG_THREADS, dyn, (cudaStream_t)stream>>>(P);
        h->last_kernel = ws_img ? "mg_rollout_img_kernel (owner / emitter warps)" : "mg_rollout_img_kernel";
    } else if (ws) {
        h->last_kernel = "mg_rollout_ws_kernel";
        if (h->obs_f32) mg_rollout_ws_kernel<false, float><<<P.total_tiles, MG_THREADS, 0, (cudaStream_t)stream>>>(P);
        else mg_rollout_ws_kernel<false, double><<<P.total_tiles, MG_THREADS, 0, (cudaStream_t)stream>>>(P);
    } else if (any_log) {
        h->last_kernel = "mg_rollout_kernel (log)";
        const dim3 grid(P.total_tiles), block(MG_THREADS);
        cudaStream_t st = (cudaStream_t)stream;
        if (h->obs_f32) {
            if (use_ring) mg_rollout_kernel<true, float, true, true><<<grid, block, 0, st>>>(P);
            else if (h->hetero) mg_rollout_kernel<true, float, false, true><<<grid, block, 0, st>>>(P);
            else mg_rollout_kernel<false, float, false, true><<<grid, block, 0, st>>>(P);
        } else {
            if (use_ring) mg_rollout_kernel<true, double, true, true><<<grid, block, 0, st>>>(P);
            else if (h->hetero) mg_rollout_kernel<true, double, false, true><<<grid, block, 0, st>>>(P);
            else mg_rollout_kernel<false, double, false, true><<<grid, block, 0, st>>>(P);
        }
    } else if (use_ring) {
        h->last_kernel = "mg_rollout_kernel (rings)";
        if (h->obs_f32) mg_rollout_kernel<true, float, true><<<P.total_tiles, MG_THREADS, 0, (cudaStream_t)stream>>>(P);
        else mg_rollout_kernel<true, double, true><<<P.total_tiles, MG_THREADS, 0, (cudaStream_t)stream>>>(P);
    } else if (h->obs_f32) {
        h->last_kernel = "mg_rollout_kernel";
        if (h->hetero) mg_rollout_kernel<true, float><<<P.total_tiles, MG_THREADS, 0, (cudaStream_t)stream>>>(P);
        else mg_rollout_kernel<false, float><<<P.total_tiles, MG_THREADS, 0, (cudaStream_t)stream>>>(P);
    } else {
        h->last_kernel = "mg_rollout_kernel";
        if (h->hetero) mg_rollout_kernel<true, double><<<P.total_tiles, MG_THREADS, 0, (cudaStream_t)stream>>>(P);
        else mg_rollout_kernel<false, double><<<P.total_tiles, MG_THREADS, 0, (cudaStream_t)stream>>>(P);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "rollout kernel launch");
    memset(h->soc_reported, 0, sizeof h->soc_reported);
    h->launches += 1;
    h->last_was_step = false;
    return MG_OK;
}

extern "C" int mg_rollout(MgHandle *h, const MgRolloutIO *io, int32_t n_steps, int32_t ring, int normalized, void *stream) {
    return launch_rollout(h, io, n_steps, ring, MODE_STEP, normalized != 0, stream);
}
extern "C" int mg_rollout_discrete(MgHandle *h, const MgRolloutIO *io, int32_t n_steps, int32_t ring, void *stream) {
    return launch_rollout(h, io, n_steps, ring, MODE_DISCRETE, 0, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// mg_rollout_host: host-resident actions / results, chunked and pipelined over three streams
// ------------------------------------------------------------------------------------------------------------------
#define MG_CUDA_TRY(call, what)                              \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) return cuda_fail(e_, what);   \
    } while (0)

static inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

extern "C" int mg_rollout_host(MgHandle *h, const MgHostRolloutIO *io, int32_t n_steps, int32_t chunk, int32_t ring,
                               int discrete, int normalized, void *stream) {
    if (!h || !io) return fail(MG_E_INVALID, "mg_rollout_host: null argument");
    if (n_steps < 1 || chunk < 1 || ring < 1) return fail(MG_E_INVALID, "mg_rollout_host: n_steps, chunk and ring must be >= 1");
    if (chunk > n_steps) chunk = n_steps;
    const int G = h->base.n_groups;
    size_t in_off[MG_MAX_GROUPS], rew_off[MG_MAX_GROUPS], done_off[MG_MAX_GROUPS], act_row[MG_MAX_GROUPS];
    size_t in_bytes = 0, out_bytes = 0;
    for (int g = 0; g < G; ++g) {
        const DevGroup &d = h->base.g[g];
        if (discrete ? !io[g].dactions : !io[g].actions) return fail(MG_E_INVALID, "mg_rollout_host: null host actions");
        if (!io[g].reward || !io[g].done) return fail(MG_E_INVALID, "mg_rollout_host: null host reward / done");
        act_row[g] = discrete ? (size_t)d.n_envs * sizeof(int32_t)
                              : (size_t)d.n_envs * d.n_act * (h->actions_f32 ? sizeof(float) : sizeof(double));   // bytes per step
        in_off[g] = in_bytes;
        in_bytes += align256(act_row[g] * chunk);
        rew_off[g] = out_bytes;
        out_bytes += align256((size_t)d.n_envs * sizeof(double) * chunk);
        done_off[g] = out_bytes;
        out_bytes += align256((size_t)d.n_envs * chunk);
    }
    cudaStream_t cur = (cudaStream_t)stream;
    if (!h->stage) {
        HostStage *st = new (std::nothrow) HostStage();
        if (!st) return fail(MG_E_INVALID, "mg_rollout_host: out of host memory");
        h->stage = st;      // (released by mg_destroy, also when the set-up below stops half way)
        MG_CUDA_TRY(cudaStreamCreateWithFlags(&st->s_in, cudaStreamNonBlocking), "mg_rollout_host: stream");
        MG_CUDA_TRY(cudaStreamCreateWithFlags(&st->s_out, cudaStreamNonBlocking), "mg_rollout_host: stream");
        for (int k = 0; k < 2; ++k) {
            MG_CUDA_TRY(cudaEventCreateWithFlags(&st->uploaded[k], cudaEventDisableTiming), "mg_rollout_host: event");
            MG_CUDA_TRY(cudaEventCreateWithFlags(&st->computed[k], cudaEventDisableTiming), "mg_rollout_host: event");
            MG_CUDA_TRY(cudaEventCreateWithFlags(&st->drained[k], cudaEventDisableTiming), "mg_rollout_host: event");
        }
        MG_CUDA_TRY(cudaEventCreateWithFlags(&st->fence, cudaEventDisableTiming), "mg_rollout_host: event");
    }
    HostStage *st = h->stage;
    if (in_bytes > st->in_cap || out_bytes > st->out_cap) {     // (re)allocate: nothing may still be using the old slots
        MG_CUDA_TRY(cudaStreamSynchronize(st->s_in), "mg_rollout_host: sync");
        MG_CUDA_TRY(cudaStreamSynchronize(st->s_out), "mg_rollout_host: sync");
        if (st->fence_recorded) MG_CUDA_TRY(cudaEventSynchronize(st->fence), "mg_rollout_host: sync");
        for (int k = 0; k < 2; ++k) {
            if (st->in[k]) cudaFree(st->in[k]);
            if (st->out[k]) cudaFree(st->out[k]);
            st->in[k] = st->out[k] = nullptr;
        }
        st->in_cap = st->out_cap = 0;
        for (int k = 0; k < 2; ++k) {
            MG_CUDA_TRY(cudaMalloc((void **)&st->in[k], in_bytes), "mg_rollout_host: cudaMalloc (action staging)");
            MG_CUDA_TRY(cudaMalloc((void **)&st->out[k], out_bytes), "mg_rollout_host: cudaMalloc (result staging)");
        }
        st->in_cap = in_bytes;
        st->out_cap = out_bytes;
    }
    // Order this call after the previous one (whatever stream that ran on) and the copy streams after the caller's
    // earlier work on `stream` (state loads, resets).
    if (st->fence_recorded) MG_CUDA_TRY(cudaStreamWaitEvent(cur, st->fence, 0), "mg_rollout_host: wait");
    MG_CUDA_TRY(cudaEventRecord(st->fence, cur), "mg_rollout_host: record");
    MG_CUDA_TRY(cudaStreamWaitEvent(st->s_in, st->fence, 0), "mg_rollout_host: wait");
    MG_CUDA_TRY(cudaStreamWaitEvent(st->s_out, st->fence, 0), "mg_rollout_host: wait");
    const int mode = discrete ? MODE_DISCRETE : MODE_STEP;
    MgRolloutIO rio[MG_MAX_GROUPS];
    int c = 0;
    for (int32_t s0 = 0; s0 < n_steps; s0 += chunk, ++c) {
        const int slot = c & 1;
        const int32_t n = n_steps - s0 < chunk ? n_steps - s0 : chunk;
        // copy-in: this slot's previous reader (chunk c-2) must have finished
        if (c >= 2) MG_CUDA_TRY(cudaStreamWaitEvent(st->s_in, st->computed[slot], 0), "mg_rollout_host: wait");
        for (int g = 0; g < G; ++g) {
            const char *src = discrete ? (const char *)io[g].dactions : (const char *)io[g].actions;
            MG_CUDA_TRY(cudaMemcpyAsync(st->in[slot] + in_off[g], src + (size_t)s0 * act_row[g], (size_t)n * act_row[g],
                                        cudaMemcpyHostToDevice, st->s_in), "mg_rollout_host: copy-in");
        }
        MG_CUDA_TRY(cudaEventRecord(st->uploaded[slot], st->s_in), "mg_rollout_host: record");
        // compute: needs the actions, and the slot's previous results (chunk c-2) must have left the device
        MG_CUDA_TRY(cudaStreamWaitEvent(cur, st->uploaded[slot], 0), "mg_rollout_host: wait");
        if (c >= 2) MG_CUDA_TRY(cudaStreamWaitEvent(cur, st->drained[slot], 0), "mg_rollout_host: wait");
        memset(rio, 0, sizeof rio);
        for (int g = 0; g < G; ++g) {
            if (discrete) rio[g].dactions = (const int32_t *)(st->in[slot] + in_off[g]);
            else rio[g].actions = (const double *)(st->in[slot] + in_off[g]);
            rio[g].obs_ring = (double *)io[g].obs_ring;
            rio[g].reward = (double *)(st->out[slot] + rew_off[g]);
            rio[g].done = (uint8_t *)(st->out[slot] + done_off[g]);
            rio[g].flags = io[g].flags;
        }
        const int rc = launch_rollout(h, rio, n, ring, mode, normalized != 0, stream);
        if (rc != MG_OK) return rc;
        MG_CUDA_TRY(cudaEventRecord(st->computed[slot], cur), "mg_rollout_host: record");
        // copy-out
        MG_CUDA_TRY(cudaStreamWaitEvent(st->s_out, st->computed[slot], 0), "mg_rollout_host: wait");
        for (int g = 0; g < G; ++g) {
            const size_t n_envs = (size_t)h->base.g[g].n_envs;
            MG_CUDA_TRY(cudaMemcpyAsync(io[g].reward + (size_t)s0 * n_envs, st->out[slot] + rew_off[g], (size_t)n * n_envs * sizeof(double),
                                        cudaMemcpyDeviceToHost, st->s_out), "mg_rollout_host: copy-out");
            MG_CUDA_TRY(cudaMemcpyAsync(io[g].done + (size_t)s0 * n_envs, st->out[slot] + done_off[g], (size_t)n * n_envs,
                                        cudaMemcpyDeviceToHost, st->s_out), "mg_rollout_host: copy-out");
        }
        MG_CUDA_TRY(cudaEventRecord(st->drained[slot], st->s_out), "mg_rollout_host: record");
    }
    // the caller's stream completes only when the last results are in host memory
    MG_CUDA_TRY(cudaStreamWaitEvent(cur, st->drained[(c - 1) & 1], 0), "mg_rollout_host: wait");
    if (c >= 2) MG_CUDA_TRY(cudaStreamWaitEvent(cur, st->drained[c & 1], 0), "mg_rollout_host: wait");
    MG_CUDA_TRY(cudaEventRecord(st->fence, cur), "mg_rollout_host: record");
    st->fence_recorded = true;
    return MG_OK;
}
