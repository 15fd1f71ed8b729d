"""ctypes mirror of include/pymgrid_b200.h -- the only place Python touches the C-ABI.

The library must exist (built in-tree by pymgrid_b200.build / __graft_entry__.build()); there is NO CPU or
PyTorch fallback: a missing or mismatching extension raises at import of the engine.
"""
import ctypes as C
import os

from . import build as _build

MG_ABI_VERSION = 2
MG_MAX_GROUPS = 8
MG_N_INFO = 16
MG_N_LOG = 24
MG_LOG_STEP, MG_LOG_CHARGE, MG_LOG_GENSET_BEFORE, MG_LOG_GENSET_AFTER, MG_LOG_REWARD, MG_LOG_DONE, MG_LOG_FLAGS, MG_LOG_INFO = 0, 1, 2, 3, 4, 5, 6, 8
MG_PLIST_WIDTH = 3
MG_OBS_GYM_SORTED, MG_OBS_CONTAINER, MG_OBS_GYM_SORTED_PV_FIRST = 0, 1, 2
MG_MOD_NONE, MG_MOD_GENSET, MG_MOD_BATTERY, MG_MOD_GRID = -1, 0, 1, 2
MG_OPT_ROLLOUT_SPECIALISED = 1
MG_OPT_ROLLOUT_RING = 2
MG_OPT_EMIT_IMAGE = 3
MG_OPT_IMAGE_SHAPE = 4
MG_OPT_RAGGED_HINT = 5
MG_OPT_STEP_OVERLAP = 6
MG_OPT_ACTIONS_F32 = 7

FLAG_NAMES = {
    1 << 0: "GENSET_GOAL_RANGE", 1 << 1: "GENSET_AS_SINK", 1 << 2: "BALANCE", 1 << 3: "BATTERY_MIN_CAP",
    1 << 4: "NEGATIVE_ABSORB", 1 << 5: "STEP_PAST_END", 1 << 6: "BAD_ACTION", 1 << 7: "SHAPER_RANGE",
    1 << 8: "CLIP_GENSET", 1 << 9: "CLIP_BATTERY", 1 << 10: "CLIP_GRID",
    1 << 12: "BATTERY_SINK", 1 << 13: "GRID_SINK", 1 << 14: "EXCESS",
}
FLAG_BATTERY_SINK, FLAG_GRID_SINK, FLAG_EXCESS = 1 << 12, 1 << 13, 1 << 14
FLAG_CLIP_MASK = 0x700
FLAG_ERROR_MASK = 0xff      # the reference raises at these; the CLIP_* bits only raise under raise_errors=True

INFO_NAMES = ("load_met", "pv_used", "curtailment", "loss_load", "overgeneration", "genset_production",
              "genset_co2", "battery_discharge", "battery_charge", "grid_import", "grid_export", "grid_co2",
              "reward_genset", "reward_battery", "reward_grid", "reward_unbalanced")

_d, _i32, _vp = C.c_double, C.c_int32, C.c_void_p


class MgConfig(C.Structure):
    _fields_ = [(n, _d) for n in (
        "bat_min_capacity", "bat_max_capacity", "bat_max_charge", "bat_max_discharge", "bat_efficiency",
        "bat_cost_cycle", "bat_act_low", "bat_act_spread", "bat_soc_low", "bat_soc_spread", "bat_charge_spread",
        "gen_running_min", "gen_running_max", "gen_cost", "gen_co2_per_unit", "gen_cost_per_unit_co2",
        "gen_act_spread", "gen_up_spread", "gen_down_spread",
        "grid_max_import", "grid_max_export", "grid_cost_per_unit_co2", "grid_act_low", "grid_act_spread",
        "loss_load_cost", "overgeneration_cost",
        "load_scale", "pv_scale", "load_low", "load_spread", "pv_low", "pv_spread", "load_fill_nrm", "pv_fill_nrm")] + [(n, _i32) for n in (
        "gen_start_up_time", "gen_wind_down_time", "gen_allow_abortion", "load_series", "pv_series", "grid_series",
        "initial_step", "final_step", "plist_offset", "plist_count", "series_scaled", "grid_status_weak", "reward_shaper")] + [("reserved", _i32 * 3)]


class MgPriorityList(C.Structure):
    _fields_ = [("module", C.c_int8 * MG_PLIST_WIDTH), ("action", C.c_int8 * MG_PLIST_WIDTH),
                ("n_elements", C.c_int8), ("_pad", C.c_int8)]


class MgGroup(C.Structure):
    _fields_ = [("has_genset", _i32), ("has_grid", _i32), ("horizon", _i32), ("obs_order", _i32),
                ("n_act", _i32), ("obs_dim", _i32), ("n_envs", C.c_int64),
                ("act_col_genset", _i32), ("act_col_battery", _i32), ("act_col_grid", _i32), ("_pad", _i32),
                ("step", _vp), ("charge", _vp), ("genset", _vp), ("cfg_index", _vp),
                ("env_initial_step", _vp), ("env_final_step", _vp), ("grid_status_bits", _vp), ("status_words", C.c_int64)]


class MgLayout(C.Structure):
    _fields_ = [("abi_version", _i32), ("n_groups", _i32), ("groups", MgGroup * MG_MAX_GROUPS),
                ("n_cfg", _i32), ("series_len", _i32), ("max_horizon", _i32),
                ("n_load", _i32), ("n_pv", _i32), ("n_grid", _i32),
                ("cfg", _vp), ("load_raw", _vp), ("pv_raw", _vp), ("grid_raw", _vp),
                ("load_nrm", _vp), ("pv_nrm", _vp), ("grid_nrm", _vp), ("bounds", _vp),
                ("plist", _vp), ("n_plist", _i32), ("flags", _i32)]


class MgStepIO(C.Structure):
    _fields_ = [("actions", _vp), ("dactions", _vp), ("obs", _vp), ("reward", _vp), ("done", _vp), ("info", _vp),
                ("flags", _vp), ("mask", _vp), ("reward_total", _vp)]


class MgRolloutIO(C.Structure):
    _fields_ = [("actions", _vp), ("dactions", _vp), ("obs_ring", _vp), ("reward", _vp), ("done", _vp),
                ("reward_sum", _vp), ("flags", _vp), ("dactions_const", C.c_int64), ("reward_total", _vp),
                ("log_slot", _vp), ("log", _vp)]


class MgHostRolloutIO(C.Structure):
    _fields_ = [("actions", _vp), ("dactions", _vp), ("reward", _vp), ("done", _vp), ("obs_ring", _vp), ("flags", _vp)]


class MgForecastNoise(C.Structure):
    _fields_ = [("load_sigma", _d), ("pv_sigma", _d), ("grid_sigma", _d * 4),
                ("load_increase", _i32), ("pv_increase", _i32), ("grid_increase", _i32), ("_pad", _i32)]


class EngineError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libpymgrid_b200.so (building it if the source is newer and nvcc is present) and verify the ABI."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path) or (_build.is_stale() and os.path.exists(_build.SRC)):
        try:
            path = _build.build()
        except Exception as exc:   # no nvcc on this machine and no prebuilt library: fail loudly
            if not os.path.exists(path):
                raise EngineError(f"CUDA extension {path} is missing and could not be built: {exc}") from exc
    L = C.CDLL(path)
    L.mg_abi_version.restype = C.c_int
    L.mg_sizeof.restype = C.c_int64
    L.mg_sizeof.argtypes = [C.c_int]
    L.mg_build_info.restype = C.c_char_p
    L.mg_last_error.restype = C.c_char_p
    L.mg_create.argtypes = [C.POINTER(MgLayout), _vp, C.POINTER(_vp)]
    L.mg_destroy.argtypes = [_vp]
    L.mg_step.argtypes = [_vp, C.POINTER(MgStepIO), C.c_int, _vp]
    L.mg_step_discrete.argtypes = [_vp, C.POINTER(MgStepIO), _vp]
    L.mg_reset.argtypes = [_vp, C.POINTER(MgStepIO), _vp]
    L.mg_observe.argtypes = [_vp, C.POINTER(MgStepIO), _vp]
    L.mg_rollout.argtypes = [_vp, C.POINTER(MgRolloutIO), _i32, _i32, C.c_int, _vp]
    L.mg_rollout_discrete.argtypes = [_vp, C.POINTER(MgRolloutIO), _i32, _i32, _vp]
    L.mg_rollout_host.argtypes = [_vp, C.POINTER(MgHostRolloutIO), _i32, _i32, _i32, C.c_int, C.c_int, _vp]
    L.mg_set_option.argtypes = [_vp, C.c_int, C.c_int]
    L.mg_forecast_noise.argtypes = [_vp, _vp, C.POINTER(_vp), C.POINTER(C.c_int64), C.c_uint64, C.c_uint64, _vp]
    L.mg_forecast_noise_at.argtypes = [_vp, _vp, C.POINTER(_vp), C.POINTER(C.c_int64), C.c_uint64, C.c_uint64, C.POINTER(_vp), C.c_int32, _vp]
    L.mg_forecast_noise_at.restype = C.c_int
    L.mg_set_reported_soc.argtypes = [_vp, C.POINTER(_vp)]
    L.mg_launch_count.argtypes = [_vp]
    L.mg_launch_count.restype = C.c_int64
    L.mg_set_trajectories.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp)]
    L.mg_set_trajectories.restype = C.c_int
    L.mg_last_kernel.argtypes = [_vp]
    L.mg_last_kernel.restype = C.c_char_p
    if L.mg_abi_version() != MG_ABI_VERSION:
        raise EngineError(f"ABI mismatch: library {L.mg_abi_version()} vs binding {MG_ABI_VERSION}")
    for which, struct in enumerate((MgConfig, MgPriorityList, MgGroup, MgLayout, MgStepIO, MgRolloutIO, MgForecastNoise,
                                    MgHostRolloutIO)):
        if L.mg_sizeof(which) != C.sizeof(struct):
            raise EngineError(f"struct {struct.__name__}: library sizeof {L.mg_sizeof(which)} != binding {C.sizeof(struct)}")
    _lib = L
    return L


EXPORTED_SYMBOLS = ("mg_abi_version", "mg_sizeof", "mg_build_info", "mg_last_error", "mg_create", "mg_destroy",
                    "mg_step", "mg_step_discrete", "mg_reset", "mg_observe", "mg_rollout", "mg_rollout_discrete",
                    "mg_rollout_host", "mg_launch_count", "mg_last_kernel", "mg_set_trajectories", "mg_set_option", "mg_forecast_noise", "mg_forecast_noise_at", "mg_set_reported_soc")


def check(code, what):
    if code != 0:
        raise EngineError(f"{what} failed ({code}): {lib().mg_last_error().decode()}")
