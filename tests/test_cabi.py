"""CPU-side checks of the C-ABI: the library loads without a GPU, exports every entry point include/pymgrid_b200.h
declares, and the ctypes struct mirrors have the compiled sizes.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "pymgrid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("mg_create", "mg_destroy", "mg_step", "mg_step_discrete", "mg_reset", "mg_observe", "mg_rollout",
                 "mg_rollout_discrete", "mg_last_error", "mg_abi_version", "mg_sizeof", "mg_launch_count", "mg_last_kernel", "mg_set_trajectories", "mg_set_option", "mg_forecast_noise", "mg_forecast_noise_at"):
        assert must in names


def test_library_loads_and_exports_every_declared_symbol():
    from pymgrid_b200 import _cabi
    L = _cabi.lib()                       # verifies ABI version and struct sizes, raises EngineError otherwise
    for name in declared_functions():
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert set(declared_functions()) == set(_cabi.EXPORTED_SYMBOLS)
    assert L.mg_abi_version() == _cabi.MG_ABI_VERSION
    assert b"sm_100a" in L.mg_build_info()
    assert L.mg_sizeof(0) == 336 == C.sizeof(_cabi.MgConfig)
    assert L.mg_sizeof(99) == -1


def test_compose_header_symbols_are_exported_and_bound():
    """include/pymgrid_b200_compose.h: every declared entry point is exported by the CUDA library, the ctypes mirrors
    have the compiled sizes, invalid layouts are refused on the host (no compute calls)."""
    from pymgrid_b200 import _cabi, compose
    text = open(os.path.join(ROOT, "include", "pymgrid_b200_compose.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mgc_[a-z_0-9]+)\s*\(", text)))
    assert set(declared) == set(compose.EXPORTED_SYMBOLS)
    L = compose.bind(_cabi.lib())         # ABI version, struct sizes and parameter-block widths, raises otherwise
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    h = C.c_void_p()
    assert L.mgc_create(None, C.byref(h)) == -1 and b"null" in L.mg_last_error()
    layout = compose.MgcLayout()
    layout.abi_version = 99
    assert L.mgc_create(C.byref(layout), C.byref(h)) == -1 and b"abi_version" in L.mg_last_error()
    assert L.mgc_run(None, None, 1, 1, 1, None) == -1
    assert L.mgc_sizeof(99) == -1 and L.mgc_param_count(99) == -1


def test_invalid_arguments_are_rejected_without_touching_the_gpu():
    from pymgrid_b200 import _cabi
    L = _cabi.lib()
    h = C.c_void_p()
    assert L.mg_create(None, None, C.byref(h)) == -1 and b"null" in L.mg_last_error()
    layout = _cabi.MgLayout()
    layout.abi_version = 99
    assert L.mg_create(C.byref(layout), None, C.byref(h)) == -1 and b"abi_version" in L.mg_last_error()
    layout.abi_version = _cabi.MG_ABI_VERSION
    layout.n_groups = 0
    assert L.mg_create(C.byref(layout), None, C.byref(h)) == -1
    assert L.mg_step(None, None, 1, None) == -1


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    from pymgrid_b200.engine import BatchedMicrogrid, EngineError
    from pymgrid_b200.scenario import load_pymgrid25
    with pytest.raises(EngineError):
        BatchedMicrogrid([load_pymgrid25(0)], np.zeros(4, dtype=np.int64))


def test_priority_tables_match_the_reference_action_lists(golden):
    """Host-side action tables == DiscreteMicrogridEnv.actions_list recorded from the reference."""
    from pymgrid_b200.priority_list import priority_lists
    from pymgrid_b200.scenario import load_pymgrid25
    z = golden["discrete"]
    for n in range(25):
        p = load_pymgrid25(n)
        table = priority_lists(p.has_genset, p.has_grid, p.genset.running_min_production if p.genset else None)
        mod, act = z[f"h23_s{n}_table_mod"], z[f"h23_s{n}_table_act"]
        assert len(table) == len(mod) == {(1, 0): 4, (0, 1): 2, (1, 1): 12}[(int(p.has_genset), int(p.has_grid))]
        for row, pl in enumerate(table):
            assert list(pl) == [(int(m), int(a)) for m, a in zip(mod[row], act[row]) if m >= 0]
    assert len(priority_lists(True, True, 0.0)) == 6          # redundant genset-off lists removed (priority_list.py:53-67)
    assert len(priority_lists(True, True, 0.0, remove_redundant_gensets=False)) == 12


def test_rbc_priority_list_matches_reference(golden):
    """RuleBasedControl's automatic priority list (algos/rbc/rbc.py:31-44) for the recorded scenarios."""
    from pymgrid_b200.priority_list import priority_lists, rbc_priority_list
    from pymgrid_b200.scenario import load_pymgrid25
    z = golden["rbc"]
    for n in (0, 1, 2, 5, 9, 13):
        p = load_pymgrid25(n)
        pl = rbc_priority_list(p)
        assert [m for m, _ in pl] == list(z[f"s{n}_list_mod"]) and [a for _, a in pl] == list(z[f"s{n}_list_act"])
        assert pl in priority_lists(p.has_genset, p.has_grid, p.genset.running_min_production if p.genset else None)


def test_oracle_rbc_rollout_matches_reference(golden):
    """The oracle driven with the RBC list reproduces the reference controller's rewards (incl. the full year of
    scenario 0: sum -956 059.6622849072)."""
    import numpy as np
    from oracle.oracle import OracleGrid
    from pymgrid_b200.priority_list import rbc_priority_list
    from pymgrid_b200.scenario import load_pymgrid25
    z = golden["rbc"]
    for n in (0, 1, 2, 5, 9, 13):
        p = load_pymgrid25(n)
        o = OracleGrid(p)
        pl = rbc_priority_list(p)
        want = z[f"s{n}_rewards"]
        got = np.empty(len(want))
        for k in range(len(want)):
            _, got[k], _, _, _ = o.run(o.priority_control(list(pl)), normalized=False)
        np.testing.assert_array_equal(got, want)
    assert float(np.add.reduce(z["s0_rewards"])) == -956059.6622849072 or abs(z["s0_rewards"].sum() + 956059.6622849072) < 1e-6
