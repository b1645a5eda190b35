"""CPU suite: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from rlgymppo_cpp_b200 import abi, build, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return engine.load_library()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "rlgym_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rlg_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/rlgym_b200.h but not exported"
    for s in engine.EXPORTS:
        assert s in syms, f"binding uses {s} which the header does not declare"


def test_struct_sizes(lib):
    assert lib.rlg_sizeof_car_state() == C.sizeof(abi.CarState)
    assert lib.rlg_sizeof_engine_cfg() == C.sizeof(abi.EngineCfg)
    assert lib.rlg_abi_version() == 1


def test_default_cfg_matches_examplemain(lib):
    c = abi.EngineCfg()
    lib.rlg_engine_cfg_default(C.byref(c))
    d = abi.default_cfg()
    for f, _ in abi.EngineCfg._fields_:
        if f == "reward_terms":
            for i in range(c.num_reward_terms):
                assert c.reward_terms[i].kind == d.reward_terms[i].kind
                assert c.reward_terms[i].weight == d.reward_terms[i].weight
                assert list(c.reward_terms[i].params) == list(d.reward_terms[i].params)
        elif f == "mutators":  # MutatorConfig(SOCCAR) defaults, byte for byte
            assert bytes(c.mutators) == bytes(d.mutators)
        else:
            assert getattr(c, f) == getattr(d, f), f


def test_action_table_no_device_needed(lib, golden_dir):
    import numpy as np

    t = engine.action_table()
    assert t.shape == (90, 8)
    assert np.array_equal(t, np.load(os.path.join(golden_dir, "action_table.npy")))


def test_no_cpu_fallback(lib):
    """Without a CUDA device engine creation must fail loudly, never fall back to a CPU path."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.EngineError) as ei:
        engine.Engine(abi.default_cfg(num_arenas=4))
    assert "CUDA" in str(ei.value) or "cuda" in str(ei.value)


def test_invalid_config_errors(lib):
    import torch

    bad = abi.default_cfg(num_arenas=0)
    with pytest.raises(engine.EngineError):
        engine.Engine(bad)
    bad = abi.default_cfg(num_arenas=4, team_size=3)
    bad.obs_kind = abi.RLG_OBS_PADDED
    bad.obs_max_players = 2  # DefaultOBSPadded: too many teammates (DefaultOBSPadded.cpp:41-45)
    with pytest.raises(engine.EngineError):
        engine.Engine(bad)
