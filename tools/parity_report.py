"""Per-fixture single-tick parity report: total ticks, ticks outside TOL_TIGHT, worst errors inside / outside it, and the list
of the non-tight ticks.  `python tools/parity_report.py host|gpu [out.json]` — host = tests/hostsim build of the device headers,
gpu = the engine through the C ABI (what the -m gpu tests run)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from rlgymppo_cpp_b200 import abi  # noqa: E402

FIXTURES = [("tick_scenarios_1v1", 1, 0, False), ("tick_random_1v1", 1, 0, False), ("tick_random_2v2", 2, 0, False), ("tick_random_3v3", 3, 0, False)] + \
    [(f"tick_random_1v1_{n}", 1, p, False) for p, n in common.CAR_PRESETS] + \
    [("tick_random_1v1_mutators", 1, 0, True), ("tick_random_2v2_mutators", 2, 0, True), ("tick_scenarios_1v1_mutators", 1, 0, True),
     ("tick_scenarios_2v2_mutators", 2, 0, True)]


def make_runner(kind, team, preset, mut):
    cfg = abi.default_cfg(num_arenas=1 if kind == "host" else 32, team_size=team)
    cfg.car_preset = preset
    if mut:
        common.apply_test_mutators(cfg)
    if kind == "host":
        from hostsim import hostsim
        hs = hostsim.HostSim(cfg)
        return (lambda c, b, p, t: hs.set_state(0, c, b, p, t)), (lambda u: hs.tick(0, u, 1)), (lambda: hs.get_state(0))
    import test_gpu_parity as tg
    return tg._gpu_runner(cfg)


def report(kind, names=None):
    out = {}
    for name, team, preset, mut in FIXTURES:
        if names and name not in names:
            continue
        s, t, g = make_runner(kind, team, preset, mut)
        res = common.check_single_tick_run(common.load_tick_file(name), s, t, g, allow_contact_frac=1.0, detail=True)
        out[name] = res
        print(name, {k: v for k, v in res.items() if k != "loose_ticks"}, flush=True)
        for lt in res["loose_ticks"]:
            print("    ", lt)
    return out


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "host"
    out = report(kind, set(sys.argv[3:]) or None)
    if len(sys.argv) > 2 and sys.argv[2] != "-":
        with open(sys.argv[2], "w") as f:
            json.dump(out, f, indent=1)
