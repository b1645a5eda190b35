"""CPU suite: the device headers (compiled for the host by tests/hostsim — a TEST TOOL, not a product path) against
the golden fixtures of the reference.  Same checks the -m gpu suite runs through the C ABI on the B200."""
import ctypes as C

import numpy as np
import pytest

import common
from hostsim import hostsim
from rlgymppo_cpp_b200 import abi


def _runner(team, car_preset=0, mutate=None):
    cfg = abi.default_cfg(num_arenas=1, team_size=team)
    cfg.car_preset = car_preset
    if mutate is not None:
        mutate(cfg)
    hs = hostsim.HostSim(cfg)
    return (lambda c, b, p, t: hs.set_state(0, c, b, p, t)), (lambda u: hs.tick(0, u, 1)), (lambda: hs.get_state(0))


def test_action_table(golden_dir):
    import os

    assert np.array_equal(hostsim.action_table(), np.load(os.path.join(golden_dir, "action_table.npy")))


def test_batched_tick_protocol_equals_the_serial_one():
    """common.check_single_tick_batch (what the -m gpu suite uses: every recorded tick in its own arena, one launch) on the host
    build, arena by arena: same tallies as the one-arena protocol."""
    cfg = abi.default_cfg(num_arenas=1, team_size=2)
    hs = hostsim.HostSim(cfg)

    def run(cars, balls, pads, ticks, ctl):
        out = []
        for i in range(len(ticks)):
            hs.set_state(0, cars[i], balls[i:i + 1], pads[i], int(ticks[i]))
            hs.tick(0, ctl[i], 1)
            out.append(hs.get_state(0))
        c = np.zeros((len(out), hs.P), dtype=abi.CAR_DTYPE)
        b = np.zeros(len(out), dtype=abi.BALL_DTYPE)
        p = np.zeros((len(out), abi.RLG_NUM_PADS), dtype=abi.PAD_DTYPE)
        for i, o in enumerate(out):
            c[i], b[i], p[i] = o[0], o[1][0], o[2]
        return c, b, p, np.array([o[3] for o in out], dtype=np.int64)

    g = common.load_tick_file("tick_random_2v2")
    a = common.check_single_tick_batch(g, run)
    s, t, gs = _runner(2)
    b = common.check_single_tick_run(g, s, t, gs)
    assert a["total"] == b["total"] == 384 and a["loose"] == b["loose"] and a["worst_tight"] == b["worst_tight"]


def test_epa_small_workspace_fallback():
    """The penetration-depth search first runs in a small workspace and is repeated in the reference-size one when that runs
    out (rl_epa.h with_epa_ws): with the small one shrunk until most evaluations overflow, the results are unchanged."""
    g = {k: v for k, v in common.load_tick_file("tick_scenarios_1v1").items() if k in ("car_into_goal", "car_up_side_ramp")}
    L = hostsim.lib()
    L.hs_epa_overflows.restype = C.c_long
    s, t, gs = _runner(1)
    before = L.hs_epa_overflows()
    a = common.check_single_tick_run(g, s, t, gs)
    assert L.hs_epa_overflows() == before  # the production capacity fits all of them
    try:
        L.hs_epa_small_capacity(2, 6)
        s, t, gs = _runner(1)
        b = common.check_single_tick_run(g, s, t, gs)
        assert L.hs_epa_overflows() >= before + 3
    finally:
        L.hs_epa_small_capacity(10, 24)
    assert a["worst_tight"] == b["worst_tight"] and a["loose"] == b["loose"]


def test_single_tick_scenarios_1v1():
    s, t, g = _runner(1)
    res = common.check_single_tick_run(common.load_tick_file("tick_scenarios_1v1"), s, t, g)
    print(res)


@pytest.mark.parametrize("team", [1, 2, 3])
def test_single_tick_random_play(team):
    s, t, g = _runner(team)
    res = common.check_single_tick_run(common.load_tick_file(f"tick_random_{team}v{team}"), s, t, g)
    print(res)


@pytest.mark.parametrize("preset,name", list(common.CAR_PRESETS))
def test_single_tick_random_play_car_presets(preset, name):
    """The five non-Octane CarConfigs (CarConfig.cpp:20-88): hitbox, wheel and suspension geometry."""
    s, t, g = _runner(1, preset)
    res = common.check_single_tick_run(common.load_tick_file(f"tick_random_1v1_{name}"), s, t, g)
    print(res)


@pytest.mark.parametrize("team", [1, 2])
def test_single_tick_random_play_mutators(team):
    """A MutatorConfig that differs from the soccar default in every honoured field (MutatorConfig.h:16-72; gravity with x/y
    components, drag, frictions / restitutions, jump and boost accelerations, pad cooldowns, demo-on-contact with team demos,
    unlimited flips): random play recorded from the reference under common.apply_test_mutators."""
    s, t, g = _runner(team, mutate=common.apply_test_mutators)
    res = common.check_single_tick_run(common.load_tick_file(f"tick_random_{team}v{team}_mutators"), s, t, g)
    print(res)


def test_single_tick_ball_mass_radius_mutators():
    """MutatorConfig::ballMass / ballRadius (45 units, 100 uu) and a carMass that the reference's Gym never applies (Gym.cpp:40-49 sets the
    mutators before the cars exist; Car.cpp:206-209 builds cars with RLConst::CAR_MASS_BT): random play + the scripted scenarios."""
    s, t, g = _runner(1, mutate=common.apply_ball_mutators)
    print(common.check_single_tick_run(common.load_tick_file("tick_random_1v1_ballmut"), s, t, g))
    s, t, g = _runner(1, mutate=common.apply_ball_mutators)
    print(common.check_single_tick_run(common.load_tick_file("tick_scenarios_1v1_ballmut"), s, t, g))


def test_single_tick_scenarios_mutators():
    """The scripted scenarios under the test mutators: demolition on contact at bump speed, 1 s respawn with 60 boost, team-mate
    demolition (2v2), pad cooldowns of 1.5 / 3 s, ball hits with the extra-impulse scale, repeated flips in one jump."""
    s, t, g = _runner(1, mutate=common.apply_test_mutators)
    print(common.check_single_tick_run(common.load_tick_file("tick_scenarios_1v1_mutators"), s, t, g))
    s, t, g = _runner(2, mutate=common.apply_test_mutators)
    print(common.check_single_tick_run(common.load_tick_file("tick_scenarios_2v2_mutators"), s, t, g))


def test_mutators_default_is_identity():
    """mutators_set = 1 with rlg_mutators_default values is the same engine as mutators_set = 0."""
    import ctypes as C

    from rlgymppo_cpp_b200 import engine

    m = abi.Mutators()
    engine.load_library().rlg_mutators_default(C.byref(m))
    d = abi.default_mutators()
    assert bytes(m) == bytes(d)


@pytest.mark.parametrize("field,value", [("car_mass", -1.0), ("ball_mass", 0.0), ("ball_radius", 120.0), ("demo_mode", 7)])
def test_mutators_unsupported_rejected(field, value):
    """Values the reference itself refuses (a ball wider than a broadphase cell: btRSBroadphase.cpp:229-230) or cannot simulate."""
    cfg = abi.default_cfg(num_arenas=1, team_size=1)
    cfg.mutators_set = 1
    setattr(cfg.mutators, field, type(getattr(cfg.mutators, field))(value))
    with pytest.raises(Exception):
        hostsim.HostSim(cfg)


@pytest.mark.parametrize("name,cfg", list(common.gym_cfgs()))
def test_gym_layer_bit_exact(name, cfg, golden_dir):
    import os

    g = np.load(os.path.join(golden_dir, name + ".npz"))
    hs = hostsim.HostSim(cfg)
    hs.set_player_order(g["player_order"])
    for i in range(len(g["tick"])):
        hs.set_state(0, g["cars"][i], g["ball"][i:i + 1], g["pads"][i], int(g["tick"][i]))
        if g["first"][i]:
            obs = hs.reset_from_current(0)
            assert common.obs_equal(cfg, g["obs"][i], obs), (name, i)
        else:
            obs, r, d = hs.eval_gym(g["actions"][i], 0)
            assert common.obs_equal(cfg, g["obs"][i], obs), (name, i)
            assert common.rewards_equal(name, g["reward"][i], r), (name, i, r, g["reward"][i])
            assert d == bool(g["done"][i]), (name, i)


def test_state_setters_statistics():
    """RandomState / KickoffState use the engine's own RNG: check structure and ranges (RandomState.cpp:8-62, Arena.cpp:112-216)."""
    cfg = abi.default_cfg(num_arenas=64, team_size=2)
    hs = hostsim.HostSim(cfg)
    for a in range(64):
        hs.reset(a)
        cars, ball, pads, tick = hs.get_state(a)
        assert np.all(np.abs(ball["pos"][0][:2]) <= [3500, 4000]) and 92.7 <= ball["pos"][0][2] <= 1820
        assert np.linalg.norm(ball["vel"][0]) <= 4000.01
        assert np.allclose(cars["pos"][:, 2], 17) and np.all(cars["is_on_ground"] == 1)
        assert np.all((cars["boost"] >= 0) & (cars["boost"] <= 100))
        assert np.all(pads["is_active"] == 1)
    cfg.state_setter = abi.RLG_SETTER_KICKOFF
    hs = hostsim.HostSim(cfg)
    spots = {(-2048, -2560), (2048, -2560), (-256, -3840), (256, -3840), (0, -4608)}
    for a in range(16):
        hs.reset(a)
        cars, ball, pads, tick = hs.get_state(a)
        assert np.allclose(ball["pos"][0], [0, 0, 93.15], atol=1e-4) and np.all(ball["vel"] == 0)
        for c in cars:
            x, y = float(c["pos"][0]), float(c["pos"][1])
            if c["team"] == 1:
                x, y = -x, -y
            assert (round(x), round(y)) in spots
            assert abs(c["boost"] - 100 / 3) < 1e-4


def _ref_reset_samples(cfg, n, seed):
    from oracle import refsim

    refsim.seed(seed)
    g = refsim.RefGym(cfg)
    out = []
    for _ in range(n):
        g.reset()
        c, b, p, _t = g.arena.get_state()
        out.append((c, b, p))
    return out


def test_random_state_distribution_matches_the_reference():
    """RandomState(true, true, true) (RandomState.cpp:8-62) on the host build vs the live reference: every marginal the setter
    defines passes a two-sample KS test (the RNG engines differ by construction, so only distributions can agree)."""
    from oracle import refsim

    if not refsim.available():
        pytest.skip("oracle/_ref not built")
    cfg = abi.default_cfg(num_arenas=1, team_size=1)
    ref = _ref_reset_samples(cfg, 3000, 3)
    hs = hostsim.HostSim(cfg)
    ours = []
    for _ in range(3000):
        hs.reset(0)
        c, b, p, _t = hs.get_state(0)
        ours.append((c, b, p))
    print(common.compare_setter_samples(ours, ref))


def test_leaf_grid_returns_the_bvh_walks_leaves_in_order():
    """The per-cell leaf lists (rl_mesh.h grid_lookup) must give exactly the stackless walks' candidate leaves, in the
    walks' order, for every query box the grid accepts; bigger boxes must fall back to the walk."""
    import ctypes as C

    h = hostsim.HostSim(abi.default_cfg(num_arenas=1, team_size=1))
    out = (C.c_int64 * 3)()
    h.L.hs_grid_check(h.h, 200000, C.c_uint64(7), C.c_float(0.3), C.c_float(2.6), out)
    answered, found, mismatches = out[0], out[1], out[2]
    assert mismatches == 0
    assert answered == 200000 and found > 10000  # every small box is answered by the grid, and plenty reach leaves
    h.L.hs_grid_check(h.h, 20000, C.c_uint64(8), C.c_float(2.0), C.c_float(6.0), out)
    assert out[2] == 0 and 0 < out[0] < 20000  # mixed: some too big for the grid -> walk, same answer
