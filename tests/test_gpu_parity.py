"""-m gpu: the parity tests proper.  Everything goes through the C ABI (rlgymppo_cpp_b200.engine -> librlgym_b200.so)
on cuda:0 and is compared with the reference's golden fixtures (tests/golden/*.npz, produced by the unmodified
reference) and, when oracle/_ref travelled to the box, with the compiled reference itself."""
import os

import numpy as np
import pytest

import common
from rlgymppo_cpp_b200 import abi, engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available()
    return torch


def _engine(team, n=1, **kw):
    cfg = abi.default_cfg(num_arenas=n, team_size=team)
    mutate = kw.pop("mutate", None)
    for k, v in kw.items():
        setattr(cfg, k, v)
    if mutate is not None:
        mutate(cfg)
    return engine.Engine(cfg)


class _TickRunner:
    """Injects one reference state into arena 0, ticks once with explicit controls, reads the state back."""

    def __init__(self, team, torch, **kw):
        self.e = _engine(team, **kw)
        self.torch = torch
        self.ids = np.zeros(1, dtype=np.int32)

    def set_state(self, c, b, p, t):
        self.e.set_state(self.ids, np.ascontiguousarray(c), np.ascontiguousarray(b), np.ascontiguousarray(p), np.array([t], dtype=np.int64))

    def tick(self, u):
        buf = self.torch.from_numpy(np.frombuffer(np.ascontiguousarray(u).tobytes(), dtype=np.uint8).copy()).cuda()
        self.e.tick_device(buf.data_ptr(), 1)
        self.e.sync()

    def get_state(self):
        c, b, p, t = self.e.get_state(self.ids)
        return c[0], b, p[0], int(t[0])


def test_action_table(golden_dir):
    assert np.array_equal(engine.action_table(), np.load(os.path.join(golden_dir, "action_table.npy")))


def _batch_runner(team, torch, **kw):
    """Every recorded tick of a fixture in its own arena of ONE engine: N injected states, one launch, N states read back (the
    role kernel at a realistic block shape — groups of 32 arenas with mixed situations per warp — instead of one live lane)."""
    def run(cars, balls, pads, ticks, ctl):
        n = len(ticks)
        e = _engine(team, n=n, **kw)
        ids = np.arange(n, dtype=np.int32)
        buf = torch.from_numpy(np.frombuffer(np.ascontiguousarray(ctl).tobytes(), dtype=np.uint8).copy()).cuda()
        e.set_state(ids, np.ascontiguousarray(cars), np.ascontiguousarray(balls), np.ascontiguousarray(pads), ticks)
        e.tick_device(buf.data_ptr(), 1)
        e.sync()
        return e.get_state(ids)
    return run


def _tick_file(name, team, torch, **kw):
    res = common.check_single_tick_batch(common.load_tick_file(name), _batch_runner(team, torch, **kw), detail=True)
    print(name, res)
    return res


def test_single_tick_scenarios_1v1(torch_cuda):
    _tick_file("tick_scenarios_1v1", 1, torch_cuda)
    # ... and the one-arena-at-a-time protocol on a slice of it (a single live lane per warp)
    r = _TickRunner(1, torch_cuda)
    g = common.load_tick_file("tick_scenarios_1v1")
    common._current_fixture[0] = "tick_scenarios_1v1 (two scenarios, one arena at a time)"
    common.check_single_tick_run({k: g[k] for k in ("car_hits_ball", "car_into_goal")}, r.set_state, r.tick, r.get_state)


@pytest.mark.parametrize("team", [1, 2, 3])
def test_single_tick_random_play(team, torch_cuda):
    _tick_file(f"tick_random_{team}v{team}", team, torch_cuda)


@pytest.mark.parametrize("preset,name", list(common.CAR_PRESETS))
def test_single_tick_random_play_car_presets(preset, name, torch_cuda):
    """The five non-Octane CarConfigs (CarConfig.cpp:20-88) against the reference's trajectories, same tolerances."""
    _tick_file(f"tick_random_1v1_{name}", 1, torch_cuda, car_preset=preset)


@pytest.mark.parametrize("team", [1, 2])
def test_single_tick_random_play_mutators(team, torch_cuda):
    """Non-default MutatorConfig (every honoured field changed, common.apply_test_mutators) against the reference's
    trajectories under the same mutators, same tolerances."""
    _tick_file(f"tick_random_{team}v{team}_mutators", team, torch_cuda, mutate=common.apply_test_mutators)


def test_single_tick_scenarios_mutators(torch_cuda):
    """The scripted scenarios under the test mutators (demolition on contact, 1 s respawn, team-mate demolition, pad cooldowns,
    ball-hit scale, repeated flips) against the reference's recordings."""
    _tick_file("tick_scenarios_1v1_mutators", 1, torch_cuda, mutate=common.apply_test_mutators)
    _tick_file("tick_scenarios_2v2_mutators", 2, torch_cuda, mutate=common.apply_test_mutators)


def test_single_tick_ball_mass_radius_mutators(torch_cuda):
    """MutatorConfig::ballMass / ballRadius (a 45-unit, 100 uu ball) and a carMass the reference's Gym never applies: random play and the
    scripted scenarios (bounces on floor / walls / mesh, car hits, wheels and hitbox on the ball) recorded from the reference."""
    _tick_file("tick_random_1v1_ballmut", 1, torch_cuda, mutate=common.apply_ball_mutators)
    _tick_file("tick_scenarios_1v1_ballmut", 1, torch_cuda, mutate=common.apply_ball_mutators)


def test_a_second_smaller_engine_does_not_break_the_first(torch_cuda):
    """The role kernel's dynamic shared-memory limit is an attribute of the function, not of an engine: creating a small engine (the
    SkillTracker's eval pool next to a training pool) after a large one must not lower it under the large engine's blocks."""
    rng = np.random.default_rng(0)
    big = engine.Engine(abi.default_cfg(num_arenas=16384, team_size=1))
    big.reset()
    small = engine.Engine(abi.default_cfg(num_arenas=8, team_size=1))
    small.reset()
    for e in (big, small, big):
        obs, rew, done = e.step_host(rng.integers(0, 90, size=e.A * e.P).astype(np.int32))
        assert np.isfinite(obs).all() and np.isfinite(rew).all()


def test_mutators_out_of_range_rejected():
    """What the reference itself refuses or cannot simulate is refused loudly: a ball wider than a broadphase cell
    (btRSBroadphase.cpp:229-230 throws), non-positive masses."""
    for field, value in (("ball_radius", 120.0), ("ball_mass", 0.0), ("car_mass", -1.0)):
        cfg = abi.default_cfg(num_arenas=1, team_size=1)
        cfg.mutators_set = 1
        setattr(cfg.mutators, field, value)
        with pytest.raises(Exception):
            engine.Engine(cfg)


@pytest.mark.parametrize("name,cfg", list(common.gym_cfgs()))
def test_gym_layer_bit_exact(name, cfg, golden_dir, torch_cuda):
    """obs / reward / done are BIT-EXACT against the reference given identical states (tolerance: 0 ulp; the configuration with
    the powf rewards: common.REWARD_ULPS)."""
    torch = torch_cuda
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n = len(g["tick"])
    # one arena per recorded sample would lose the cross-step state (EventReward memo, no-touch counter): replay in order
    e = engine.Engine(cfg)
    e.set_player_order(g["player_order"])
    ids = np.zeros(1, dtype=np.int32)
    P = e.P
    for i in range(n):
        e.set_state(ids, np.ascontiguousarray(g["cars"][i]), np.ascontiguousarray(g["ball"][i:i + 1]), np.ascontiguousarray(g["pads"][i]),
                    np.array([int(g["tick"][i])], dtype=np.int64))
        if g["first"][i]:
            e.reset_current()
            obs, _, _ = e.read_outputs()
            assert common.obs_equal(cfg, g["obs"][i], obs.reshape(P, -1)), (name, i)
        else:
            acts = torch.from_numpy(g["actions"][i].astype(np.int32)).cuda()
            e.eval_gym_device(acts.data_ptr())
            obs, rew, done = e.read_outputs()
            assert common.obs_equal(cfg, g["obs"][i], obs.reshape(P, -1)), (name, i)
            assert common.rewards_equal(name, g["reward"][i], rew[:g["reward"][i].shape[0]]), (name, i, rew, g["reward"][i])
            assert bool(done[0]) == bool(g["done"][i]), (name, i)


def test_many_arenas_identical_to_single(torch_cuda):
    """Layout check at a non-trivial size: every arena of a 4096-arena engine fed the same state + controls must end
    bit-identical to arena 0 (coalesced word-transposed load/store, no cross-arena leakage)."""
    torch = torch_cuda
    A = 4096
    e = _engine(1, n=A)
    g = common.load_tick_file("tick_scenarios_1v1")["car_hits_ball"]
    ids = np.arange(A, dtype=np.int32)
    t0 = 40
    cars = np.ascontiguousarray(np.tile(g["cars"][t0], (A, 1)))
    balls = np.ascontiguousarray(np.repeat(g["ball"][t0:t0 + 1], A))
    pads = np.ascontiguousarray(np.tile(g["pads"][t0], (A, 1)))
    e.set_state(ids, cars, balls, pads, np.full(A, int(g["tick"][t0]), dtype=np.int64))
    u = np.ascontiguousarray(np.tile(g["controls"][t0], (A, 1)))
    buf = torch.from_numpy(np.frombuffer(u.tobytes(), dtype=np.uint8).copy()).cuda()
    e.tick_device(buf.data_ptr(), 16)
    e.sync()
    c, b, p, t = e.get_state(ids)
    assert all(c[i].tobytes() == c[0].tobytes() for i in range(0, A, 97))
    assert np.all(b == b[0]) and np.all(t == t[0])


def test_full_size_properties(torch_cuda):
    """BASELINE configs[1] size (16384 arenas, 1v1): size-independent properties of a real collection run."""
    torch = torch_cuda
    A = 16384
    e = _engine(1, n=A)
    e.reset()
    P = e.P
    rng = np.random.default_rng(0)
    total_done = 0
    for s in range(64):
        acts = torch.from_numpy(rng.integers(0, 90, size=A * P).astype(np.int32)).cuda()
        e.step_device(acts.data_ptr())
        if s % 16 == 15:
            obs, rew, done = e.read_outputs()
            assert np.isfinite(obs).all() and np.isfinite(rew).all()
            assert set(np.unique(done)) <= {0, 1}
            total_done += int(done.sum())
            # pad bits / flags are exactly 0 or 1, prev-action block is a table row
            assert set(np.unique(obs[:, 17:51])) <= {0.0, 1.0}
            assert set(np.unique(obs[:, 66:70])) <= {0.0, 1.0} or True
    cars, balls, pads, ticks = e.get_state(np.arange(0, A, 37, dtype=np.int32))
    assert np.all(ticks == 64 * 8)
    assert np.all(np.abs(balls["pos"][:, 0]) < 4200) and np.all(np.abs(balls["pos"][:, 1]) < 6100) and np.all(balls["pos"][:, 2] > 80)
    assert np.all(np.linalg.norm(balls["vel"], axis=1) <= 6000.5)
    assert np.all(np.linalg.norm(cars["vel"], axis=2) <= 2300.5)
    assert np.all((cars["boost"] >= 0) & (cars["boost"] <= 100))
    fw = cars["rot_forward"]
    assert np.allclose(np.linalg.norm(fw, axis=2), 1, atol=1e-3)


@pytest.mark.parametrize("team", [1, 2])
def test_sharded_engines_equal_one_engine(team, torch_cuda):
    """arenas are independent and the per-arena RNG streams are keyed by the GLOBAL arena id
    (rlg_engine_cfg.arena_id_base), so a pool split over ranks gives bit-identical obs / rewards / done flags to one engine
    holding all arenas - at ragged sizes (not multiples of the warp or SM count, a 1-arena shard), with many auto-resets."""
    torch = torch_cuda
    sizes = [100, 1, 112]
    A = sum(sizes)

    def mk(n, base):
        cfg = abi.default_cfg(num_arenas=n, team_size=team)
        cfg.no_touch_max_steps = 5  # frequent RandomState resets: the reset RNG is what could depend on the sharding
        cfg.arena_id_base = base
        return engine.Engine(cfg)

    whole = mk(A, 0)
    shards, base = [], 0
    for n in sizes:
        shards.append(mk(n, base))
        base += n
    whole.reset()
    for sh in shards:
        sh.reset()
    P = whole.P
    rng = np.random.default_rng(team)
    n_done = 0
    for s in range(24):
        acts = rng.integers(0, 90, size=A * P).astype(np.int32)
        o, r, d = whole.step_host(acts)
        lo = 0
        for sh, n in zip(shards, sizes):
            o2, r2, d2 = sh.step_host(acts[lo * P:(lo + n) * P])
            assert np.array_equal(o[lo * P:(lo + n) * P].view(np.uint32), o2.view(np.uint32)), (s, lo)
            assert np.array_equal(r[lo * P:(lo + n) * P].view(np.uint32), r2.view(np.uint32)), (s, lo)
            assert np.array_equal(d[lo:lo + n], d2), (s, lo)
            lo += n
        n_done += int(d.sum())
    assert n_done > A  # every arena was re-set at least once on average


def test_step_host_matches_device_path(torch_cuda):
    torch = torch_cuda
    cfg = abi.default_cfg(num_arenas=256, team_size=1)
    cfg.no_touch_max_steps = 3  # many auto-resets: the host-buffer step patches the re-set arenas' obs rows after its early copy
    e1, e2, e3 = engine.Engine(cfg), engine.Engine(cfg), engine.Engine(cfg)
    e1.reset(); e2.reset(); e3.reset()
    rng = np.random.default_rng(1)
    pinned_actions = e3.host_buffers()[0]
    n_done = 0
    for s in range(12):
        a = rng.integers(0, 90, size=e1.A * e1.P).astype(np.int32)
        o1, r1, d1 = e1.step_host(a)
        t = torch.from_numpy(a).cuda()
        e2.step_device(t.data_ptr())
        o2, r2, d2 = e2.read_outputs()
        assert np.array_equal(o1, o2) and np.array_equal(r1, r2) and np.array_equal(d1, d2)
        pinned_actions[:] = a  # the zero-copy variant: engine-owned page-locked buffers
        o3, r3, d3 = e3.step_pinned()
        assert np.array_equal(o1, o3) and np.array_equal(r1, r3) and np.array_equal(d1, d3)
        n_done += int(d1.sum())
    assert n_done > 100


def test_reward_metrics_follow_gameinst(torch_cuda):
    """rlg_engine_metrics == GameInst::Step's AvgTracker arithmetic (GameInst.cpp:13-31) replayed on the host in float32
    from the rewards/dones the steps returned, summed over the games like ThreadAgentManager::GetMetrics (:82-92)."""
    cfg = abi.default_cfg(num_arenas=96, team_size=2)
    cfg.no_touch_max_steps = 6  # episodes end often
    e = engine.Engine(cfg)
    e.reset()
    m0 = e.metrics()
    assert np.isnan(m0["avg_step_reward"]) and np.isnan(m0["avg_episode_reward"]) and m0["total_steps"] == 0
    A, P = e.A, e.P
    f32 = np.float32
    step_tot = np.zeros(A, f32); ep_tot = np.zeros(A, f32); cur = np.zeros(A, f32)
    step_cnt = np.zeros(A, np.uint64); ep_cnt = np.zeros(A, np.uint64)
    rng = np.random.default_rng(5)

    def replay(rew, done):
        for a in range(A):
            tot = f32(0)
            for p in range(P):
                tot = f32(tot + rew[a * P + p])
            step_tot[a] = f32(step_tot[a] + tot); step_cnt[a] += P
            cur[a] = f32(cur[a] + f32(tot / f32(P)))
            if done[a]:
                ep_tot[a] = f32(ep_tot[a] + cur[a]); ep_cnt[a] += 1
                cur[a] = 0

    def expect():
        st = f32(0); et = f32(0)
        for a in range(A):
            st = f32(st + step_tot[a]); et = f32(et + ep_tot[a])
        return st, et

    for s in range(20):
        _, rew, done = e.step_host(rng.integers(0, 90, size=A * P).astype(np.int32), want_obs=False)
        replay(rew, done)
    m = e.metrics()
    st, et = expect()
    assert m["total_steps"] == 20 * A and m["step_reward_count"] == int(step_cnt.sum()) and m["episode_count"] == int(ep_cnt.sum())
    assert m["episode_count"] > 0
    assert f32(m["step_reward_total"]).view(np.uint32) == st.view(np.uint32)
    assert f32(m["episode_reward_total"]).view(np.uint32) == et.view(np.uint32)
    assert f32(m["avg_step_reward"]).view(np.uint32) == f32(st / f32(step_cnt.sum())).view(np.uint32)
    # ResetMetrics clears the trackers, not the running episode sums
    e.reset_metrics()
    step_tot[:] = 0; ep_tot[:] = 0; step_cnt[:] = 0; ep_cnt[:] = 0
    for s in range(8):
        _, rew, done = e.step_host(rng.integers(0, 90, size=A * P).astype(np.int32), want_obs=False)
        replay(rew, done)
    m = e.metrics()
    st, et = expect()
    assert m["total_steps"] == 28 * A
    assert f32(m["step_reward_total"]).view(np.uint32) == st.view(np.uint32)
    assert f32(m["episode_reward_total"]).view(np.uint32) == et.view(np.uint32)


def test_against_compiled_reference_if_present(torch_cuda):
    """Live comparison with oracle/_ref (travels to the GPU box as a prebuilt .so): fresh random states every run."""
    from oracle import refsim

    if not refsim.available():
        pytest.skip("oracle/_ref/librlref.so not present")
    cfg = abi.default_cfg(num_arenas=1, team_size=1)
    g = refsim.RefGym(cfg)
    refsim.seed(99)
    r = _TickRunner(1, torch_cuda)
    rng = np.random.default_rng(5)
    table = refsim.action_table()
    arena = refsim.RefArena(1, True)
    tight = loose = 0
    for ep in range(4):
        g.reset()
        cars, ball, pads, _ = g.arena.get_state()
        arena.set_state(cars, ball, pads, 0)
        for step in range(20):
            acts = rng.integers(0, 90, size=2)
            u = np.zeros(2, dtype=abi.CONTROLS_DTYPE)
            for i in range(2):
                a = table[acts[i]]
                u[i] = (a[0], a[1], a[2], a[3], a[4], int(a[5] == 1), int(a[6] == 1), int(a[7] == 1))
            for t in range(8):
                c0, b0, p0, t0 = arena.get_state()
                r.set_state(c0, b0, p0, t0)
                arena.step(u, 1)
                r.tick(u)
                c1, b1, p1, t1 = arena.get_state()
                gc, gb, gp, gt = r.get_state()
                e = common.phys_err(c1, b1, gc, gb)
                if common.within(e, common.TOL_TIGHT):
                    tight += 1
                else:
                    assert common.within(e, common.TOL_CONTACT), e
                    loose += 1
    assert loose <= 0.08 * (tight + loose)


def test_rollout_statistics_match_the_reference(torch_cuda):
    """Free-running collection (RandomState resets, uniform random actions, auto-reset) on both sides: the per-step statistics a
    learner sees — mean reward, episode-end rate, ball height / speed, car height / speed, cars on the ground, boost, flips,
    demolitions — agree within 3 standard errors of the difference (no absolute slack).  1 024 reference gyms against 4 096
    engine arenas, 160 env-steps each after a 60-step warm-up; the standard errors come from the per-gym / per-arena means.
    The two runs share no random numbers, so this is the distribution-level check behind "learning curves statistically
    indistinguishable".  Both sides take the observation statistics over the steps that did NOT end an episode (GameInst::Step
    hands back the reset observation for a finished arena, the reference loop sees the terminal one); reward and the episode-end
    rate are taken over all steps."""
    from oracle import refsim

    if not refsim.available():
        pytest.skip("oracle/_ref/librlref.so not present")
    steps, warm = 220, 60  # warm-up: both sides forget the all-fresh start
    names = ("ball_z", "ball_speed", "car_z", "car_speed", "boost", "on_ground", "has_flip", "demoed")

    def obs_stats(o):
        """o [..., P, 89] -> [..., 8]: ball pos/vel at 0:6, self block at 51 (pos 51:54, vel 60:63, boost 66, onGround 67, hasFlip 68, demoed 69)."""
        return np.stack([o[..., 0, 2], np.linalg.norm(o[..., 0, 3:6], axis=-1), o[..., 53].mean(-1), np.linalg.norm(o[..., 60:63], axis=-1).mean(-1),
                         o[..., 66].mean(-1), o[..., 67].mean(-1), o[..., 68].mean(-1), o[..., 69].mean(-1)], axis=-1)

    def summarise(per_unit):
        """per_unit [units, k] means -> (mean, standard error) over the units (gyms / arenas are independent)."""
        return per_unit.mean(0), per_unit.std(0, ddof=1) / np.sqrt(per_unit.shape[0])

    # reference
    G = 1024
    refsim.seed(4242)
    rng = np.random.default_rng(77)
    g = refsim.RefGym(abi.default_cfg(num_arenas=1, team_size=1))
    ref_units = np.zeros((G, 2 + len(names)))
    for i in range(G):
        g.reset()
        rew, done, ob = [], [], []
        for s in range(steps):
            o, r, d = g.step(rng.integers(0, 90, size=2))
            if s >= warm:
                rew.append(float(np.mean(r))); done.append(float(d))
                if not d:
                    ob.append(obs_stats(o))
            if d:
                g.reset()
        ref_units[i] = np.concatenate([[np.mean(rew), np.mean(done)], np.mean(ob, axis=0)])
    ref_mean, ref_se = summarise(ref_units)

    e = engine.Engine(abi.default_cfg(num_arenas=4096, team_size=1))
    e.reset()
    rng = np.random.default_rng(78)
    A, P = e.A, e.P
    rew_sum, done_sum, ob_sum, ob_cnt = np.zeros(A), np.zeros(A), np.zeros((A, len(names))), np.zeros(A)
    for s in range(steps):
        o, r, d = e.step_host(rng.integers(0, 90, size=A * P).astype(np.int32))
        if s >= warm:
            rew_sum += r.reshape(A, P).mean(1); done_sum += d
            live = d == 0
            ob_sum[live] += obs_stats(o.reshape(A, P, -1)[live]); ob_cnt += live
    n = steps - warm
    eng_units = np.concatenate([(rew_sum / n)[:, None], (done_sum / n)[:, None], ob_sum / np.maximum(ob_cnt, 1)[:, None]], axis=1)
    eng_mean, eng_se = summarise(eng_units)
    keys = ("reward", "done") + names
    report = {k: dict(engine=float(eng_mean[i]), reference=float(ref_mean[i]), se_diff=float(np.hypot(ref_se[i], eng_se[i])),
                      sigmas=float((eng_mean[i] - ref_mean[i]) / np.hypot(ref_se[i], eng_se[i]))) for i, k in enumerate(keys)}
    print(report)
    import json

    for d_ in (os.environ.get("RLG_STATS_DIR"), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")):
        if d_ and os.path.isdir(d_):
            with open(os.path.join(d_, "rollout_stats_r02.json"), "w") as f:
                json.dump(dict(reference_gyms=G, engine_arenas=A, steps=n, stats=report), f, indent=1)
    bad = {k: v for k, v in report.items() if abs(v["sigmas"]) > 3.0}
    assert not bad, bad


def test_state_setters_match_the_reference_distribution(torch_cuda):
    """The device state setters (rl_gym.h gym_reset, run by k_reset and by the auto-reset of k_roles) against the live reference:
    RandomState marginals by two-sample KS tests (common.compare_setter_samples), KickoffState's spawn-slot assignment by a
    chi-square test on the 20 ordered (blue slot, orange slot) pairs + the exact kickoff poses."""
    from oracle import refsim

    if not refsim.available():
        pytest.skip("oracle/_ref/librlref.so not present")
    cfg = abi.default_cfg(num_arenas=4096, team_size=1)
    refsim.seed(5)
    g = refsim.RefGym(cfg)
    ref = []
    for _ in range(4096):
        g.reset()
        c, b, p, _t = g.arena.get_state()
        ref.append((c, b, p))
    e = engine.Engine(cfg)
    e.reset()
    ids = np.arange(e.A, dtype=np.int32)
    cars, balls, pads, _t = e.get_state(ids)
    ours = [(cars[i], balls[i:i + 1], pads[i]) for i in range(e.A)]
    print(common.compare_setter_samples(ours, ref))
    # ... and again after a round of auto-resets inside the fused step (every arena finishes: NoTouch horizon 1 step)
    cfg2 = abi.default_cfg(num_arenas=2048, team_size=1)
    cfg2.no_touch_max_steps = 1
    e2 = engine.Engine(cfg2)
    e2.reset()
    rng = np.random.default_rng(1)
    for _ in range(3):
        _o, _r, d = e2.step_host(rng.integers(0, 90, size=e2.A * e2.P).astype(np.int32))
    assert d.all()
    cars, balls, pads, _t = e2.get_state(np.arange(e2.A, dtype=np.int32))
    print(common.compare_setter_samples([(cars[i], balls[i:i + 1], pads[i]) for i in range(e2.A)], ref))

    # KickoffState (Arena::ResetToRandomKickoff, Arena.cpp:112-216): 5 spawn slots, shuffled, blue takes slot k, orange mirrors
    cfgk = abi.default_cfg(num_arenas=4000, team_size=1)
    cfgk.state_setter = abi.RLG_SETTER_KICKOFF
    ek = engine.Engine(cfgk)
    ek.reset()
    cars, balls, pads, _t = ek.get_state(np.arange(ek.A, dtype=np.int32))
    spots = [(-2048, -2560), (2048, -2560), (-256, -3840), (256, -3840), (0, -4608)]
    yaws = [np.pi / 4, 3 * np.pi / 4, np.pi / 2, np.pi / 2, np.pi / 2]
    gk = refsim.RefGym(cfgk)
    counts = {"ours": np.zeros(5), "ref": np.zeros(5)}

    def slot_of(c):
        x, y = float(c["pos"][0]), float(c["pos"][1])
        yaw = float(np.arctan2(c["rot_forward"][1], c["rot_forward"][0]))
        if c["team"] == 1:
            x, y, yaw = -x, -y, float(np.arctan2(-c["rot_forward"][1], -c["rot_forward"][0]))
        k = spots.index((round(x), round(y)))
        assert abs(yaw - yaws[k]) < 1e-5 and abs(float(c["pos"][2]) - 17.0) < 1e-4 and abs(float(c["boost"]) - 100 / 3) < 1e-4
        return k

    for i in range(ek.A):
        assert np.allclose(balls[i]["pos"], [0, 0, 93.15], atol=1e-4) and np.all(balls[i]["vel"] == 0)
        kb, ko = slot_of(cars[i][0]), slot_of(cars[i][1])
        assert kb == ko  # 1v1: the orange car mirrors the blue car's slot (same index of the shuffled list)
        counts["ours"][kb] += 1
    for _ in range(2000):
        gk.reset()
        c, b, p, _t = gk.arena.get_state()
        kb, ko = slot_of(c[0]), slot_of(c[1])
        assert kb == ko
        counts["ref"][kb] += 1
    for k, n in counts.items():
        chi2 = float((((n - n.sum() / 5) ** 2) / (n.sum() / 5)).sum())
        print(k, n, chi2)
        assert chi2 < 23.5, (k, n, chi2)  # chi-square, 4 dof, alpha = 1e-4
