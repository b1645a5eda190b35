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


def test_mutators_unsupported_rejected():
    cfg = abi.default_cfg(num_arenas=1, team_size=1)
    cfg.mutators_set = 1
    cfg.mutators.ball_radius = 120.0
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
    learner sees — mean reward, episode-end rate, ball height / speed, cars on the ground, boost, demolitions — agree within
    the reference sample's own standard error (4 sigma + a small absolute slack).  The two runs share no random numbers, so
    this is the distribution-level check behind "learning curves statistically indistinguishable"."""
    from oracle import refsim

    if not refsim.available():
        pytest.skip("oracle/_ref/librlref.so not present")
    cfg = abi.default_cfg(num_arenas=4096, team_size=1)
    steps, warm = 220, 60  # warm-up: both sides forget the all-fresh start

    def stats(obs, rew, done):
        """obs [n, P, 89]: ball pos/vel at 0:6 (scaled by 1/2300, 1/2300), self block at 51 (boost 66, onGround 67, demoed 69)."""
        o = obs.reshape(-1, obs.shape[-1])
        return dict(reward=rew.reshape(-1), done=done.reshape(-1).astype(np.float64), ball_z=o[:, 2], ball_speed=np.linalg.norm(o[:, 3:6], axis=1),
                    car_z=o[:, 53], car_speed=np.linalg.norm(o[:, 60:63], axis=1), boost=o[:, 66], on_ground=o[:, 67], has_flip=o[:, 68],
                    demoed=o[:, 69])

    # reference: G gyms, one (correlated) sample per gym-step; the standard error is taken over per-gym means
    G = 96
    refsim.seed(4242)
    gyms = [refsim.RefGym(abi.default_cfg(num_arenas=1, team_size=1)) for _ in range(G)]
    rng = np.random.default_rng(77)
    per_gym = []
    for g in gyms:
        g.reset()
        acc = {}
        for s in range(steps):
            o, r, d = g.step(rng.integers(0, 90, size=2))
            if s >= warm:
                for k, v in stats(o[None], r[None], np.array([d])).items():
                    acc.setdefault(k, []).append(np.mean(v))
            if d:
                g.reset()
        per_gym.append({k: float(np.mean(v)) for k, v in acc.items()})
    ref_mean = {k: float(np.mean([p[k] for p in per_gym])) for k in per_gym[0]}
    ref_se = {k: float(np.std([p[k] for p in per_gym], ddof=1) / np.sqrt(G)) for k in per_gym[0]}

    e = engine.Engine(cfg)
    e.reset()
    rng = np.random.default_rng(78)
    acc = {}
    for s in range(steps):
        o, r, d = e.step_host(rng.integers(0, 90, size=e.A * e.P).astype(np.int32))
        if s >= warm:
            # GameInst::Step hands back the RESET observation for a finished arena; the reference loop above looks at the
            # terminal one and discards the reset's, so finished arenas are left out of the observation statistics here
            live = d == 0
            st = stats(o.reshape(e.A, e.P, -1)[live], r.reshape(e.A, e.P)[live], d[live])
            st["done"], st["reward"] = d.astype(np.float64), r
            for k, v in st.items():
                acc.setdefault(k, []).append(np.mean(v))
    got = {k: float(np.mean(v)) for k, v in acc.items()}
    slack = dict(reward=0.01, done=0.001, ball_z=0.01, ball_speed=0.01, car_z=0.003, car_speed=0.01, boost=0.01, on_ground=0.008, has_flip=0.01,
                 demoed=0.001)
    out = os.environ.get("RLG_STATS_OUT")
    if out:  # evidence file for profiles/: engine mean, reference mean, reference standard error per statistic
        import json

        with open(out, "w") as f:
            json.dump({k: dict(engine=got[k], reference=ref_mean[k], reference_se=ref_se[k]) for k in got}, f, indent=1)
    bad = {k: (got[k], ref_mean[k], ref_se[k]) for k in got if abs(got[k] - ref_mean[k]) > 4 * ref_se[k] + slack[k]}
    assert not bad, bad
