"""Generates the golden fixtures under tests/golden/ from the UNMODIFIED reference (oracle/_ref).

Run in the dev container (needs /root/reference to have been compiled by `make -C oracle ref`):

    python tests/golden/make_golden.py

Outputs (committed):
  tick_<name>.npz   reference Arena::Step trajectories: full car/ball/pad state after every tick +
                    the controls applied, per scenario class (free flight, ground drive, jump/flip,
                    ball bounces, ball-mesh, car-ball, kickoff, pads, car-world, random play, car-car).
  gym_<name>.npz    Gym-layer sequences: injected arena states + action indices -> obs / reward / done
                    as Match::BuildObservations / GetRewards / IsDone compute them, per config.
  action_table.npy  the 90x8 DiscreteAction table.
The reference has no golden vectors of its own; these ARE the reference's outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refsim  # noqa: E402
from rlgymppo_cpp_b200 import abi  # noqa: E402
from parity_tools import make_controls, yaw_rot  # noqa: E402


def record(arena, cars, ball, pads, controls_fn, nticks):
    """One trajectory, recorded so that EVERY tick is a single-tick experiment the tests can repeat exactly: the state the
    reference reports (GetState, uu) is re-injected (SetState) before each tick, so the recorded after-state is the reference's
    answer to precisely the recorded before-state — without it the body keeps sub-ulp information the uu round trip drops and
    even the reference cannot reproduce its own recording from the recorded states on sensitive contact ticks.  The arena is
    stepped once before the recording starts: a world that has never stepped still carries btContactSolverInfo's default
    time step (1/60) into its first vehicle update (cold-start quirk of tick count 0, covered by its own live test)."""
    arena.step(None, 1)
    arena.set_state(cars, ball, pads, -1)  # keep the tick count (>= 1: "has stepped" is tick count != 0 on both sides)
    P = arena.num_cars
    C = np.zeros((nticks + 1, P), dtype=abi.CAR_DTYPE)
    B = np.zeros(nticks + 1, dtype=abi.BALL_DTYPE)
    Pd = np.zeros((nticks + 1, abi.RLG_NUM_PADS), dtype=abi.PAD_DTYPE)
    T = np.zeros(nticks + 1, dtype=np.int64)
    U = np.zeros((nticks, P), dtype=abi.CONTROLS_DTYPE)
    c, b, p, t = arena.get_state()
    C[0], B[0], Pd[0], T[0] = c, b[0], p, t
    for i in range(nticks):
        u = controls_fn(i)
        U[i] = u
        arena.set_state(c, b, p, t)
        arena.step(u, 1)
        c, b, p, t = arena.get_state()
        C[i + 1], B[i + 1], Pd[i + 1], T[i + 1] = c, b[0], p, t
    return dict(cars=C, ball=B, pads=Pd, tick=T, controls=U)


def base(P=2):
    cars = abi.new_cars(P)
    ball = abi.new_balls(1)
    pads = abi.new_pads(abi.RLG_NUM_PADS)
    spots = [(3000, -2000), (-3000, 2000), (3000, 2000), (-3000, -2000), (0, -3000), (0, 3000)]
    for i in range(P):
        cars["pos"][i] = (spots[i][0], spots[i][1], 17)
    return cars, ball, pads


def scenarios_1v1(arena):
    out = {}
    z = make_controls(2)
    # free flight with air control
    cars, ball, pads = base()
    cars["pos"][0] = (0, -1000, 800); cars["pos"][1] = (500, 1000, 1200)
    cars["vel"][0] = (300, 500, 200); cars["ang_vel"][0] = (1, 2, -1.5)
    cars["vel"][1] = (-300, 100, -200); cars["ang_vel"][1] = (0.5, -2, 3)
    cars["is_on_ground"] = 0
    yaw_rot(cars, 0, 0.3, 0.2, 0.1); yaw_rot(cars, 1, -2.0, -0.5, 1.0)
    ball["pos"][0] = (100, 200, 900); ball["vel"][0] = (800, -300, 400); ball["ang_vel"][0] = (1, 2, 3)
    u = make_controls(2, throttle=1, pitch=0.5, yaw=-1, roll=1, boost=1)
    out["free_flight"] = record(arena, cars, ball, pads, lambda t: u, 60)
    # ground drive
    cars, ball, pads = base()
    cars["pos"][0] = (0, -2000, 17); cars["pos"][1] = (1000, 2000, 17); yaw_rot(cars, 0, 1.0); yaw_rot(cars, 1, -2.0)
    u1 = make_controls(2, throttle=1, steer=0.5)
    out["ground_drive"] = record(arena, cars, ball, pads, lambda t: u1, 120)
    u2 = make_controls(2, throttle=1, steer=-1, boost=1, handbrake=1)
    out["ground_powerslide"] = record(arena, cars, ball, pads, lambda t: u2, 120)

    def jumpctl(t):
        c = make_controls(2, throttle=1)
        c["jump"] = 1 if (t < 10 or (30 <= t < 32)) else 0
        c["pitch"] = -1 if t >= 30 else 0
        return c
    out["jump_flip"] = record(arena, cars, ball, pads, jumpctl, 150)

    def djctl(t):
        c = make_controls(2)
        c["jump"] = 1 if (t < 4 or (20 <= t < 22)) else 0
        c["yaw"] = 1 if t > 40 else 0
        c["roll"] = -1 if t > 60 else 0
        return c
    out["double_jump_air_roll"] = record(arena, cars, ball, pads, djctl, 120)
    # ball alone
    cars, ball, pads = base()
    ball["pos"][0] = (0, 0, 500); ball["vel"][0] = (500, 300, -200); ball["ang_vel"][0] = (2, 1, 0)
    out["ball_bounce_floor"] = record(arena, cars, ball, pads, lambda t: z, 400)
    cars, ball, pads = base()
    ball["pos"][0] = (3500, 0, 800); ball["vel"][0] = (2000, 300, 100)
    out["ball_side_wall"] = record(arena, cars, ball, pads, lambda t: z, 200)
    cars, ball, pads = base()
    ball["pos"][0] = (2000, 4500, 800); ball["vel"][0] = (100, 2500, 100); ball["ang_vel"][0] = (1, 0, 0)
    out["ball_back_wall_mesh"] = record(arena, cars, ball, pads, lambda t: z, 200)
    cars, ball, pads = base()
    ball["pos"][0] = (3000, 4000, 500); ball["vel"][0] = (1500, 1500, 0); ball["ang_vel"][0] = (0, 0, 2)
    out["ball_corner_mesh"] = record(arena, cars, ball, pads, lambda t: z, 200)
    cars, ball, pads = base()
    ball["pos"][0] = (200, 4500, 300); ball["vel"][0] = (50, 3000, 100)
    out["ball_into_goal"] = record(arena, cars, ball, pads, lambda t: z, 200)
    cars, ball, pads = base()
    ball["pos"][0] = (3300, 3000, 93.15); ball["vel"][0] = (800, 900, 0)
    out["ball_roll_ramp"] = record(arena, cars, ball, pads, lambda t: z, 300)
    # car-ball
    cars, ball, pads = base()
    cars["pos"][0] = (0, -1000, 17); yaw_rot(cars, 0, np.pi / 2); cars["vel"][0] = (0, 1000, 0)
    ball["pos"][0] = (30, 0, 93.15)
    ub = make_controls(2, throttle=1, boost=1)
    out["car_hits_ball"] = record(arena, cars, ball, pads, lambda t: ub, 200)
    cars, ball, pads = base()
    cars["pos"][0] = (-2048, -2560, 17); yaw_rot(cars, 0, np.pi / 4)
    cars["pos"][1] = (2048, 2560, 17); yaw_rot(cars, 1, np.pi / 4 + np.pi)
    out["kickoff"] = record(arena, cars, ball, pads, lambda t: ub, 300)
    cars, ball, pads = base()
    cars["pos"][0] = (0, 0, 600); yaw_rot(cars, 0, 0.5, 0.8, 1.5); cars["vel"][0] = (300, 0, -800); cars["is_on_ground"] = 0
    out["car_lands_tumbling_hits_ball"] = record(arena, cars, ball, pads, lambda t: z, 200)
    # pads
    cars, ball, pads = base()
    cars["pos"][0] = (-3584, -1000, 17); yaw_rot(cars, 0, np.pi / 2); cars["boost"][0] = 10
    cars["pos"][1] = (0, 1500, 17); yaw_rot(cars, 1, -np.pi / 2); cars["boost"][1] = 0
    ut = make_controls(2, throttle=1)
    out["boost_pads"] = record(arena, cars, ball, pads, lambda t: ut, 400)
    # car-world
    cars, ball, pads = base()
    cars["pos"][0] = (3000, 0, 17); yaw_rot(cars, 0, 0.0); cars["vel"][0] = (1500, 0, 0)
    out["car_up_side_ramp"] = record(arena, cars, ball, pads, lambda t: ut, 200)
    cars, ball, pads = base()
    cars["pos"][0] = (0, 4000, 17); yaw_rot(cars, 0, np.pi / 2); cars["vel"][0] = (0, 1200, 0)
    out["car_into_goal"] = record(arena, cars, ball, pads, lambda t: ut, 250)
    # car-car
    cars, ball, pads = base()
    cars["pos"][0] = (0, -800, 17); yaw_rot(cars, 0, np.pi / 2); cars["vel"][0] = (0, 1000, 0)
    cars["pos"][1] = (10, 800, 17); yaw_rot(cars, 1, -np.pi / 2); cars["vel"][1] = (0, -1000, 0)
    ball["pos"][0] = (2000, 0, 93.15)
    out["car_car_head_on"] = record(arena, cars, ball, pads, lambda t: ut, 150)
    cars, ball, pads = base()
    cars["pos"][0] = (0, -2500, 17); yaw_rot(cars, 0, np.pi / 2); cars["vel"][0] = (0, 2250, 0); cars["boost"][0] = 100
    cars["pos"][1] = (0, 500, 17); yaw_rot(cars, 1, 0.0)
    ball["pos"][0] = (2000, 0, 93.15)

    def demo_ctl(t):
        c = make_controls(2)
        c["throttle"][0] = 1; c["boost"][0] = 1
        return c
    out["car_car_demo"] = record(arena, cars, ball, pads, demo_ctl, 500)
    return out


def random_play(team, nsteps, seed, car_preset=0, mutate=None):
    """Random DiscreteAction play from RandomState resets: the workload distribution of the benchmark."""
    cfg = abi.default_cfg(num_arenas=1, team_size=team)
    cfg.car_preset = car_preset
    if mutate is not None:
        mutate(cfg)
    g = refsim.RefGym(cfg)
    refsim.seed(seed)
    rng = np.random.default_rng(seed)
    table = refsim.action_table()
    P = 2 * team
    arena = refsim.RefArena(cfg=cfg)
    chunks = []
    for ep in range(nsteps):
        g.reset()
        cars, ball, pads, _ = g.arena.get_state()
        ctl_seq = []
        for s in range(12):
            acts = rng.integers(0, 90, size=P)
            u = np.zeros(P, dtype=abi.CONTROLS_DTYPE)
            for i in range(P):
                a = table[acts[i]]
                u[i] = (a[0], a[1], a[2], a[3], a[4], int(a[5] == 1), int(a[6] == 1), int(a[7] == 1))
            ctl_seq += [u] * 8
        chunks.append(record(arena, cars, ball[0:1], pads, lambda t: ctl_seq[t], len(ctl_seq)))
    return chunks


def gym_sequence(cfg, nepisodes, nsteps, seed):
    """Injected states + actions -> obs/reward/done via the reference plugins (ref_gym_eval_current)."""
    g = refsim.RefGym(cfg)
    team = cfg.team_size
    P = g.P
    phys = refsim.RefArena(team, bool(cfg.spawn_opponents))  # physics runs here so that Gym callbacks do not fire
    refsim.seed(seed)
    rng = np.random.default_rng(seed)
    table = refsim.action_table()
    order = g.arena.player_order()
    recs = dict(cars=[], ball=[], pads=[], tick=[], actions=[], obs=[], reward=[], done=[], first=[])
    for ep in range(nepisodes):
        g.reset()
        st = g.arena.get_state()
        phys.set_state(*st)
        st = phys.get_state()
        g.arena.set_state(*st)
        obs0 = g.reset_from_current()
        recs["cars"].append(st[0]); recs["ball"].append(st[1][0]); recs["pads"].append(st[2]); recs["tick"].append(st[3])
        recs["actions"].append(np.zeros(P, dtype=np.int32)); recs["obs"].append(obs0)
        recs["reward"].append(np.zeros(P, dtype=np.float32)); recs["done"].append(0); recs["first"].append(1)
        for s in range(nsteps):
            acts = rng.integers(0, 90, size=P).astype(np.int32)
            u = np.zeros(P, dtype=abi.CONTROLS_DTYPE)
            for i in range(P):
                a = table[acts[i]]
                u[order[i] - 1] = (a[0], a[1], a[2], a[3], a[4], int(a[5] == 1), int(a[6] == 1), int(a[7] == 1))
            phys.step(u, cfg.tick_skip)
            st = phys.get_state()
            g.arena.set_state(*st)
            o, r, d = g.eval_current(acts)
            recs["cars"].append(st[0]); recs["ball"].append(st[1][0]); recs["pads"].append(st[2]); recs["tick"].append(st[3])
            recs["actions"].append(acts); recs["obs"].append(o); recs["reward"].append(r); recs["done"].append(int(d)); recs["first"].append(0)
            if d:
                break
    out = {k: np.stack(v) if k not in ("tick", "done", "first") else np.asarray(v) for k, v in recs.items()}
    out["player_order"] = np.asarray(order, dtype=np.int32)
    return out


def save(name, d):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(name, {k: v.shape for k, v in d.items()}, os.path.getsize(path) // 1024, "KiB")


def main():
    arena = refsim.RefArena(1, True)
    sc = scenarios_1v1(arena)
    flat = {}
    for name, d in sc.items():
        for k, v in d.items():
            flat[f"{name}/{k}"] = v
    save("tick_scenarios_1v1", flat)
    for team, n, seed in ((1, 6, 11), (2, 4, 12), (3, 3, 13)):
        chunks = random_play(team, n, seed)
        flat = {}
        for i, d in enumerate(chunks):
            for k, v in d.items():
                flat[f"ep{i}/{k}"] = v
        save(f"tick_random_{team}v{team}", flat)
    np.save(os.path.join(HERE, "action_table.npy"), refsim.action_table())
    if len(sys.argv) > 1 and sys.argv[1] == "ticks":  # python tests/golden/make_golden.py ticks — every tick_* file, nothing else
        presets(); mutators(); mutator_scenarios(); ball_mutators()
        return
    # gym layer
    cfg = abi.default_cfg(num_arenas=1, team_size=1)
    for k in range(11):
        cfg.reward_terms[3].params[k] = 0.1 * (k + 1)
    save("gym_1v1_default", gym_sequence(cfg, 12, 40, 21))
    cfg = abi.default_cfg(num_arenas=1, team_size=2)
    cfg.zero_sum = 1; cfg.team_spirit = 0.3
    cfg.num_reward_terms = 5
    cfg.reward_terms[4].kind = abi.RLG_REW_VELOCITY; cfg.reward_terms[4].weight = 0.25
    save("gym_2v2_zerosum", gym_sequence(cfg, 8, 40, 22))
    cfg = abi.default_cfg(num_arenas=1, team_size=3)
    cfg.obs_kind = abi.RLG_OBS_PADDED; cfg.obs_max_players = 3
    save("gym_3v3_padded", gym_sequence(cfg, 5, 40, 23))
    cfg = abi.default_cfg(num_arenas=1, team_size=2)
    cfg.obs_kind = abi.RLG_OBS_PADDED; cfg.obs_max_players = 3
    cfg.zero_sum = 1; cfg.team_spirit = 0.3
    cfg.state_setter = abi.RLG_SETTER_KICKOFF
    save("gym_2v2_padded_zerosum_kickoff", gym_sequence(cfg, 5, 40, 24))
    extra()


def extra():
    """Fixtures added after the first set (kept separate so that the earlier files are not rewritten):
    python tests/golden/make_golden.py extra"""
    import common

    save("gym_1v1_extra_rewards", gym_sequence(common.extra_rewards_cfg(), 14, 40, 25))
    presets()


def ppo():
    """python tests/golden/make_golden.py ppo — the reference's own ComputeGAE / DiscretePolicy / ValueEstimator outputs
    (oracle/_ref/librlref_ppo.so: the unmodified reference TUs against the pip libtorch) on seeded inputs."""
    from oracle import refppo

    rng = np.random.default_rng(41)
    out = {}
    for i, (n, gamma, lam, std, clip) in enumerate([(257, 0.99, 0.95, 1.0, 10.0), (64, 0.995, 0.9, 2.5, 0.5), (40, 0.9, 1.0, 0.0, 10.0), (33, 0.99, 0.95, 0.37, 0.0)]):
        r = (rng.standard_normal(n) * 3).astype(np.float32)
        d = (rng.random(n) < 0.08).astype(np.float32)
        t = ((rng.random(n) < 0.05) & (d == 0)).astype(np.float32)
        v = rng.standard_normal(n + 1).astype(np.float32)
        adv, tgt, ret = refppo.compute_gae(r, d, t, v, gamma, lam, std, clip)
        out.update({f"gae{i}/rews": r, f"gae{i}/dones": d, f"gae{i}/truncated": t, f"gae{i}/values": v, f"gae{i}/params": np.array([gamma, lam, std, clip], np.float64),
                    f"gae{i}/adv": adv, f"gae{i}/target": tgt, f"gae{i}/ret": ret})
    dims = [(64, 89), (64, 64), (90, 64)]
    layers = [((rng.standard_normal(s) * (1.5 / np.sqrt(s[1]))).astype(np.float32), (rng.standard_normal(s[0]) * 0.1).astype(np.float32)) for s in dims]
    obs = rng.uniform(-1.5, 1.5, size=(48, 89)).astype(np.float32)
    acts = rng.integers(0, 90, size=48)
    for temp, tag in ((1.0, "t1"), (0.7, "t07")):
        probs, arg, lp, ent = refppo.policy(layers, obs, acts, temp)
        out.update({f"policy_{tag}/probs": probs, f"policy_{tag}/argmax": arg, f"policy_{tag}/logprob": lp, f"policy_{tag}/entropy": np.float32(ent)})
    cl = layers[:-1] + [((rng.standard_normal((1, 64)) * 0.2).astype(np.float32), np.array([0.05], np.float32))]
    out["critic/values"] = refppo.critic(cl, obs)
    for l, (W, b) in enumerate(layers):
        out[f"net/W{l}"] = W; out[f"net/b{l}"] = b
    out["net/Wc"], out["net/bc"] = cl[-1]
    out["net/obs"], out["net/acts"] = obs, acts
    # ExperienceBuffer FIFO: (maxSize, submit sizes) incl. overflow, a submit larger than the buffer, exact fill
    for i, (max_size, sizes, bs) in enumerate([(10, [4, 4, 4, 3], 4), (8, [20], 3), (12, [5, 7, 1], 5), (16, [3, 2], 4)]):
        starts = np.cumsum([0] + sizes[:-1]).astype(np.float32) * 1.0 + 100 * np.arange(len(sizes), dtype=np.float32)
        cur, st, ac, ad, nb = refppo.buffer_fifo(max_size, 3, sizes, starts, bs)
        out.update({f"buf{i}/max_size": np.int64(max_size), f"buf{i}/sizes": np.array(sizes, np.int32), f"buf{i}/starts": starts, f"buf{i}/batch": np.int64(bs),
                    f"buf{i}/cur": np.int64(cur), f"buf{i}/states": st, f"buf{i}/actions": ac, f"buf{i}/advantages": ad, f"buf{i}/num_batches": np.int64(nb)})
    save("ppo_reference", out)


def mutators():
    """python tests/golden/make_golden.py mutators — random play under a non-default MutatorConfig (common.mutated_cfg)"""
    import common

    for team, n, seed in ((1, 6, 51), (2, 3, 52)):
        chunks = random_play(team, n, seed, mutate=common.apply_test_mutators)
        flat = {}
        for i, d in enumerate(chunks):
            for k, v in d.items():
                flat[f"ep{i}/{k}"] = v
        save(f"tick_random_{team}v{team}_mutators", flat)


def mutator_scenarios():
    """python tests/golden/make_golden.py mutator_scenarios — the scripted 1v1 scenarios (bumps, demos, ball hits, pads, jumps) and
    two mutator-specific ones, recorded from the reference under common.apply_test_mutators."""
    import common

    cfg = common.apply_test_mutators(abi.default_cfg(num_arenas=1, team_size=1))
    arena = refsim.RefArena(cfg=cfg)
    sc = scenarios_1v1(arena)
    # unlimitedFlips: a second dodge in the same jump (Car.cpp:665-671)
    cars, ball, pads = base()
    cars["pos"][0] = (0, -2000, 17); cars["pos"][1] = (1000, 2000, 17); yaw_rot(cars, 0, 1.0); yaw_rot(cars, 1, -2.0)

    def flips(t):
        c = make_controls(2, throttle=1)
        c["jump"] = 1 if (t < 10 or (30 <= t < 32) or (75 <= t < 77) or (120 <= t < 122)) else 0
        c["pitch"] = -1 if (30 <= t < 40 or 120 <= t < 130) else 0
        c["yaw"] = 1 if 75 <= t < 85 else 0
        return c
    sc["unlimited_flips"] = record(arena, cars, ball, pads, flips, 220)
    flat = {}
    for name, d in sc.items():
        for k, v in d.items():
            flat[f"{name}/{k}"] = v
    save("tick_scenarios_1v1_mutators", flat)
    # enableTeamDemos + DemoMode::ON_CONTACT: a slow bump between team mates demolishes (Arena.cpp:375-391), 2v2
    cfg2 = common.apply_test_mutators(abi.default_cfg(num_arenas=1, team_size=2))
    arena2 = refsim.RefArena(cfg=cfg2)
    out = {}
    for name, victim in (("team_mate_demo", 2), ("opponent_demo", 1)):  # car ids 1, 3 are blue (Gym.cpp:46-50 add order)
        cars, ball, pads = base(4)
        cars["pos"][0] = (0, -800, 17); yaw_rot(cars, 0, np.pi / 2); cars["vel"][0] = (0, 900, 0)
        cars["pos"][victim] = (10, 300, 17); yaw_rot(cars, victim, 0.0)
        ball["pos"][0] = (2500, 0, 93.15)

        def ram(t):
            c = make_controls(4)
            c["throttle"][0] = 1
            return c
        out[name] = record(arena2, cars, ball, pads, ram, 260)
    flat = {}
    for name, d in out.items():
        for k, v in d.items():
            flat[f"{name}/{k}"] = v
    save("tick_scenarios_2v2_mutators", flat)


def ball_mutators():
    """python tests/golden/make_golden.py ball_mutators — MutatorConfig::ballMass / ballRadius / carMass (common.apply_ball_mutators): random
    play and the scripted 1v1 scenarios (ball hits, dribbles, wall and ground bounces, wheels on the ball) from the reference"""
    import common

    chunks = random_play(1, 6, 61, mutate=common.apply_ball_mutators)
    flat = {}
    for i, d in enumerate(chunks):
        for k, v in d.items():
            flat[f"ep{i}/{k}"] = v
    save("tick_random_1v1_ballmut", flat)
    cfg = common.apply_ball_mutators(abi.default_cfg(num_arenas=1, team_size=1))
    arena = refsim.RefArena(cfg=cfg)
    flat = {}
    for name, d in scenarios_1v1(arena).items():
        for k, v in d.items():
            flat[f"{name}/{k}"] = v
    save("tick_scenarios_1v1_ballmut", flat)


def presets():
    """python tests/golden/make_golden.py presets — random play with the five non-Octane CarConfigs"""
    for preset, name in ((1, "dominus"), (2, "plank"), (3, "breakout"), (4, "hybrid"), (5, "merc")):
        chunks = random_play(1, 4, 30 + preset, car_preset=preset)
        flat = {}
        for i, d in enumerate(chunks):
            for k, v in d.items():
                flat[f"ep{i}/{k}"] = v
        save(f"tick_random_1v1_{name}", flat)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ticks":
        main()
    elif len(sys.argv) > 1 and sys.argv[1] == "extra":
        extra()
    elif len(sys.argv) > 1 and sys.argv[1] == "presets":
        presets()
    elif len(sys.argv) > 1 and sys.argv[1] == "mutators":
        mutators()
    elif len(sys.argv) > 1 and sys.argv[1] == "ball_mutators":
        ball_mutators()
    elif len(sys.argv) > 1 and sys.argv[1] == "mutator_scenarios":
        mutator_scenarios()
    elif len(sys.argv) > 1 and sys.argv[1] == "ppo":
        ppo()
    else:
        main()
