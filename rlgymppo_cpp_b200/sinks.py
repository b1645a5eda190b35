"""Render and metric sinks of the learner (host glue, no GPU work).

Reference: /root/reference/RLGymPPO_CPP/src/public/RLGymPPO_CPP/Util/RenderSender.cpp:22-97 (GameState -> JSON ->
python_scripts/render_receiver.py -> RocketSimVis over UDP 127.0.0.1:9273) and Util/MetricSender.cpp:7-41
(Report -> python_scripts/metric_receiver.py -> wandb).  The JSON keys are the reference's; the values of one arena come
from the engine: physical state through rlg_engine_get_state, the gym-layer fields (pads in CommonValues order,
boost_amount, on_ground, has_flip) from the arena's own observation rows, i.e. exactly what the fused obs builder computed.
"""
from __future__ import annotations

import json
import socket
from typing import Dict, List, Optional

import numpy as np

UDP_IP, UDP_PORT = "127.0.0.1", 9273  # render_receiver.py:13-14
OBS_PADS, OBS_SELF = slice(17, 51), 51  # DefaultOBS layout: ball 9, prevAction 8, pads 34, self 19 (DefaultOBS.cpp:28-38)


def _v(x) -> List[float]:
    return [float(x[0]), float(x[1]), float(x[2])]


def phys_to_json(pos, fwd, right, up, vel, ang_vel) -> Dict:
    """PhysToJSON (RenderSender.cpp:26-39)."""
    return {"pos": _v(pos), "forward": _v(fwd), "right": _v(right), "up": _v(up), "vel": _v(vel), "ang_vel": _v(ang_vel)}


def game_state_json(cars: np.ndarray, ball: np.ndarray, obs_rows: np.ndarray, player_order, tick: int, tick_skip: int, score=(0, 0),
                    actions: Optional[np.ndarray] = None, gamemode: str = "soccar") -> Dict:
    """RenderSender::Send's document (RenderSender.cpp:41-96) for ONE arena.  cars: [P] abi.CAR_DTYPE in car-id order, ball:
    abi.BALL_DTYPE scalar, obs_rows: [P, obs] the arena's observation rows in player order, player_order: car ids."""
    players = []
    ident = (np.array([1.0, 0, 0]), np.array([0, 1.0, 0]), np.array([0, 0, 1.0]))
    blue_row = None
    for p, car_id in enumerate(player_order):
        c = cars[int(car_id) - 1]
        row = obs_rows[p]
        if int(c["team"]) == 0 and blue_row is None:
            blue_row = row
        touched = bool(c["hit_valid"]) and int(c["hit_tick"]) >= int(tick) - int(tick_skip)  # PlayerData.cpp:20-22
        players.append({
            "car_id": int(c["car_id"]), "team_num": int(c["team"]),
            "phys": phys_to_json(c["pos"], c["rot_forward"], c["rot_right"], c["rot_up"], c["vel"], c["ang_vel"]),
            "boost_pickups": 0,  # not exported by the engine's state readback (RocketSimVis does not draw it)
            "is_demoed": bool(c["is_demoed"]), "on_ground": bool(row[OBS_SELF + 16] != 0), "ball_touched": touched,
            "has_flip": bool(row[OBS_SELF + 17] != 0), "boost_amount": float(row[OBS_SELF + 15]),
        })
    pads_row = blue_row if blue_row is not None else obs_rows[0][::1]
    pads = [float(x) for x in pads_row[OBS_PADS]]
    if blue_row is None:  # orange-only arena: its rows list the pads mirrored (GameState.cpp:84-91)
        pads = pads[::-1]
    state = {"ball": phys_to_json(ball["pos"], *ident, ball["vel"], ball["ang_vel"]), "players": players, "boost_pads": pads,
             "team_goals": [int(score[0]), int(score[1])]}
    doc = {"gamemode": gamemode, "state": state}
    doc["actions"] = [] if actions is None else [[float(v) for v in a] for a in np.asarray(actions, dtype=np.float32)]
    return doc


def rocketsimvis_packet(doc: Dict) -> bytes:
    """render_receiver.py:18-31: what RocketSimVis listens for."""
    st = doc["state"] if "state" in doc else doc
    ball = {k: v for k, v in st["ball"].items() if k not in ("forward", "right", "up")}
    out = {"gamemode": doc.get("gamemode", "soccar"), "ball_phys": ball, "cars": list(st["players"]), "boost_pad_states": st["boost_pads"]}
    return json.dumps(out).encode()


class RenderSender:
    """RLGPC::RenderSender over the engine: send(arena) ships one arena's current state to RocketSimVis."""

    def __init__(self, engine, ip: str = UDP_IP, port: int = UDP_PORT):
        self.engine, self.addr = engine, (ip, port)
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
        self.order = [int(i) for i in engine.player_order()]
        from . import engine as _e  # action table (DiscreteAction.cpp:3-67)

        self.table = _e.action_table()

    def document(self, arena: int = 0, action_idx=None) -> Dict:
        e = self.engine
        cars, balls, _, ticks = e.get_state(np.array([arena], dtype=np.int32))
        obs, _, _ = e.read_outputs()
        rows = obs.reshape(e.A, e.P, -1)[arena]
        score = e.score_lines()[arena]
        actions = None if action_idx is None else self.table[np.asarray(action_idx, dtype=np.int64)]
        return game_state_json(cars[0], balls[0], rows, self.order, int(ticks[0]), int(e.cfg.tick_skip), score, actions)

    def send(self, arena: int = 0, action_idx=None):
        self.sock.sendto(rocketsimvis_packet(self.document(arena, action_idx)), self.addr)


class MetricSender:
    """RLGPC::MetricSender (MetricSender.cpp:7-41 + metric_receiver.py): wandb.init(project, group, name[, id, resume]) and
    run.log(report)."""

    def __init__(self, project: str, group: str, name: str, run_id: str = ""):
        try:
            import wandb
        except Exception as ex:  # noqa: BLE001
            raise RuntimeError(f"MetricSender: Failed to import metrics receiver, exception: {ex!r}")
        kw = dict(project=project, group=group, name=name)
        if run_id:
            kw.update(id=run_id, resume="allow")
        self.run = wandb.init(**kw)
        self.cur_run_id = self.run_id = self.run.id

    def send(self, report: Dict[str, float]):
        self.run.log({k: v for k, v in report.items()})
