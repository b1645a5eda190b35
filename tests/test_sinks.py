"""Render sink (rlgymppo_cpp_b200/sinks.py) against the reference's JSON schema (RenderSender.cpp:22-96, render_receiver.py)."""
import json
import socket

import numpy as np

from rlgymppo_cpp_b200 import abi, sinks


def _arena():
    cars = abi.new_cars(2)
    cars["car_id"] = [1, 2]; cars["team"] = [0, 1]
    cars["pos"][1] = (100, -200, 17); cars["vel"][0] = (500, 0, 0)
    cars["hit_valid"][0] = 1; cars["hit_tick"][0] = 995
    ball = abi.new_balls(1)[0]
    obs = np.zeros((2, 89), dtype=np.float32)
    obs[:, 17:51] = 1; obs[0, 17 + 3] = 0            # blue's view: pad 3 is down
    obs[1, 17:51] = obs[0, 17:51][::-1]              # orange sees the mirrored list
    obs[0, 51 + 15] = 0.33; obs[0, 51 + 16] = 1; obs[0, 51 + 17] = 1
    obs[1, 51 + 15] = 0.9
    return cars, ball, obs


def test_game_state_document_schema():
    cars, ball, obs = _arena()
    acts = np.arange(16, dtype=np.float32).reshape(2, 8)
    doc = sinks.game_state_json(cars, ball, obs, [1, 2], tick=1000, tick_skip=8, score=(2, 1), actions=acts)
    assert set(doc) == {"gamemode", "state", "actions"} and doc["gamemode"] == "soccar"
    st = doc["state"]
    assert set(st) == {"ball", "players", "boost_pads", "team_goals"} and st["team_goals"] == [2, 1]
    assert set(st["ball"]) == {"pos", "forward", "right", "up", "vel", "ang_vel"} and st["ball"]["pos"][2] == np.float32(93.15)
    p0, p1 = st["players"]
    assert set(p0) == {"car_id", "team_num", "phys", "boost_pickups", "is_demoed", "on_ground", "ball_touched", "has_flip", "boost_amount"}
    assert (p0["car_id"], p0["team_num"], p1["car_id"], p1["team_num"]) == (1, 0, 2, 1)
    assert p0["ball_touched"] is True and p1["ball_touched"] is False      # hit 5 ticks ago, tickSkip 8
    assert p0["on_ground"] and p0["has_flip"] and abs(p0["boost_amount"] - 0.33) < 1e-6 and abs(p1["boost_amount"] - 0.9) < 1e-6
    assert len(st["boost_pads"]) == 34 and st["boost_pads"][3] == 0.0 and sum(st["boost_pads"]) == 33
    assert doc["actions"] == acts.tolist()
    json.dumps(doc)  # serialisable


def test_rocketsimvis_packet_over_udp():
    cars, ball, obs = _arena()
    doc = sinks.game_state_json(cars, ball, obs, [1, 2], 1000, 8)
    rx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    rx.bind(("127.0.0.1", 0)); rx.settimeout(5)
    tx = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    tx.sendto(sinks.rocketsimvis_packet(doc), rx.getsockname())
    got = json.loads(rx.recv(65536))
    assert set(got) == {"gamemode", "ball_phys", "cars", "boost_pad_states"}                 # render_receiver.py:18-29
    assert set(got["ball_phys"]) == {"pos", "vel", "ang_vel"} and len(got["cars"]) == 2 and len(got["boost_pad_states"]) == 34
