"""Shared helpers for the parity tests: scenario builders and state comparison.

A *scenario* is (cars, ball, pads, controls_fn, nticks): the same start state is injected into the
reference arena (oracle/_ref) and into the implementation under test, both are stepped tick by
tick with the same controls, and the states are compared after every tick.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rlgymppo_cpp_b200 import abi  # noqa: E402

PHYS_FIELDS = ["pos", "vel", "ang_vel", "rot_forward", "rot_up"]
FLAG_FIELDS = ["is_on_ground", "has_jumped", "has_double_jumped", "has_flipped", "is_flipping", "is_jumping",
               "is_supersonic", "is_auto_flipping", "is_demoed", "hit_valid", "world_contact_has"]
SCALAR_FIELDS = ["jump_time", "flip_time", "air_time", "air_time_since_jump", "boost", "time_spent_boosting",
                 "supersonic_time", "handbrake_val", "auto_flip_timer", "car_contact_cooldown", "demo_respawn_timer",
                 "wheel_steer_angle", "wheel_engine_force", "wheel_brake"]


def yaw_rot(cars, i, yaw, pitch=0.0, roll=0.0):
    """Angle(yaw,pitch,roll).ToRotMat() in numpy float32 (reference MathTypes.cpp:73-78)."""
    y, p, r = np.float32(yaw), np.float32(-pitch), np.float32(-roll)
    ci, cj, ch = np.cos(r), np.cos(p), np.cos(y)
    si, sj, sh = np.sin(r), np.sin(p), np.sin(y)
    cc, cs, sc, ss = ci * ch, ci * sh, si * ch, si * sh
    m = np.array([[cj * ch, sj * sc - cs, sj * cc + ss], [cj * sh, sj * ss + cc, sj * cs - sc], [-sj, cj * si, cj * ci]],
                 dtype=np.float32)
    cars["rot_forward"][i] = m[:, 0]
    cars["rot_right"][i] = m[:, 1]
    cars["rot_up"][i] = m[:, 2]


def make_controls(n, **kw):
    c = np.zeros(n, dtype=abi.CONTROLS_DTYPE)
    for k, v in kw.items():
        c[k] = v
    return c


def compare_states(ref, got, what=""):
    """-> dict of max abs errors for the physical fields + list of mismatching discrete fields."""
    rc, rb, rp, rt = ref
    gc, gb, gp, gt = got
    err = {}
    for f in PHYS_FIELDS:
        err["car_" + f] = float(np.max(np.abs(rc[f].astype(np.float64) - gc[f].astype(np.float64)))) if len(rc) else 0.0
    for f in ("pos", "vel", "ang_vel"):
        err["ball_" + f] = float(np.max(np.abs(rb[f].astype(np.float64) - gb[f].astype(np.float64))))
    for f in SCALAR_FIELDS:
        err[f] = float(np.max(np.abs(rc[f].astype(np.float64) - gc[f].astype(np.float64)))) if len(rc) else 0.0
    mism = []
    for f in FLAG_FIELDS:
        if not np.array_equal(rc[f] != 0, gc[f] != 0):
            mism.append(f)
    if not np.array_equal(rc["wheels_with_contact"] != 0, gc["wheels_with_contact"] != 0):
        mism.append("wheels_with_contact")
    if not np.array_equal(rp["is_active"] != 0, gp["is_active"] != 0):
        mism.append("pads_active")
    if rt != gt:
        mism.append("tick")
    return err, mism


def run_pair(ref_arena, impl, cars, ball, pads, controls_fn, nticks, arena=0, resync_every=0):
    """Step both sides; returns per-tick (err, mism). With resync_every=k the implementation is
    re-seeded from the reference state every k ticks (single-tick parity from identical states when k=1)."""
    ref_arena.set_state(cars, ball, pads, 0)
    impl.set_state(arena, cars, ball, pads, 0)
    out = []
    for t in range(nticks):
        if resync_every and t % resync_every == 0 and t > 0:
            rc, rb, rp, rt = ref_arena.get_state()
            impl.set_state(arena, rc, rb, rp, rt)
        ctl = controls_fn(t)
        ref_arena.step(ctl, 1)
        impl.tick(arena, ctl, 1)
        out.append(compare_states(ref_arena.get_state(), impl.get_state(arena)))
    return out
