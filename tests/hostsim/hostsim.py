"""TEST TOOL ONLY: ctypes wrapper for tests/hostsim/libhostsim.so (the device headers compiled
with g++ so the CPU-only suite can exercise the kernel logic). Never used by the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
sys.path.insert(0, _ROOT)
from rlgymppo_cpp_b200 import abi, meshes  # noqa: E402

LIB = os.path.join(_HERE, "libhostsim.so")


def build(force=False):
    src = os.path.join(_HERE, "hostsim.cpp")
    hdr_dir = os.path.join(_ROOT, "rlgymppo_cpp_b200", "csrc")
    newest = max([os.path.getmtime(src)] + [os.path.getmtime(os.path.join(hdr_dir, f)) for f in os.listdir(hdr_dir) if f.endswith(".h")])
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", src, "-o", LIB])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.hs_create.restype = C.c_void_p
        L.hs_sizeof_arena.restype = C.c_size_t
        _lib = L
    return _lib


class HostSim:
    def __init__(self, cfg: abi.EngineCfg, blobs=None):
        self.L = lib()
        self.cfg = cfg
        blobs = meshes.generate_placeholder_soccar() if blobs is None else blobs
        n = len(blobs)
        self._blobs = blobs
        arr = (C.c_void_p * n)(*[C.cast(C.c_char_p(b), C.c_void_p) for b in blobs])
        sizes = (C.c_size_t * n)(*[len(b) for b in blobs])
        self.h = C.c_void_p(self.L.hs_create(C.byref(cfg), arr, sizes, n))
        if not self.h:
            raise RuntimeError("hs_create failed")
        self.P = self.L.hs_num_cars(self.h)
        self.obs_size = self.L.hs_obs_size(self.h)

    def __del__(self):
        try:
            if self.h:
                self.L.hs_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_player_order(self, car_ids):
        ids = np.ascontiguousarray(car_ids, dtype=np.int32)
        self.L.hs_set_player_order(self.h, ids.ctypes.data_as(C.c_void_p))

    def set_state(self, arena=0, cars=None, ball=None, pads=None, tick_count=-1):
        cp = cars.ctypes.data_as(C.c_void_p) if cars is not None else None
        bp = ball.ctypes.data_as(C.c_void_p) if ball is not None else None
        pp = pads.ctypes.data_as(C.c_void_p) if pads is not None else None
        self.L.hs_set_state(self.h, arena, cp, bp, pp, C.c_int64(tick_count))

    def get_state(self, arena=0):
        cars = np.zeros(self.P, dtype=abi.CAR_DTYPE)
        ball = np.zeros(1, dtype=abi.BALL_DTYPE)
        pads = np.zeros(abi.RLG_NUM_PADS, dtype=abi.PAD_DTYPE)
        tick = C.c_int64(0)
        self.L.hs_get_state(self.h, arena, cars.ctypes.data_as(C.c_void_p), ball.ctypes.data_as(C.c_void_p),
                            pads.ctypes.data_as(C.c_void_p), C.byref(tick))
        return cars, ball, pads, tick.value

    def tick(self, arena=0, controls=None, nticks=1):
        cp = controls.ctypes.data_as(C.c_void_p) if controls is not None else None
        self.L.hs_tick(self.h, arena, cp, nticks)

    def reset_from_current(self, arena=0):
        obs = np.zeros((self.P, self.obs_size), dtype=np.float32)
        self.L.hs_reset_from_current(self.h, arena, obs.ctypes.data_as(C.c_void_p))
        return obs

    def reset(self, arena=0):
        obs = np.zeros((self.P, self.obs_size), dtype=np.float32)
        self.L.hs_reset(self.h, arena, obs.ctypes.data_as(C.c_void_p))
        return obs

    def step(self, actions, arena=0):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((self.P, self.obs_size), dtype=np.float32)
        rew = np.zeros(self.P, dtype=np.float32)
        done = C.c_uint8(0)
        self.L.hs_step(self.h, arena, actions.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p),
                       rew.ctypes.data_as(C.c_void_p), C.byref(done))
        return obs, rew, bool(done.value)

    def eval_gym(self, actions, arena=0):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((self.P, self.obs_size), dtype=np.float32)
        rew = np.zeros(self.P, dtype=np.float32)
        done = C.c_uint8(0)
        self.L.hs_eval_gym(self.h, arena, actions.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p),
                           rew.ctypes.data_as(C.c_void_p), C.byref(done))
        return obs, rew, bool(done.value)

    def gym_state(self, arena=0):
        score = np.zeros(2, dtype=np.int32)
        last_touch = C.c_int32(0)
        counters = np.zeros((self.P, 8), dtype=np.int32)
        touched = np.zeros(self.P, dtype=np.uint8)
        self.L.hs_gym_state(self.h, arena, score.ctypes.data_as(C.c_void_p), C.byref(last_touch),
                            counters.ctypes.data_as(C.c_void_p), touched.ctypes.data_as(C.c_void_p))
        return score, last_touch.value, counters, touched


def dump_contacts(max_rows=64):
    out = np.zeros((max_rows, 16), dtype=np.float32)
    n = lib().hs_dump_contacts(out.ctypes.data_as(C.c_void_p), max_rows)
    return out[:n]


def action_table():
    t = np.zeros((256, 8), dtype=np.float32)
    n = lib().hs_action_table(t.ctypes.data_as(C.c_void_p))
    return t[:n].copy()
