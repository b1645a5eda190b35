"""refsim.py — TEST INFRASTRUCTURE ONLY (oracle). ctypes wrapper around oracle/_ref/librlref.so,
i.e. the UNMODIFIED reference compiled by oracle/Makefile plus oracle/ref_harness.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module. The product (rlgymppo_cpp_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from rlgymppo_cpp_b200 import abi, meshes  # noqa: E402

LIB_PATH = os.path.join(_HERE, "_ref", "librlref.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"reference oracle not built: {LIB_PATH} (run `make -C oracle ref`)")
        L = C.CDLL(LIB_PATH)
        L.ref_arena_create.restype = C.c_void_p
        L.ref_gym_create.restype = C.c_void_p
        L.ref_gym_arena.restype = C.c_void_p
        L.ref_bench_collect.restype = C.c_double
        L.ref_bench_create.restype = C.c_void_p
        L.ref_bench_run.restype = C.c_double
        L.ref_sizeof_car_state.restype = C.c_size_t
        assert L.ref_sizeof_car_state() == C.sizeof(abi.CarState)
        blobs = meshes.generate_placeholder_soccar()
        n = len(blobs)
        arr = (C.c_void_p * n)(*[C.cast(C.c_char_p(b), C.c_void_p) for b in blobs])
        sizes = (C.c_size_t * n)(*[len(b) for b in blobs])
        rc = L.ref_init(arr, sizes, n)
        if rc != 0:
            raise RuntimeError("ref_init failed")
        _lib = L
    return _lib


def seed(s: int):
    lib().ref_seed(C.c_uint32(s))


class RefArena:
    """Raw RocketSim Arena (soccar), cars added in Gym::Gym order; indices are car id - 1."""

    def __init__(self, team_size=1, spawn_opponents=True, handle=None, car_preset=0, cfg=None):
        self.L = lib()
        self._own = handle is None
        self.L.ref_arena_create_preset.restype = C.c_void_p
        self.L.ref_arena_create_cfg.restype = C.c_void_p
        if handle is None and cfg is not None:  # car preset + mutators of an EngineCfg
            handle, self._own = self.L.ref_arena_create_cfg(C.byref(cfg)), True
        self.h = C.c_void_p(handle if handle is not None else self.L.ref_arena_create_preset(team_size, int(spawn_opponents), int(car_preset)))
        self.num_cars = self.L.ref_arena_num_cars(self.h)

    def close(self):
        if self._own and self.h:
            self.L.ref_arena_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_state(self, cars=None, ball=None, pads=None, tick_count=-1):
        cp = cars.ctypes.data_as(C.c_void_p) if cars is not None else None
        bp = ball.ctypes.data_as(C.c_void_p) if ball is not None else None
        pp = pads.ctypes.data_as(C.c_void_p) if pads is not None else None
        if cars is not None:
            assert cars.dtype == abi.CAR_DTYPE and len(cars) == self.num_cars
        self.L.ref_arena_set_state(self.h, cp, bp, pp, C.c_int64(tick_count))

    def get_state(self):
        cars = np.zeros(self.num_cars, dtype=abi.CAR_DTYPE)
        ball = np.zeros(1, dtype=abi.BALL_DTYPE)
        pads = np.zeros(abi.RLG_NUM_PADS, dtype=abi.PAD_DTYPE)
        tick = C.c_int64(0)
        self.L.ref_arena_get_state(self.h, cars.ctypes.data_as(C.c_void_p), ball.ctypes.data_as(C.c_void_p),
                                   pads.ctypes.data_as(C.c_void_p), C.byref(tick))
        return cars, ball, pads, tick.value

    def step(self, controls=None, nticks=1):
        cp = None
        if controls is not None:
            assert controls.dtype == abi.CONTROLS_DTYPE and len(controls) == self.num_cars
            cp = controls.ctypes.data_as(C.c_void_p)
        self.L.ref_arena_step(self.h, cp, nticks)

    def dump_contacts(self, max_rows=64):
        out = np.zeros((max_rows, 16), dtype=np.float32)
        n = self.L.ref_arena_dump_contacts(self.h, out.ctypes.data_as(C.c_void_p), max_rows)
        return out[:n]

    def player_order(self):
        ids = np.zeros(self.num_cars, dtype=np.int32)
        self.L.ref_arena_player_order(self.h, ids.ctypes.data_as(C.c_void_p))
        return ids


class RefGym:
    """Reference Gym + Match with the built-in plugins configured from an EngineCfg."""

    def __init__(self, cfg: abi.EngineCfg):
        self.L = lib()
        self.cfg = cfg
        self.h = C.c_void_p(self.L.ref_gym_create(C.byref(cfg)))
        if not self.h:
            raise RuntimeError("ref_gym_create failed")
        self.P = self.L.ref_gym_num_players(self.h)
        self.obs_size = abi.obs_size(cfg)
        self.arena = RefArena(handle=self.L.ref_gym_arena(self.h))

    def close(self):
        if self.h:
            self.L.ref_gym_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        obs = np.zeros((self.P, self.obs_size), dtype=np.float32)
        w = self.L.ref_gym_reset(self.h, obs.ctypes.data_as(C.c_void_p))
        assert w == self.obs_size, (w, self.obs_size)
        return obs

    def reset_from_current(self):
        obs = np.zeros((self.P, self.obs_size), dtype=np.float32)
        w = self.L.ref_gym_reset_from_current(self.h, obs.ctypes.data_as(C.c_void_p))
        assert w == self.obs_size, (w, self.obs_size)
        return obs

    def step(self, actions):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((self.P, self.obs_size), dtype=np.float32)
        rew = np.zeros(self.P, dtype=np.float32)
        done = C.c_uint8(0)
        self.L.ref_gym_step(self.h, actions.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p),
                            rew.ctypes.data_as(C.c_void_p), C.byref(done))
        return obs, rew, bool(done.value)

    def eval_current(self, actions):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        obs = np.zeros((self.P, self.obs_size), dtype=np.float32)
        rew = np.zeros(self.P, dtype=np.float32)
        done = C.c_uint8(0)
        self.L.ref_gym_eval_current(self.h, actions.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p),
                                    rew.ctypes.data_as(C.c_void_p), C.byref(done))
        return obs, rew, bool(done.value)

    def last_state(self):
        score = np.zeros(2, dtype=np.int32)
        last_touch = C.c_int32(0)
        counters = np.zeros((self.P, 8), dtype=np.int32)
        touched = np.zeros(self.P, dtype=np.uint8)
        self.L.ref_gym_last_state(self.h, score.ctypes.data_as(C.c_void_p), C.byref(last_touch),
                                  counters.ctypes.data_as(C.c_void_p), touched.ctypes.data_as(C.c_void_p))
        return score, last_touch.value, counters, touched


def action_table() -> np.ndarray:
    t = np.zeros((256, 8), dtype=np.float32)
    n = lib().ref_action_table(t.ctypes.data_as(C.c_void_p))
    return t[:n].copy()


def bench_collect(cfg: abi.EngineCfg, num_threads: int, gyms_per_thread: int, warmup_steps: int,
                  timed_steps: int, seed_: int = 1) -> float:
    """player-steps/s of the reference's threaded Gym::Step loop (sim only, random actions)."""
    return float(lib().ref_bench_collect(C.byref(cfg), num_threads, gyms_per_thread, warmup_steps, timed_steps,
                                         C.c_uint32(seed_)))


class RefPool:
    """Persistent worker threads stepping T x G reference gyms with the caller's actions (GameInst::Step semantics)."""

    def __init__(self, cfg: abi.EngineCfg, num_threads: int, gyms_per_thread: int, seed_: int = 1):
        self.L = lib()
        self.L.ref_pool_create.restype = C.c_void_p
        self.h = C.c_void_p(self.L.ref_pool_create(C.byref(cfg), num_threads, gyms_per_thread, C.c_uint32(seed_)))
        self.G = num_threads * gyms_per_thread
        self.P = abi.num_players(cfg)
        self.obs_size = int(self.L.ref_pool_obs_size(self.h))
        self.obs = np.zeros((self.G * self.P, self.obs_size), dtype=np.float32)
        self.rew = np.zeros(self.G * self.P, dtype=np.float32)
        self.done = np.zeros(self.G, dtype=np.uint8)

    def reset(self) -> np.ndarray:
        self.L.ref_pool_reset(self.h, self.obs.ctypes.data_as(C.c_void_p))
        return self.obs

    def step(self, actions: np.ndarray):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        self.L.ref_pool_step(self.h, a.ctypes.data_as(C.c_void_p), self.obs.ctypes.data_as(C.c_void_p), self.rew.ctypes.data_as(C.c_void_p),
                             self.done.ctypes.data_as(C.c_void_p))
        return self.obs, self.rew, self.done

    def close(self):
        if self.h:
            self.L.ref_pool_destroy(self.h)
            self.h = None


class RefBench:
    """Persistent multithreaded reference collection loop (Gym::Step + auto-reset, random actions)."""

    def __init__(self, cfg: abi.EngineCfg, num_threads: int, gyms_per_thread: int, seed_: int = 1):
        self.L = lib()
        self.T, self.G = num_threads, gyms_per_thread
        self.h = C.c_void_p(self.L.ref_bench_create(C.byref(cfg), num_threads, gyms_per_thread, C.c_uint32(seed_)))
        self.P = abi.num_players(cfg)

    def run(self, steps: int) -> float:
        """-> seconds for `steps` env-steps on every gym"""
        return float(self.L.ref_bench_run(self.h, steps))

    def player_steps(self, steps: int) -> int:
        return self.T * self.G * steps * self.P

    def close(self):
        if self.h:
            self.L.ref_bench_destroy(self.h)
            self.h = None
