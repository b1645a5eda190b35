"""ctypes binding of the C-ABI collection engine (csrc/librlgym_b200.so, include/rlgym_b200.h).

This is the same binding a reference maintainer would write against the C ABI (see
INTEGRATION.md); the host-side mirror of the reference's Gym/Match surface lives in
``rlgymppo_cpp_b200.gym``.  There is NO CPU fallback: constructing an Engine without a
usable CUDA device (or without the built extension) raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import abi, meshes

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RLG_B200_LIB") or os.path.join(_HERE, "csrc", "librlgym_b200.so")  # override: A/B builds while profiling

EXPORTS = [
    "rlg_last_error", "rlg_abi_version", "rlg_sizeof_car_state", "rlg_sizeof_engine_cfg", "rlg_engine_cfg_default",
    "rlg_engine_create", "rlg_engine_destroy", "rlg_engine_load_meshes", "rlg_engine_reset", "rlg_engine_reset_current",
    "rlg_engine_set_state", "rlg_engine_get_state", "rlg_engine_tick", "rlg_engine_eval_gym", "rlg_engine_step", "rlg_engine_step_noreset",
    "rlg_engine_outputs", "rlg_engine_obs_size", "rlg_engine_num_players", "rlg_engine_num_arenas",
    "rlg_engine_state_bytes_per_arena", "rlg_engine_player_order", "rlg_engine_set_player_order", "rlg_action_table",
    "rlg_engine_step_host", "rlg_engine_host_buffers", "rlg_engine_step_pinned", "rlg_engine_read_outputs", "rlg_engine_launch_count", "rlg_engine_stream", "rlg_engine_sync",
    "rlg_engine_metrics", "rlg_engine_reset_metrics", "rlg_engine_score_lines", "rlg_gemm_tf32", "rlg_gemm_tf32_fused",
    # collector / plumbing (bound in rlgymppo_cpp_b200.collector)
    "rlg_engine_step_to", "rlg_engine_step_ready", "rlg_engine_step_to_after", "rlg_engine_set_action_table", "rlg_engine_num_actions", "rlg_engine_device", "rlg_engine_arena_id_base", "rlg_engine_copy_to_host", "rlg_engine_copy_to_device",
    "rlg_collector_create", "rlg_collector_destroy", "rlg_collector_set_layer", "rlg_collector_infer", "rlg_collector_collect",
    "rlg_collector_gae", "rlg_collector_view", "rlg_collector_export", "rlg_collector_launch_count",
    "rlg_collector_enable_timing", "rlg_collector_kernel_times", "rlg_collector_set_reset_hook", "rlg_engine_reset_current_to",
    "rlg_collector_set_layer_device", "rlg_collector_return_stats", "rlg_collector_load_external", "rlg_collector_set_step_hook",
    "rlg_engine_step_begin", "rlg_engine_step_end", "rlg_engine_export_gamestates", "rlg_engine_export_gamestates_async", "rlg_engine_export_wait",
    "rlg_engine_reset_to", "rlg_host_alloc", "rlg_host_free", "rlg_device_alloc", "rlg_device_free", "rlg_set_last_error", "rlg_sizeof_gym_state",
    "rlg_sizeof_gym_player",
    # device PPO learner (bound in rlgymppo_cpp_b200.ppo)
    "rlg_ppo_create", "rlg_ppo_destroy", "rlg_ppo_init_weights", "rlg_ppo_set_layer", "rlg_ppo_get_layer", "rlg_ppo_adam_steps", "rlg_ppo_flat",
    "rlg_ppo_set_lr", "rlg_ppo_set_allreduce_hook", "rlg_ppo_submit", "rlg_ppo_submit_collector", "rlg_ppo_buffer_size", "rlg_ppo_buffer_read",
    "rlg_ppo_peek_shuffle", "rlg_ppo_shuffle_counter", "rlg_ppo_learn", "rlg_ppo_push_weights", "rlg_ppo_stream", "rlg_ppo_launch_count",
    "rlg_ppo_model_updates",
]

_lib = None


class EngineError(RuntimeError):
    """Mirrors the reference's RG_ERR_CLOSE -> std::runtime_error (RLGymSim_CPP/Framework.h:17-22)."""


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"CUDA extension not built: {LIB_PATH} missing (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.rlg_last_error.restype = C.c_char_p
        L.rlg_sizeof_car_state.restype = C.c_size_t
        L.rlg_sizeof_engine_cfg.restype = C.c_size_t
        L.rlg_engine_state_bytes_per_arena.restype = C.c_size_t
        L.rlg_engine_launch_count.restype = C.c_uint64
        L.rlg_engine_stream.restype = C.c_void_p
        if L.rlg_sizeof_car_state() != C.sizeof(abi.CarState) or L.rlg_sizeof_engine_cfg() != C.sizeof(abi.EngineCfg):
            raise EngineError("ABI mismatch between rlgymppo_cpp_b200.abi and include/rlgym_b200.h")
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise EngineError(load_library().rlg_last_error().decode("utf-8", "replace"))


def action_table() -> np.ndarray:
    t = np.zeros((abi.RLG_NUM_ACTIONS, 8), dtype=np.float32)
    _check(load_library().rlg_action_table(t.ctypes.data_as(C.c_void_p)))
    return t


class Engine:
    """A device-resident pool of arenas on one GPU."""

    def __init__(self, cfg: abi.EngineCfg, mesh_blobs: Optional[Sequence[bytes]] = None, load_meshes: bool = True):
        self.L = load_library()
        self.cfg = cfg
        self.h = C.c_void_p()
        _check(self.L.rlg_engine_create(C.byref(cfg), C.byref(self.h)))
        self.A = self.L.rlg_engine_num_arenas(self.h)
        self.P = self.L.rlg_engine_num_players(self.h)
        self.obs_size = self.L.rlg_engine_obs_size(self.h)
        self.state_bytes = self.L.rlg_engine_state_bytes_per_arena(self.h)
        if load_meshes:
            self.load_meshes(meshes.generate_placeholder_soccar() if mesh_blobs is None else mesh_blobs)

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.L.rlg_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- setup ----------------------------------------------------------------------------------
    def load_meshes(self, blobs: Sequence[bytes]):
        n = len(blobs)
        self._blobs = list(blobs)
        arr = (C.c_void_p * n)(*[C.cast(C.c_char_p(b), C.c_void_p) for b in self._blobs])
        sizes = (C.c_size_t * n)(*[len(b) for b in self._blobs])
        _check(self.L.rlg_engine_load_meshes(self.h, arr, sizes, n))

    def set_player_order(self, car_ids):
        ids = np.ascontiguousarray(car_ids, dtype=np.int32)
        _check(self.L.rlg_engine_set_player_order(self.h, ids.ctypes.data_as(C.c_void_p)))

    def player_order(self) -> np.ndarray:
        ids = np.zeros(self.P, dtype=np.int32)
        _check(self.L.rlg_engine_player_order(self.h, ids.ctypes.data_as(C.c_void_p)))
        return ids

    # -- state injection / extraction -------------------------------------------------------------
    def set_state(self, arena_ids, cars=None, balls=None, pads=None, tick_counts=None):
        ids = np.ascontiguousarray(arena_ids, dtype=np.int32)
        n = len(ids)

        def p(a, dt, cnt):
            if a is None:
                return None
            assert a.dtype == dt and a.size == cnt, (a.dtype, a.size, cnt)
            return np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)

        tc = None if tick_counts is None else np.ascontiguousarray(tick_counts, dtype=np.int64)
        _check(self.L.rlg_engine_set_state(self.h, ids.ctypes.data_as(C.c_void_p), n, p(cars, abi.CAR_DTYPE, n * self.P),
                                           p(balls, abi.BALL_DTYPE, n), p(pads, abi.PAD_DTYPE, n * abi.RLG_NUM_PADS),
                                           None if tc is None else tc.ctypes.data_as(C.c_void_p)))

    def get_state(self, arena_ids):
        ids = np.ascontiguousarray(arena_ids, dtype=np.int32)
        n = len(ids)
        cars = np.zeros((n, self.P), dtype=abi.CAR_DTYPE)
        balls = np.zeros(n, dtype=abi.BALL_DTYPE)
        pads = np.zeros((n, abi.RLG_NUM_PADS), dtype=abi.PAD_DTYPE)
        ticks = np.zeros(n, dtype=np.int64)
        _check(self.L.rlg_engine_get_state(self.h, ids.ctypes.data_as(C.c_void_p), n, cars.ctypes.data_as(C.c_void_p),
                                           balls.ctypes.data_as(C.c_void_p), pads.ctypes.data_as(C.c_void_p),
                                           ticks.ctypes.data_as(C.c_void_p)))
        return cars, balls, pads, ticks

    # -- stepping ---------------------------------------------------------------------------------
    def reset(self, mask: Optional[np.ndarray] = None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        _check(self.L.rlg_engine_reset(self.h, None if m is None else m.ctypes.data_as(C.c_void_p), None))

    def reset_current(self, mask: Optional[np.ndarray] = None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        _check(self.L.rlg_engine_reset_current(self.h, None if m is None else m.ctypes.data_as(C.c_void_p), None))

    def tick_device(self, controls_ptr: int, nticks: int):
        """controls_ptr: device pointer to rlg_controls[A*P] (car-id order) or 0 to keep the current controls."""
        _check(self.L.rlg_engine_tick(self.h, C.c_void_p(controls_ptr) if controls_ptr else None, nticks, None))

    def step_device(self, actions_ptr: int, auto_reset: bool = True):
        fn = self.L.rlg_engine_step if auto_reset else self.L.rlg_engine_step_noreset
        _check(fn(self.h, C.c_void_p(actions_ptr), None))

    def set_action_table(self, table: np.ndarray):
        """A user ActionParser as its table [n_actions, 8] (throttle, steer, pitch, yaw, roll, jump, boost, handbrake) in place of the
        DiscreteAction table; call it before a Collector is created on this engine (its policy head gets n_actions outputs)."""
        t = np.ascontiguousarray(table, dtype=np.float32)
        if t.ndim != 2 or t.shape[1] != 8:
            raise EngineError("action table must be [n_actions, 8]")
        _check(self.L.rlg_engine_set_action_table(self.h, t.ctypes.data_as(C.c_void_p), int(t.shape[0])))
        self._action_table = t.copy()

    @property
    def action_table(self) -> np.ndarray:
        """The table the fused step parses action indices with: DiscreteAction's 90 rows unless set_action_table replaced it."""
        t = getattr(self, "_action_table", None)
        return t.copy() if t is not None else action_table()

    @property
    def num_actions(self) -> int:
        return int(self.L.rlg_engine_num_actions(self.h))

    def step_ready(self):
        """(device address of the per-block completion flags of the last fused step or 0, its sequence number, arenas per block)."""
        flags, seq, apb = C.c_void_p(), C.c_uint32(), C.c_int()
        _check(self.L.rlg_engine_step_ready(self.h, C.byref(flags), C.byref(seq), C.byref(apb)))
        return int(flags.value or 0), int(seq.value), int(apb.value)

    def eval_gym_device(self, actions_ptr: int):
        _check(self.L.rlg_engine_eval_gym(self.h, C.c_void_p(actions_ptr), None))

    def step_host(self, actions: np.ndarray, want_obs=True):
        """The reference-facing call with HOST buffers: H2D actions, fused step, D2H obs/reward/done."""
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        assert actions.size == self.A * self.P
        obs = np.empty((self.A * self.P, self.obs_size), dtype=np.float32) if want_obs else None
        rew = np.empty(self.A * self.P, dtype=np.float32)
        done = np.empty(self.A, dtype=np.uint8)
        _check(self.L.rlg_engine_step_host(self.h, actions.ctypes.data_as(C.c_void_p),
                                           None if obs is None else obs.ctypes.data_as(C.c_void_p),
                                           rew.ctypes.data_as(C.c_void_p), done.ctypes.data_as(C.c_void_p)))
        return obs, rew, done

    def host_buffers(self):
        """numpy views of the engine's page-locked host buffers: (action_idx [A*P] i32, obs [A*P, obs] f32,
        reward [A*P] f32, done [A] u8); valid for the engine's lifetime."""
        if getattr(self, "_hb", None) is None:
            a, o, r, d = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
            _check(self.L.rlg_engine_host_buffers(self.h, C.byref(a), C.byref(o), C.byref(r), C.byref(d)))
            n = self.A * self.P

            def view(ptr, ctype, count, dtype, shape):
                return np.frombuffer((ctype * count).from_address(ptr.value), dtype=dtype).reshape(shape)

            self._hb = (view(a, C.c_int32, n, np.int32, (n,)), view(o, C.c_float, n * self.obs_size, np.float32, (n, self.obs_size)),
                        view(r, C.c_float, n, np.float32, (n,)), view(d, C.c_uint8, self.A, np.uint8, (self.A,)))
        return self._hb

    def step_pinned(self, want_obs=True):
        """Gym::Step for every arena through the page-locked host buffers of host_buffers() (no staging copies)."""
        _check(self.L.rlg_engine_step_pinned(self.h, 1 if want_obs else 0))
        return self.host_buffers()[1:]

    def read_outputs(self):
        obs = np.empty((self.A * self.P, self.obs_size), dtype=np.float32)
        rew = np.empty(self.A * self.P, dtype=np.float32)
        done = np.empty(self.A, dtype=np.uint8)
        _check(self.L.rlg_engine_read_outputs(self.h, obs.ctypes.data_as(C.c_void_p), rew.ctypes.data_as(C.c_void_p),
                                              done.ctypes.data_as(C.c_void_p)))
        return obs, rew, done

    def output_ptrs(self):
        o, r, d = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(self.L.rlg_engine_outputs(self.h, C.byref(o), C.byref(r), C.byref(d)))
        return o.value, r.value, d.value

    def metrics(self) -> dict:
        """ThreadAgentManager::GetMetrics' reward entries ("Average Step Reward", "Average Episode Reward") + raw totals."""
        m = abi.MetricsHost()
        _check(self.L.rlg_engine_metrics(self.h, C.byref(m)))
        return {k: getattr(m, k) for k, _ in abi.MetricsHost._fields_}

    def score_lines(self) -> np.ndarray:
        """[A, 2] int32: GameState::scoreLine per arena (index 0 = blue scored)."""
        out = np.zeros((self.A, 2), dtype=np.int32)
        _check(self.L.rlg_engine_score_lines(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def reset_metrics(self):
        _check(self.L.rlg_engine_reset_metrics(self.h))

    def sync(self):
        _check(self.L.rlg_engine_sync(self.h))

    @property
    def stream(self) -> int:
        return self.L.rlg_engine_stream(self.h)

    @property
    def launch_count(self) -> int:
        return int(self.L.rlg_engine_launch_count(self.h))
