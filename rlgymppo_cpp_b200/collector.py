"""ctypes binding of the device-resident collector (csrc/collector.cu, include/rlgym_b200.h `rlg_collector_*`).

Mirrors what ThreadAgentManager hands to Learner (reference ThreadAgentManager.cpp:16-80 / GameTrajectory.h:5-18):
after ``collect(n)`` + ``gae(...)`` the seven trajectory tensors plus value targets and advantages exist on the device,
either as T-major views (``view()``) or exported in the reference's concatenated row order (``export_rows``).
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from . import abi
from .engine import Engine, _check, load_library


class CollectorCfg(C.Structure):
    _fields_ = [
        ("num_hidden", C.c_int32), ("policy_hidden", C.c_int32 * 4), ("critic_hidden", C.c_int32 * 4), ("max_steps", C.c_int32),
        ("seed", C.c_uint64), ("temperature", C.c_float), ("deterministic", C.c_int32),
    ]


class TrajView(C.Structure):
    _fields_ = [
        ("T", C.c_int32), ("N", C.c_int32), ("A", C.c_int32), ("P", C.c_int32), ("obs_size", C.c_int32),
        ("obs", C.c_void_p), ("action", C.c_void_p), ("logprob", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p),
        ("value", C.c_void_p), ("advantage", C.c_void_p), ("value_target", C.c_void_p), ("ret", C.c_void_p),
    ]


def default_linear_init(layer_dims: Sequence[Tuple[int, int]], seed: int):
    """torch.nn.Linear's default init (kaiming_uniform(a=sqrt 5) == U(+-1/sqrt(in)) for W and b), the init the reference's
    DiscretePolicy / ValueEstimator get from libtorch (DiscretePolicy.cpp:13-27). numpy Generator, not torch's stream."""
    rng = np.random.default_rng(seed)
    out = []
    for o, i in layer_dims:
        bound = 1.0 / np.sqrt(i)
        out.append((rng.uniform(-bound, bound, size=(o, i)).astype(np.float32), rng.uniform(-bound, bound, size=o).astype(np.float32)))
    return out


class Collector:
    def __init__(self, engine: Engine, policy_hidden=(256, 256, 256), critic_hidden=(256, 256, 256), max_steps=8, seed=123,
                 temperature=1.0, deterministic=False):
        self.L = load_library()
        self.L.rlg_collector_launch_count.restype = C.c_uint64
        self.engine = engine
        assert len(policy_hidden) == len(critic_hidden)
        cfg = CollectorCfg()
        cfg.num_hidden = len(policy_hidden)
        for i, (a, b) in enumerate(zip(policy_hidden, critic_hidden)):
            cfg.policy_hidden[i] = a
            cfg.critic_hidden[i] = b
        cfg.max_steps = max_steps
        cfg.seed = seed
        cfg.temperature = temperature
        cfg.deterministic = int(deterministic)
        self.cfg = cfg
        self.h = C.c_void_p()
        _check(self.L.rlg_collector_create(engine.h, C.byref(cfg), C.byref(self.h)))
        self.policy_dims = self._dims(policy_hidden, engine.num_actions)  # ActionParser::GetActionAmount (90 unless Engine.set_action_table)
        self.critic_dims = self._dims(critic_hidden, 1)
        self.weights = [None, None]

    def _dims(self, hidden, out):
        dims, i = [], self.engine.obs_size
        for h in hidden:
            dims.append((h, i))
            i = h
        dims.append((out, i))
        return dims

    def close(self):
        if getattr(self, "h", None):
            self.L.rlg_collector_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights ------------------------------------------------------------------------------------
    def set_weights(self, net: int, layers):
        """layers: [(W [out,in] f32, b [out] f32), ...] in torch nn.Linear convention; net 0 = policy, 1 = critic."""
        dims = self.policy_dims if net == 0 else self.critic_dims
        assert len(layers) == len(dims)
        keep = []
        for l, ((W, b), (o, i)) in enumerate(zip(layers, dims)):
            W = np.ascontiguousarray(W, dtype=np.float32)
            b = np.ascontiguousarray(b, dtype=np.float32)
            assert W.shape == (o, i) and b.shape == (o,), (W.shape, b.shape, o, i)
            _check(self.L.rlg_collector_set_layer(self.h, net, l, W.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), o, i))
            keep.append((W, b))
        self.weights[net] = keep

    def init_default(self, seed=0):
        self.set_weights(0, default_linear_init(self.policy_dims, seed))
        self.set_weights(1, default_linear_init(self.critic_dims, seed + 1))

    # -- host StateSetter (StateSetter::ResetState(Arena*), G/Utils/StateSetters/StateSetter.h:9) ------------------------------
    def set_state_setter(self, fn):
        """A user StateSetter on the host: ``fn(arena_ids, cars, balls) -> None`` fills ``cars`` ([n, P] abi.CAR_DTYPE, car-id
        order, pre-filled with default CarStates carrying the right car_id / team) and ``balls`` ([n] abi.BALL_DTYPE) for the
        arenas whose episode ended; boost pads come back active (Match.cpp:66-67).  The engine must have been created with
        ``state_setter = RLG_SETTER_HOST``.  The collector calls it between steps (rlg_collector_set_reset_hook) and
        ``reset_with_setter()`` applies it to every arena (GameInst::Start)."""
        e = self.engine
        if e.cfg.state_setter != abi.RLG_SETTER_HOST:
            raise RuntimeError("set_state_setter: create the engine with state_setter = RLG_SETTER_HOST")
        self._setter = fn
        teams = np.array([(c & 1) if e.cfg.spawn_opponents else 0 for c in range(e.P)], dtype=np.int32)

        def apply(ids: np.ndarray, obs_out_ptr):
            n = len(ids)
            if n == 0:
                return
            cars = np.stack([abi.new_cars(e.P) for _ in range(n)])
            cars["car_id"] = np.arange(1, e.P + 1, dtype=np.int32)[None, :]
            cars["team"] = teams[None, :]
            balls = abi.new_balls(n)
            fn(ids, cars, balls)
            pads = np.zeros((n, abi.RLG_NUM_PADS), dtype=abi.PAD_DTYPE)
            pads["is_active"] = 1
            e.set_state(ids, cars, balls, pads, np.full(n, -1, dtype=np.int64))
            mask = np.zeros(e.A, dtype=np.uint8)
            mask[ids] = 1
            if obs_out_ptr:
                _check(self.L.rlg_engine_reset_current_to(e.h, mask.ctypes.data_as(C.c_void_p), C.c_void_p(obs_out_ptr), None))
            else:
                e.reset_current(mask)
            e.sync()

        self._apply_setter = apply
        HOOK = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.c_void_p)

        def hook(user, ids_ptr, n, obs_out):
            apply(np.ctypeslib.as_array(ids_ptr, shape=(n,)).astype(np.int32).copy(), obs_out)

        self._hook = HOOK(hook)  # keep alive
        _check(self.L.rlg_collector_set_reset_hook(self.h, self._hook, None))

    def reset_with_setter(self):
        """GameInst::Start for every arena through the host StateSetter."""
        self._apply_setter(np.arange(self.engine.A, dtype=np.int32), 0)

    # -- calls ----------------------------------------------------------------------------------------
    def infer(self, obs_ptr: int, n_rows: int, counter: int, action_ptr=0, logprob_ptr=0, value_ptr=0):
        v = lambda p: C.c_void_p(p) if p else None
        _check(self.L.rlg_collector_infer(self.h, C.c_void_p(obs_ptr), n_rows, C.c_uint64(counter), v(action_ptr), v(logprob_ptr), v(value_ptr), None))

    def collect(self, n_steps: int):
        _check(self.L.rlg_collector_collect(self.h, n_steps, None))

    def gae(self, gamma=0.99, lam=0.95, return_std=1.0, clip_range=10.0):
        _check(self.L.rlg_collector_gae(self.h, C.c_float(gamma), C.c_float(lam), C.c_float(return_std), C.c_float(clip_range), None))

    def view(self) -> TrajView:
        v = TrajView()
        _check(self.L.rlg_collector_view(self.h, C.byref(v)))
        return v

    def export_rows(self, states=0, actions=0, log_probs=0, rewards=0, next_states=0, dones=0, truncateds=0, value_targets=0, advantages=0):
        v = lambda p: C.c_void_p(p) if p else None
        _check(self.L.rlg_collector_export(self.h, v(states), v(actions), v(log_probs), v(rewards), v(next_states), v(dones), v(truncateds),
                                           v(value_targets), v(advantages), None))

    def read(self, name: str) -> np.ndarray:
        """D2H copy of one T-major view (test plumbing)."""
        v = self.view()
        T, N, A, O = v.T, v.N, v.A, v.obs_size
        shapes = {"obs": ((T + 1, N, O), np.float32), "action": ((T, N), np.int32), "logprob": ((T, N), np.float32),
                  "reward": ((T, N), np.float32), "done": ((T, A), np.uint8), "value": ((T + 1, N), np.float32),
                  "advantage": ((T, N), np.float32), "value_target": ((T, N), np.float32), "ret": ((T, N), np.float32)}
        shape, dt = shapes[name]
        out = np.empty(shape, dtype=dt)
        _check(self.L.rlg_engine_copy_to_host(self.engine.h, out.ctypes.data_as(C.c_void_p), C.c_void_p(getattr(v, name)), C.c_size_t(out.nbytes)))
        return out

    def return_stats(self, n_first: int = 0):
        """(means [3] = mean |returns|, |advantages|, |value targets| of the last collect + GAE; the first n_first returns in the
        reference's concatenation order) — reduced on the device, one small D2H (Learner.cpp:660-682)."""
        means = np.zeros(3, dtype=np.float64)
        first = np.zeros(max(n_first, 1), dtype=np.float32)
        _check(self.L.rlg_collector_return_stats(self.h, means.ctypes.data_as(C.c_void_p), first.ctypes.data_as(C.c_void_p), int(n_first), None))
        v = self.view()
        return means, first[: min(n_first, v.T * v.N)]

    def set_weights_device(self, net: int, layer: int, w_ptr: int, ldw: int, b_ptr: int, out_dim: int, in_dim: int, stream: int = 0):
        _check(self.L.rlg_collector_set_layer_device(self.h, net, layer, C.c_void_p(w_ptr), ldw, C.c_void_p(b_ptr), out_dim, in_dim,
                                                     C.c_void_p(stream) if stream else None))

    def load_external(self, obs, action, logprob, reward, done, value):
        """A trajectory collected elsewhere (host arrays, T-major like the ring) becomes the last collect (rlg_collector_load_external)."""
        T = action.shape[0]
        arrs = [np.ascontiguousarray(obs, np.float32), np.ascontiguousarray(action, np.int32), np.ascontiguousarray(logprob, np.float32),
                np.ascontiguousarray(reward, np.float32), np.ascontiguousarray(done, np.uint8), np.ascontiguousarray(value, np.float32)]
        N, A = self.engine.A * self.engine.P, self.engine.A
        assert arrs[0].shape == (T + 1, N, self.engine.obs_size) and arrs[4].shape == (T, A) and arrs[5].shape == (T + 1, N)
        _check(self.L.rlg_collector_load_external(self.h, T, *[a.ctypes.data_as(C.c_void_p) for a in arrs]))

    def enable_timing(self, on=True):
        _check(self.L.rlg_collector_enable_timing(self.h, int(on)))

    def kernel_times(self):
        """(step_ms_total, step_launches, infer_ms_total, infer_launches) of the last collect (syncs on its events)."""
        sm, im = C.c_double(), C.c_double()
        sn, inn = C.c_int32(), C.c_int32()
        _check(self.L.rlg_collector_kernel_times(self.h, C.byref(sm), C.byref(sn), C.byref(im), C.byref(inn)))
        return sm.value, sn.value, im.value, inn.value

    @property
    def launch_count(self) -> int:
        return int(self.L.rlg_collector_launch_count(self.h))
