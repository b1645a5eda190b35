"""ctypes binding of the device PPO learner (csrc/ppo.cu, include/rlgym_b200.h `rlg_ppo_*`): PPOLearner::Learn and the
ExperienceBuffer as hand-written CUDA kernels + the tcgen05 GEMM, no autograd.

Reference: /root/reference/RLGymPPO_CPP/src/private/RLGymPPO_CPP/PPO/PPOLearner.cpp:17-349, PPO/ExperienceBuffer.cpp:12-121.
There is no CPU path: constructing a DevicePPO without the CUDA library or a device raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from .engine import _check, load_library


class PpoCfg(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("obs_size", C.c_int32), ("num_actions", C.c_int32), ("num_hidden", C.c_int32),
        ("policy_hidden", C.c_int32 * 4), ("critic_hidden", C.c_int32 * 4), ("batch_size", C.c_int64), ("mini_batch_size", C.c_int64),
        ("epochs", C.c_int32), ("policy_lr", C.c_float), ("critic_lr", C.c_float), ("ent_coef", C.c_float), ("clip_range", C.c_float),
        ("temperature", C.c_float), ("exp_buffer_size", C.c_int64), ("seed", C.c_uint64), ("world", C.c_int32),
    ]


class PpoReport(C.Structure):
    _fields_ = [
        ("entropy", C.c_double), ("kl", C.c_double), ("ratio", C.c_double), ("value_loss", C.c_double), ("clip_fraction", C.c_double),
        ("policy_update_magnitude", C.c_double), ("critic_update_magnitude", C.c_double),
        ("batches", C.c_int64), ("minibatches", C.c_int64), ("cumulative_model_updates", C.c_int64), ("device_ms", C.c_double),
    ]


ALLREDUCE_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)
PARAMS, GRADS, EXP_AVG, EXP_AVG_SQ = 0, 1, 2, 3


class _DevArray:
    """A float32 device range as a __cuda_array_interface__ object (so torch.as_tensor can alias it for the collective)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class DevicePPO:
    def __init__(self, obs_size: int, num_actions: int, policy_hidden: Sequence[int], critic_hidden: Sequence[int], batch_size: int,
                 mini_batch_size: int = 0, epochs: int = 1, policy_lr: float = 3e-4, critic_lr: float = 3e-4, ent_coef: float = 0.005,
                 clip_range: float = 0.2, temperature: float = 1.0, exp_buffer_size: int = 100_000, seed: int = 0, device: int = 0, world: int = 1):
        self.L = load_library()
        L = self.L
        L.rlg_ppo_buffer_size.restype = C.c_int64
        L.rlg_ppo_model_updates.restype = C.c_int64
        L.rlg_ppo_launch_count.restype = C.c_uint64
        L.rlg_ppo_shuffle_counter.restype = C.c_uint64
        L.rlg_ppo_stream.restype = C.c_void_p
        assert len(policy_hidden) == len(critic_hidden)
        cfg = PpoCfg()
        cfg.device, cfg.obs_size, cfg.num_actions, cfg.num_hidden = device, obs_size, num_actions, len(policy_hidden)
        for i, (a, b) in enumerate(zip(policy_hidden, critic_hidden)):
            cfg.policy_hidden[i], cfg.critic_hidden[i] = a, b
        cfg.batch_size, cfg.mini_batch_size, cfg.epochs = batch_size, mini_batch_size, epochs
        cfg.policy_lr, cfg.critic_lr, cfg.ent_coef, cfg.clip_range, cfg.temperature = policy_lr, critic_lr, ent_coef, clip_range, temperature
        cfg.exp_buffer_size, cfg.seed, cfg.world = exp_buffer_size, seed, world
        self.cfg = cfg
        self.h = C.c_void_p()
        _check(L.rlg_ppo_create(C.byref(cfg), C.byref(self.h)))
        self.dims = [self._dims(obs_size, policy_hidden, num_actions), self._dims(obs_size, critic_hidden, 1)]
        self._hook = None

    @staticmethod
    def _dims(obs, hidden, out) -> List[Tuple[int, int]]:
        dims, i = [], obs
        for h in hidden:
            dims.append((h, i))
            i = h
        dims.append((out, i))
        return dims

    def close(self):
        if getattr(self, "h", None):
            self.L.rlg_ppo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters / optimiser state -----------------------------------------------------------------------------------------------
    def init_weights(self, seed: int):
        _check(self.L.rlg_ppo_init_weights(self.h, C.c_uint64(seed)))

    def set_layers(self, net: int, layers, which: int = PARAMS):
        """layers: [(W [out, in], b [out]), ...] in torch nn.Linear convention; net 0 = policy, 1 = critic."""
        assert len(layers) == len(self.dims[net])
        for l, ((W, b), (o, i)) in enumerate(zip(layers, self.dims[net])):
            W, b = np.ascontiguousarray(W, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)
            assert W.shape == (o, i) and b.shape == (o,), (W.shape, b.shape, o, i)
            _check(self.L.rlg_ppo_set_layer(self.h, which, net, l, W.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), o, i))

    def get_layers(self, net: int, which: int = PARAMS):
        out = []
        for l, (o, i) in enumerate(self.dims[net]):
            W, b = np.empty((o, i), dtype=np.float32), np.empty(o, dtype=np.float32)
            _check(self.L.rlg_ppo_get_layer(self.h, which, net, l, W.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), o, i))
            out.append((W, b))
        return out

    def adam_steps(self) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        _check(self.L.rlg_ppo_adam_steps(self.h, C.byref(a), C.byref(b), 0))
        return a.value, b.value

    def set_adam_steps(self, policy_steps: int, critic_steps: int):
        a, b = C.c_int64(policy_steps), C.c_int64(critic_steps)
        _check(self.L.rlg_ppo_adam_steps(self.h, C.byref(a), C.byref(b), 1))

    def flat(self, which: int = GRADS):
        """(device pointer, float count, policy float count) of one flat vector."""
        ptr, n, n0 = C.c_void_p(), C.c_int64(), C.c_int64()
        _check(self.L.rlg_ppo_flat(self.h, which, C.byref(ptr), C.byref(n), C.byref(n0)))
        return ptr.value, n.value, n0.value

    def set_lr(self, policy_lr: float, critic_lr: float):
        _check(self.L.rlg_ppo_set_lr(self.h, C.c_float(policy_lr), C.c_float(critic_lr)))

    def set_allreduce(self, fn, world: int):
        """fn(ptr, count, stream): in-place SUM all-reduce of `count` floats at device pointer `ptr`, enqueued on `stream`."""
        def hook(_user, ptr, count, stream):
            fn(ptr, count, stream)

        self._hook = ALLREDUCE_HOOK(hook) if fn is not None else None
        _check(self.L.rlg_ppo_set_allreduce_hook(self.h, self._hook, None, int(world)))

    # -- experience buffer ----------------------------------------------------------------------------------------------------------
    def submit(self, states_ptr: int, actions_ptr: int, log_probs_ptr: int, value_targets_ptr: int, advantages_ptr: int, n: int, stream: int = 0):
        _check(self.L.rlg_ppo_submit(self.h, C.c_void_p(states_ptr), C.c_void_p(actions_ptr), C.c_void_p(log_probs_ptr), C.c_void_p(value_targets_ptr),
                                     C.c_void_p(advantages_ptr), C.c_int64(n), C.c_void_p(stream) if stream else None))

    def submit_collector(self, collector, stream: int = 0):
        _check(self.L.rlg_ppo_submit_collector(self.h, collector.h, C.c_void_p(stream) if stream else None))

    @property
    def buffer_size(self) -> int:
        return int(self.L.rlg_ppo_buffer_size(self.h))

    def buffer_read(self):
        n, o = self.buffer_size, self.cfg.obs_size
        out = {"states": np.empty((n, o), np.float32), "actions": np.empty(n, np.int64), "log_probs": np.empty(n, np.float32),
               "values": np.empty(n, np.float32), "advantages": np.empty(n, np.float32)}
        _check(self.L.rlg_ppo_buffer_read(self.h, *[out[k].ctypes.data_as(C.c_void_p) for k in ("states", "actions", "log_probs", "values", "advantages")]))
        return out

    def peek_shuffle(self, counter: int = None) -> np.ndarray:
        perm = np.empty(self.buffer_size, dtype=np.int32)
        c = int(self.L.rlg_ppo_shuffle_counter(self.h)) if counter is None else counter
        _check(self.L.rlg_ppo_peek_shuffle(self.h, perm.ctypes.data_as(C.c_void_p), C.c_uint64(c)))
        return perm

    # -- learn ------------------------------------------------------------------------------------------------------------------------
    def learn(self, stream: int = 0, want_report: bool = True):
        rep = PpoReport()
        _check(self.L.rlg_ppo_learn(self.h, C.byref(rep) if want_report else None, C.c_void_p(stream) if stream else None))
        return rep

    def push_weights(self, collector, stream: int = 0):
        _check(self.L.rlg_ppo_push_weights(self.h, collector.h, C.c_void_p(stream) if stream else None))

    @property
    def stream(self) -> int:
        return int(self.L.rlg_ppo_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.L.rlg_ppo_launch_count(self.h))

    @property
    def model_updates(self) -> int:
        return int(self.L.rlg_ppo_model_updates(self.h))
