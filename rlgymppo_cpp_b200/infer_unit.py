"""InferUnit on the device engine: run a saved policy OR critic on game states, outside of training.

Reference: /root/reference/RLGymPPO_CPP/src/public/RLGymPPO_CPP/Util/InferUnit.h:8-43, InferUnit.cpp:11-138.  The
reference's unit owns an OBSBuilder + ActionParser + one network loaded from ``PPO_POLICY.lt`` / ``PPO_CRITIC.lt`` and turns
a ``GameState`` into actions (or values).  Here the OBS builder and the action parser are the engine's (``rlg_engine_cfg``
obs kind, the DiscreteAction table), the observations are the engine's device obs rows of its CURRENT states (or rows the
caller hands in: states that come from somewhere else, e.g. a live game), and the network runs on the tcgen05 inference
kernel (``rlg_collector_infer`` -> ``k_mlp_infer``); the action DISTRIBUTION, which that kernel does not export, is the
same network through the hand-written TF32 GEMM (``gemm.MLPTF32``) + softmax / clamp.  There is no CPU path: the
reference's ``gpu=false`` default has no counterpart, the unit needs the CUDA library and a device.

Rows are (arena, player-slot) pairs in the engine's order, ``row = arena * P + slot`` (``Engine.player_order`` gives the car
id of a slot), which is the order of ``GameState::players`` per arena (InferUnit.cpp:46-50).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import abi
from .engine import Engine, EngineError, action_table

ACTION_MIN_PROB = 1e-11  # DiscretePolicy.h:9


class InferUnit:
    def __init__(self, engine: Engine, model_path: str, is_policy: bool, obs_size: int, layer_sizes: Sequence[int], seed: int = 0):
        """InferUnit::InferUnit (InferUnit.cpp:11-43).  ``model_path``: a ``PPO_POLICY.lt`` / ``PPO_CRITIC.lt`` written by the
        reference or by ``checkpoint.save_seq``; ``obs_size`` must be the engine's (the reference trusts the caller and fails
        inside torch; here it is checked)."""
        import torch

        from . import checkpoint
        from .learner import make_mlp

        if not torch.cuda.is_available():
            raise EngineError("InferUnit needs a CUDA device; there is no CPU fallback")
        if obs_size != engine.obs_size:
            raise EngineError(f"InferUnit: obsSize {obs_size} does not match the engine's OBS builder ({engine.obs_size})")
        self.torch = torch
        self.engine = engine
        self.is_policy = bool(is_policy)
        self.layer_sizes = tuple(int(x) for x in layer_sizes)
        self.num_actions = engine.num_actions  # actionParser->GetActionAmount()
        self.device = torch.device("cuda", engine.L.rlg_engine_device(engine.h))
        self.table = engine.action_table  # ActionParser::ParseActions lookup (DiscreteAction.cpp unless the engine carries a user table)
        seq = make_mlp(obs_size, list(self.layer_sizes), self.num_actions if self.is_policy else 1)
        checkpoint.load_seq(seq, os.fspath(model_path))  # raises like RG_ERR_CLOSE("Failed to load model ...")
        self.seq = seq.to(self.device)
        self.seed = seed
        self._collectors: Dict[Tuple[bool, float], object] = {}
        self._mlp = None
        self._counter = 0
        self._rows = engine.A * engine.P
        self._act = torch.empty(self._rows, dtype=torch.int32, device=self.device)
        self._val = torch.empty(self._rows, dtype=torch.float32, device=self.device)

    # ---- plumbing ---------------------------------------------------------------------------------------------------------
    def _collector(self, deterministic: bool, temperature: float):
        """One inference context per (deterministic, temperature): both are launch constants of k_mlp_infer
        (rlg_collector_cfg), the reference sets policy->temperature per call (InferUnit.cpp:67)."""
        from .collector import Collector
        from .learner import mlp_layers_numpy

        key = (bool(deterministic), float(temperature))
        c = self._collectors.get(key)
        if c is None:
            c = Collector(self.engine, policy_hidden=self.layer_sizes, critic_hidden=self.layer_sizes, max_steps=1, seed=self.seed,
                          temperature=key[1], deterministic=key[0])
            mine = mlp_layers_numpy(self.seq)
            own, other = (0, 1) if self.is_policy else (1, 0)
            c.set_weights(own, mine)
            # the network this unit was NOT created for: zero weights, never read back (its output pointer stays NULL)
            c.set_weights(other, [(np.zeros(d, dtype=np.float32), np.zeros(d[0], dtype=np.float32))
                                  for d in (c.critic_dims if self.is_policy else c.policy_dims)])
            self._collectors[key] = c
        return c

    def _assert_type(self, want_policy: bool):
        if want_policy != self.is_policy:  # ASSERT_RIGHT_TYPE (InferUnit.cpp:52-53)
            a, b = ("policy", "critic") if want_policy else ("critic", "policy")
            raise EngineError(f"InferUnit: Failed to infer the {a} because this inference unit was created to infer the {b}")

    def _obs_dev(self, obs: Optional[np.ndarray]):
        """Device obs rows: the engine's current ones, or caller rows [n, obs_size] copied in."""
        torch = self.torch
        if obs is None:
            ptr, _, _ = self.engine.output_ptrs()
            self.engine.sync()
            return ptr, self._rows, None
        o = np.ascontiguousarray(obs, dtype=np.float32).reshape(-1, self.engine.obs_size)
        if o.shape[0] > self._rows:
            raise EngineError(f"InferUnit: at most {self._rows} rows per call (engine arenas x players)")
        t = torch.from_numpy(o).to(self.device)
        torch.cuda.current_stream(self.device).synchronize()  # the kernel runs on the engine's stream, not torch's
        return t.data_ptr(), o.shape[0], t

    def get_obs(self) -> np.ndarray:
        """InferUnit::GetObs(state, prevActions) (InferUnit.cpp:45-50): the OBS builder's rows for the engine's current states,
        [arenas * players, obs_size]; the previous action is the one the engine last applied (Match keeps it, G/Match.cpp)."""
        obs, _, _ = self.engine.read_outputs()
        return obs

    # ---- policy -----------------------------------------------------------------------------------------------------------
    def infer_policy_indices(self, deterministic: bool, temperature: float = 1.0, obs: Optional[np.ndarray] = None) -> np.ndarray:
        """The action INDICES DiscretePolicy::GetAction returns (argmax, or a multinomial draw), [rows] int32."""
        self._assert_type(True)
        ptr, n, keep = self._obs_dev(obs)
        c = self._collector(deterministic, temperature)
        self._counter += 1
        with self.torch.cuda.device(self.device):
            c.infer(ptr, n, self._counter, action_ptr=self._act.data_ptr())
            self.engine.sync()
        del keep
        return self._act[:n].cpu().numpy()

    def infer_policy_all(self, deterministic: bool, temperature: float = 1.0, obs: Optional[np.ndarray] = None) -> np.ndarray:
        """InferUnit::InferPolicyAll (InferUnit.cpp:55-72): parsed actions, [rows, 8] = throttle, steer, pitch, yaw, roll, jump,
        boost, handbrake (DiscreteAction lookup of the chosen indices)."""
        return self.table[self.infer_policy_indices(deterministic, temperature, obs)]

    def infer_policy_single(self, row: int, deterministic: bool, temperature: float = 1.0, obs: Optional[np.ndarray] = None) -> np.ndarray:
        """InferUnit::InferPolicySingle (InferUnit.cpp:74-98) for one (arena, player) row: [8]."""
        o = self.get_obs()[row] if obs is None else np.asarray(obs, dtype=np.float32).reshape(-1)
        return self.infer_policy_all(deterministic, temperature, o[None, :])[0]

    def infer_policy_single_distrib(self, row: int, temperature: float = 1.0, obs: Optional[np.ndarray] = None) -> np.ndarray:
        """InferUnit::InferPolicySingleDistrib (InferUnit.cpp:100-112): GetActionProbs = clamp(softmax(logits / temperature),
        1e-11, 1) (DiscretePolicy.cpp:44-49), [num_actions]."""
        self._assert_type(True)
        torch = self.torch
        from . import gemm

        o = self.get_obs()[row] if obs is None else np.asarray(obs, dtype=np.float32).reshape(-1)
        if self._mlp is None:
            self._mlp = gemm.MLPTF32(self.seq)
        x = torch.zeros((4, self.engine.obs_size), dtype=torch.float32, device=self.device)  # the GEMM wants rows % 4 == 0
        x[0] = torch.from_numpy(np.ascontiguousarray(o)).to(self.device)
        with torch.no_grad():
            logits = self._mlp(x)[0]
            probs = torch.softmax(logits / float(temperature), dim=-1).clamp(ACTION_MIN_PROB, 1)
        return probs.cpu().numpy()

    # ---- critic -----------------------------------------------------------------------------------------------------------
    def infer_critic_all(self, obs: Optional[np.ndarray] = None) -> np.ndarray:
        """InferUnit::InferCriticAll (InferUnit.cpp:114-122): ValueEstimator::Forward per row, [rows] float32."""
        self._assert_type(False)
        ptr, n, keep = self._obs_dev(obs)
        c = self._collector(True, 1.0)
        with self.torch.cuda.device(self.device):
            c.infer(ptr, n, 0, value_ptr=self._val.data_ptr())
            self.engine.sync()
        del keep
        return self._val[:n].cpu().numpy()

    def infer_critic_single(self, row: int, obs: Optional[np.ndarray] = None) -> float:
        """InferUnit::InferCriticSingle (InferUnit.cpp:124-132)."""
        o = self.get_obs()[row] if obs is None else np.asarray(obs, dtype=np.float32).reshape(-1)
        return float(self.infer_critic_all(o[None, :])[0])

    def close(self):
        for c in self._collectors.values():
            c.close()
        self._collectors.clear()
