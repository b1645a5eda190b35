"""ppo_oracle.py — TEST INFRASTRUCTURE ONLY (oracle).

Plain numpy-float32 restatement of the learner-side arithmetic of the collection path
(paths under /root/reference/RLGymPPO_CPP/src/):

* ``mlp_forward`` / ``policy_probs``   — DiscretePolicy / ValueEstimator: Linear+ReLU stack, final Linear,
  softmax(logits / temperature), clamp(1e-11, 1)           (private/RLGymPPO_CPP/PPO/DiscretePolicy.cpp:7-42, .h:21-33,
  PPO/ValueEstimator.cpp:6-27)
* ``concat_reference_order``           — ThreadAgentManager::CollectTimesteps: per-player trajectories back to back, last
  step marked truncated unless done                         (private/RLGymPPO_CPP/Threading/ThreadAgentManager.cpp:36-66)
* ``compute_gae``                      — TorchFuncs::ComputeGAE, scalar reverse scan over the WHOLE concatenation
                                                            (private/RLGymPPO_CPP/Util/TorchFuncs.cpp:5-52)
* ``ExperienceBufferOracle``           — ExperienceBuffer::SubmitExperience FIFO (PPO/ExperienceBuffer.cpp:12-70)

Pinning: the reference has no tests or golden vectors for this code, so it is pinned against the reference's own
binaries instead — TorchFuncs.cpp, DiscretePolicy.cpp and ValueEstimator.cpp compiled UNMODIFIED against the pip libtorch
(oracle/Makefile ``ref_ppo`` -> oracle/_ref/librlref_ppo.so, bound by oracle/refppo.py); their outputs on seeded inputs are
tests/golden/ppo_reference.npz.  tests/test_ppo_oracle.py: ``compute_gae`` bit-exact, probabilities / log-probs / entropy
/ critic values within 2e-6 .. 2e-5 (ATen sums in another order).  The buffer and the concatenation (first-party scalar
code) additionally have hand-derived known-answer cases.  Not pinned: torch.multinomial's random stream.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
ACTION_MIN_PROB = f32(1e-11)


def mlp_forward(layers, x):
    """layers: [(W [out,in], b [out])...]; ReLU after every layer but the last. float32 throughout (float64 accumulate
    inside numpy's matmul is avoided by using float32 arrays; the summation ORDER is unspecified in libtorch too)."""
    h = np.asarray(x, dtype=np.float32)
    for i, (W, b) in enumerate(layers):
        h = h @ np.asarray(W, dtype=np.float32).T + np.asarray(b, dtype=np.float32)
        if i + 1 < len(layers):
            h = np.maximum(h, f32(0))
    return h.astype(np.float32)


def policy_probs(logits, temperature=1.0):
    """DiscretePolicy::GetOutput + GetActionProbs (DiscretePolicy.h:21-26, .cpp:37-42)."""
    z = (np.asarray(logits, dtype=np.float32) / f32(temperature)).astype(np.float32)
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z, dtype=np.float32)
    p = (e / e.sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    return np.clip(p, ACTION_MIN_PROB, f32(1))


def concat_reference_order(tmajor, done_ta, P):
    """tmajor: dict name -> [T, N, ...] arrays (N = A*P rows, row n = arena*P + player); done_ta [T, A].
    Returns dict of [N*T, ...] arrays in the reference's concatenation order (row i = n*T + t) plus 'dones' and
    'truncateds' (float32) exactly as CollectTimesteps leaves them."""
    T, A = done_ta.shape
    N = A * P
    out = {k: np.ascontiguousarray(np.swapaxes(v, 0, 1)).reshape((N * T,) + v.shape[2:]) for k, v in tmajor.items()}
    d = np.repeat(done_ta.astype(np.float32), P, axis=1)  # [T, N]
    dn = np.ascontiguousarray(d.T)                         # [N, T]
    tr = np.zeros_like(dn)
    tr[:, T - 1] = (dn[:, T - 1] == 0).astype(np.float32)  # ThreadAgentManager.cpp:53
    out["dones"] = dn.reshape(N * T)
    out["truncateds"] = tr.reshape(N * T)
    return out


def compute_gae(rews, dones, truncated, values, gamma, lam, return_std, clip_range):
    """TorchFuncs.cpp:5-52, line for line, float32 scalars. values has len(rews)+1 entries."""
    rews = np.asarray(rews, dtype=np.float32)
    dones = np.asarray(dones, dtype=np.float32)
    truncated = np.asarray(truncated, dtype=np.float32)
    values = np.asarray(values, dtype=np.float32)
    n = len(rews)
    assert len(values) == n + 1
    next_values = values[1:]
    gamma, lam, return_std, clip_range = f32(gamma), f32(lam), f32(return_std), f32(clip_range)
    with np.errstate(divide="ignore", invalid="ignore"):
        return_scale = f32(1) / return_std
    if np.isnan(return_scale):
        return_scale = f32(0)
    last_gae = f32(0)
    last_return = f32(0)
    adv = np.zeros(n, dtype=np.float32)
    returns = np.zeros(n, dtype=np.float32)
    for step in range(n - 1, -1, -1):
        done = f32(1) - dones[step]
        trunc = f32(1) - truncated[step]
        if return_std != 0:
            norm_rew = f32(rews[step] * return_scale)
            if clip_range > 0:
                norm_rew = min(max(norm_rew, -clip_range), clip_range)
        else:
            norm_rew = rews[step]
        pred_ret = f32(norm_rew + f32(f32(gamma * next_values[step]) * done))
        delta = f32(pred_ret - values[step])
        ret = f32(rews[step] + f32(f32(f32(last_return * gamma) * done) * trunc))
        returns[step] = ret
        last_return = ret
        last_gae = f32(delta + f32(f32(f32(f32(gamma * lam) * done) * trunc) * last_gae))
        adv[step] = last_gae
    value_targets = (values[:-1] + adv).astype(np.float32)
    return adv, value_targets, returns


def gae_reference_order(reward_tn, done_ta, value_t1n, P, gamma, lam, return_std, clip_range):
    """End to end: T-major collector arrays -> reference concatenation -> ComputeGAE -> back to T-major [T, N]."""
    T, N = reward_tn.shape
    cat = concat_reference_order({"rewards": reward_tn, "values": value_t1n[:T]}, done_ta, P)
    values = np.concatenate([cat["values"], value_t1n[T, N - 1:N]]).astype(np.float32)  # + V(nextStates[count-1]) (Learner.cpp:618-640)
    adv, tgt, ret = compute_gae(cat["rewards"], cat["dones"], cat["truncateds"], values, gamma, lam, return_std, clip_range)
    back = lambda a: np.ascontiguousarray(a.reshape(N, T).T)
    return back(adv), back(tgt), back(ret)


class ExperienceBufferOracle:
    """ExperienceBuffer::SubmitExperience (ExperienceBuffer.cpp:12-70): FIFO of max_size rows over several tensors."""

    def __init__(self, max_size):
        self.max_size = int(max_size)
        self.cur = 0
        self.data = {}

    def submit(self, tensors):
        empty = self.cur == 0
        first = None
        for k, add in tensors.items():
            add = np.asarray(add)
            n = add.shape[0]
            if first is None:
                first = n
            if n > self.max_size:
                add = add[n - self.max_size:]
                n = self.max_size
            overflow = max(self.cur + n - self.max_size, 0)
            start, end = self.cur - overflow, self.cur + n - overflow
            if empty:
                self.data[k] = np.full((self.max_size,) + add.shape[1:], np.nan, dtype=np.float64).astype(add.dtype) if add.dtype.kind == "f" else np.zeros((self.max_size,) + add.shape[1:], dtype=add.dtype)
            elif overflow > 0:
                self.data[k][:self.cur - overflow] = self.data[k][overflow:self.cur].copy()
            self.data[k][start:end] = add
        self.cur = min(self.cur + first, self.max_size)
