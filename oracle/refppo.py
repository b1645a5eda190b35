"""ctypes binding of oracle/_ref/librlref_ppo.so: the UNMODIFIED reference's ComputeGAE / DiscretePolicy / ValueEstimator
compiled against the pip libtorch (oracle/Makefile target ref_ppo, oracle/ref_ppo_harness.cpp).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "librlref_ppo.so")
_lib = None


def available() -> bool:
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        import torch  # noqa: F401  (loads libtorch / libc10 into the process first)

        _lib = C.CDLL(_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def compute_gae(rews, dones, truncated, values, gamma, lam, return_std, clip_range):
    """TorchFuncs::ComputeGAE (TorchFuncs.cpp:5-52): returns (advantages, value_targets, returns)."""
    r = np.ascontiguousarray(rews, np.float32); d = np.ascontiguousarray(dones, np.float32)
    t = np.ascontiguousarray(truncated, np.float32); v = np.ascontiguousarray(values, np.float32)
    n = len(r)
    assert len(v) == n + 1
    adv = np.empty(n, np.float32); tgt = np.empty(n, np.float32); ret = np.empty(n, np.float32)
    rc = lib().ref_ppo_gae(n, _p(r), _p(d), _p(t), _p(v), C.c_float(gamma), C.c_float(lam), C.c_float(return_std), C.c_float(clip_range),
                           _p(adv), _p(tgt), _p(ret))
    assert rc == 0, rc
    return adv, tgt, ret


def _layer_ptrs(layers):
    Ws = [np.ascontiguousarray(W, np.float32) for W, _ in layers]
    bs = [np.ascontiguousarray(b, np.float32) for _, b in layers]
    pw = (C.c_void_p * len(Ws))(*[w.ctypes.data for w in Ws])
    pb = (C.c_void_p * len(bs))(*[b.ctypes.data for b in bs])
    return Ws, bs, pw, pb


def policy(layers, obs, acts, temperature=1.0):
    """DiscretePolicy (DiscretePolicy.cpp:28-75) with the given nn.Linear tensors: (probs [rows, A], argmax [rows],
    log-prob of acts [rows], mean entropy)."""
    obs = np.ascontiguousarray(obs, np.float32); acts = np.ascontiguousarray(acts, np.int64)
    rows, in_dim = obs.shape
    hidden = np.array([W.shape[0] for W, _ in layers[:-1]], np.int32)
    n_act = layers[-1][0].shape[0]
    keep = _layer_ptrs(layers)
    probs = np.empty((rows, n_act), np.float32); arg = np.empty(rows, np.int64); lp = np.empty(rows, np.float32); ent = C.c_float(0)
    rc = lib().ref_ppo_policy(rows, in_dim, n_act, len(hidden), _p(hidden), keep[2], keep[3], C.c_float(temperature), _p(obs), _p(acts),
                              _p(probs), _p(arg), _p(lp), C.byref(ent))
    assert rc == 0, rc
    return probs, arg, lp, float(ent.value)


def critic(layers, obs):
    """ValueEstimator::Forward (ValueEstimator.cpp:6-27, .h:15)."""
    obs = np.ascontiguousarray(obs, np.float32)
    rows, in_dim = obs.shape
    hidden = np.array([W.shape[0] for W, _ in layers[:-1]], np.int32)
    keep = _layer_ptrs(layers)
    out = np.empty(rows, np.float32)
    rc = lib().ref_ppo_critic(rows, in_dim, len(hidden), _p(hidden), keep[2], keep[3], _p(obs), _p(out))
    assert rc == 0, rc
    return out


def buffer_fifo(max_size, width, sizes, starts, batch_size=0):
    """ExperienceBuffer(maxSize).SubmitExperience x len(sizes) with counting data (ref_ppo_harness.cpp ref_ppo_buffer): returns
    (curSize, states [maxSize, width], actions [maxSize], advantages [maxSize], number of shuffled full batches)."""
    sizes = np.ascontiguousarray(sizes, np.int32); starts = np.ascontiguousarray(starts, np.float32)
    st = np.empty((max_size, width), np.float32); ac = np.empty(max_size, np.float32); ad = np.empty(max_size, np.float32)
    cur = C.c_int64(0); nb = C.c_int32(0)
    rc = lib().ref_ppo_buffer(C.c_int64(max_size), width, len(sizes), _p(sizes), _p(starts), _p(st), _p(ac), _p(ad), C.byref(cur),
                              C.c_int64(batch_size), C.byref(nb))
    assert rc == 0, rc
    return int(cur.value), st, ac, ad, int(nb.value)
