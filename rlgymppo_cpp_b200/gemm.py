"""ctypes binding of csrc/gemm.cu (`rlg_gemm_tf32`) on torch CUDA tensors + the autograd Linear that routes the PPO
update's dense contractions through it (PPOLearner::Learn's three GEMMs per torch::nn::Linear: forward, dX, dW;
reference /root/reference/RLGymPPO_CPP/src/private/RLGymPPO_CPP/PPO/PPOLearner.cpp:125-290).  No fallback: without the
CUDA library or a device these raise."""
from __future__ import annotations

import ctypes as C

import torch

from .engine import _check, load_library

RELU, ACCUMULATE, ATOMIC = 1, 2, 4


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor = None, bias: torch.Tensor = None, relu=False, accumulate=False, atomic=False,
         split_k: int = 1) -> torch.Tensor:
    """out[M, N] (+)= a[M, K] @ b[N, K].T (+ bias) (ReLU), TF32 tensor cores, fp32 accumulate."""
    assert a.is_cuda and b.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1] and a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        assert not (accumulate or atomic)
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype == torch.float32
    flags = (RELU if relu else 0) | (ACCUMULATE if accumulate else 0) | (ATOMIC if atomic else 0)
    L = load_library()
    stream = torch.cuda.current_stream(a.device).cuda_stream
    with torch.cuda.device(a.device):
        _check(L.rlg_gemm_tf32(M, N, K, _ptr(a), a.stride(0), _ptr(b), b.stride(0), _ptr(out), out.stride(0),
                               _ptr(bias) if bias is not None else None, flags, int(split_k), C.c_void_p(stream)))
    return out


def _pad4(n: int) -> int:
    return (n + 3) & ~3


class _LinearTF32(torch.autograd.Function):
    """y = x W^T + b with the three contractions on csrc/gemm.cu.  x: [rows, in_pad] (columns >= in are zero), returns
    [rows, out_pad]; rows must be a multiple of 4 (it is the K of the weight-gradient GEMM)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        out_f, in_f = weight.shape
        in_p, out_p = x.shape[1], _pad4(out_f)
        wp = weight if (in_p == in_f and out_p == out_f) else torch.nn.functional.pad(weight, (0, in_p - in_f, 0, out_p - out_f))
        bp = bias if out_p == out_f else torch.nn.functional.pad(bias, (0, out_p - out_f))
        y = gemm(x, wp.contiguous(), bias=bp.contiguous(), relu=relu)
        ctx.save_for_backward(x, wp, y if relu else None)
        ctx.relu, ctx.shape = relu, (out_f, in_f)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wp, y = ctx.saved_tensors
        out_f, in_f = ctx.shape
        dy = dy.contiguous()
        if ctx.relu:
            dy = dy * (y > 0)
        rows = x.shape[0]
        dx = gemm(dy, wp.t().contiguous()) if ctx.needs_input_grad[0] else None           # [rows, in_pad]
        dw = torch.zeros((wp.shape[0], wp.shape[1]), dtype=torch.float32, device=x.device)
        split = max(1, min(rows // 256, 148 // max(1, (wp.shape[0] + 127) // 128)))       # ~ one wave of CTAs
        gemm(dy.t().contiguous(), x.t().contiguous(), out=dw, atomic=True, split_k=split)  # [out_pad, in_pad], K = rows
        db = dy.sum(0)
        return dx, dw[:out_f, :in_f], db[:out_f], None


class MLPTF32(torch.nn.Module):
    """The reference's Sequential(Linear, ReLU, ..., Linear) (DiscretePolicy.cpp:11-27, ValueEstimator.cpp:10-24) evaluated
    with _LinearTF32; shares the parameters of an existing torch Sequential (same state_dict keys, same optimizer)."""

    def __init__(self, seq: torch.nn.Sequential):
        super().__init__()
        self.seq = seq
        self.linears = [m for m in seq if isinstance(m, torch.nn.Linear)]

    def forward(self, x):
        rows, in_f = x.shape
        assert rows % 4 == 0, "minibatch rows must be a multiple of 4"
        if in_f % 4:
            x = torch.nn.functional.pad(x, (0, _pad4(in_f) - in_f))
        x = x.contiguous()
        for i, lin in enumerate(self.linears):
            last = i == len(self.linears) - 1
            x = _LinearTF32.apply(x, lin.weight, lin.bias, not last)
        return x[:, : self.linears[-1].out_features]
