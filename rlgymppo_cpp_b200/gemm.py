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
         split_k: int = 1, mask: torch.Tensor = None, out_t: torch.Tensor = None) -> torch.Tensor:
    """out[M, N] (+)= a[M, K] @ b[N, K].T (+ bias) (ReLU) (zeroed where mask <= 0), TF32 tensor cores, fp32 accumulate;
    out_t [N, M] optionally receives the transposed result as well."""
    assert a.is_cuda and b.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1] and a.stride(1) == 1 and b.stride(1) == 1
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        assert not (accumulate or atomic)
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype == torch.float32
    if mask is not None:
        assert mask.shape == (M, N) and mask.stride(1) == 1 and mask.dtype == torch.float32
    if out_t is not None:
        assert out_t.shape == (N, M) and out_t.stride(1) == 1 and out_t.dtype == torch.float32
    flags = (RELU if relu else 0) | (ACCUMULATE if accumulate else 0) | (ATOMIC if atomic else 0)
    L = load_library()
    stream = torch.cuda.current_stream(a.device).cuda_stream
    with torch.cuda.device(a.device):
        _check(L.rlg_gemm_tf32_fused(M, N, K, _ptr(a), a.stride(0), _ptr(b), b.stride(0), _ptr(out), out.stride(0),
                                     _ptr(bias) if bias is not None else None, flags, int(split_k),
                                     _ptr(mask) if mask is not None else None, mask.stride(0) if mask is not None else 0,
                                     _ptr(out_t) if out_t is not None else None, out_t.stride(0) if out_t is not None else 0, C.c_void_p(stream)))
    return out


def _pad4(n: int) -> int:
    return (n + 3) & ~3


class _MLPFunction(torch.autograd.Function):
    """The whole Sequential(Linear, ReLU, ..., Linear) as one autograd node: 3 GEMMs per layer on csrc/gemm.cu.

    forward : y_l = gemm(x_l, W_l) + b_l (ReLU fused for hidden layers); the epilogue also writes y_l^T
    backward: dW_l = gemm(dy_l^T, x_l^T) split over K = rows (atomic), db_l = row sums of dy_l^T,
              dx_l = gemm(dy_l, W_l^T) masked by x_l > 0 (ReLU backward fused), epilogue also writes dx_l^T
    so no tensor is transposed by a separate pass except the network input (once) and the loss gradient (a few columns).
    Feature dimensions are padded to multiples of 4 floats (89 -> 92, 90 -> 92, 1 -> 4) with zero rows / columns."""

    @staticmethod
    def forward(ctx, x, n_layers, *params):
        weights, biases = params[:n_layers], params[n_layers:]
        rows = x.shape[0]
        dev = x.device
        xs, xts, wps = [], [], []
        cur = x
        cur_t = x.t().contiguous()
        for l in range(n_layers):
            W, b = weights[l], biases[l]
            out_f, in_f = W.shape
            in_p, out_p = cur.shape[1], _pad4(out_f)
            wp = W if (in_p == in_f and out_p == out_f) else torch.nn.functional.pad(W, (0, in_p - in_f, 0, out_p - out_f))
            bp = b if out_p == out_f else torch.nn.functional.pad(b, (0, out_p - out_f))
            wp = wp.contiguous()
            last = l == n_layers - 1
            y_t = None if last else torch.empty((out_p, rows), dtype=torch.float32, device=dev)
            y = gemm(cur, wp, bias=bp.contiguous(), relu=not last, out_t=y_t)
            xs.append(cur); xts.append(cur_t); wps.append(wp)
            cur, cur_t = y, y_t
        ctx.n_layers = n_layers
        ctx.shapes = [tuple(W.shape) for W in weights]
        ctx.save_for_backward(*xs, *xts, *wps)
        return cur

    @staticmethod
    def backward(ctx, dy):
        n = ctx.n_layers
        saved = ctx.saved_tensors
        xs, xts, wps = saved[:n], saved[n:2 * n], saved[2 * n:]
        rows = xs[0].shape[0]
        dev = dy.device
        dy = dy.contiguous()
        dy_t = dy.t().contiguous()
        dws, dbs = [None] * n, [None] * n
        for l in range(n - 1, -1, -1):
            out_f, in_f = ctx.shapes[l]
            wp = wps[l]
            dw = torch.zeros((wp.shape[0], wp.shape[1]), dtype=torch.float32, device=dev)
            split = max(1, min(rows // 256, 148 // max(1, (wp.shape[0] + 127) // 128)))  # about one wave of CTAs
            gemm(dy_t, xts[l], out=dw, atomic=True, split_k=split)
            dws[l] = dw[:out_f, :in_f]
            dbs[l] = dy_t.sum(1)[:out_f]
            if l > 0:
                dx_t = torch.empty((wp.shape[1], rows), dtype=torch.float32, device=dev)
                dy = gemm(dy, wp.t().contiguous(), mask=xs[l], out_t=dx_t)  # xs[l] is layer l-1's ReLU output
                dy_t = dx_t
        return (None, None, *dws, *dbs)


class MLPTF32(torch.nn.Module):
    """The reference's Sequential(Linear, ReLU, ..., Linear) (DiscretePolicy.cpp:11-27, ValueEstimator.cpp:10-24) evaluated
    through _MLPFunction; shares the parameters of an existing torch Sequential (same state_dict keys, same optimizer)."""

    def __init__(self, seq: torch.nn.Sequential):
        super().__init__()
        self.seq = seq
        self.linears = [m for m in seq if isinstance(m, torch.nn.Linear)]

    def forward(self, x):
        rows, in_f = x.shape
        assert rows % 4 == 0, "minibatch rows must be a multiple of 4"
        if in_f % 4:
            x = torch.nn.functional.pad(x, (0, _pad4(in_f) - in_f))
        y = _MLPFunction.apply(x.contiguous(), len(self.linears), *[m.weight for m in self.linears], *[m.bias for m in self.linears])
        return y[:, : self.linears[-1].out_features]
