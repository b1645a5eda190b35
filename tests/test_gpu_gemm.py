"""-m gpu: csrc/gemm.cu (tcgen05 TF32 GEMM of the PPO update) against torch fp32, and the autograd Linear built on it against
torch.nn.Linear's own gradients."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tf32(x):
    import torch

    u = x.contiguous().view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,N,K,kw", [
    (1000, 256, 92, dict(bias=True, relu=True)),      # forward, first layer (obs padded to 92)
    (4096, 92, 256, dict(bias=True)),                 # policy head (90 -> 92)
    (300, 4, 256, dict(bias=True)),                   # critic head (1 -> 4)
    (2048, 256, 256, dict()),                         # dX
    (128, 300, 36, dict()),                           # N > 256 (two column tiles), K tail
    (256, 92, 8192, dict(split_k=32)),                # dW, split over K = rows
    (92, 256, 4096, dict(split_k=7)),
    (5, 16, 4, dict(accumulate=True)),
])
def test_gemm_matches_fp32(M, N, K, kw):
    import torch

    from rlgymppo_cpp_b200 import gemm as G

    assert torch.cuda.is_available()
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g) * 0.3
    bias = torch.randn(N, device="cuda", generator=g) if kw.get("bias") else None
    split = kw.get("split_k", 1)
    base = torch.randn(M, N, device="cuda", generator=g) if (split > 1 or kw.get("accumulate")) else None
    out = base.clone() if base is not None else None
    res = G.gemm(a, b, out=out, bias=bias, relu=kw.get("relu", False), accumulate=kw.get("accumulate", False), atomic=split > 1, split_k=split)
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = False
    # what the tensor core multiplies: fp32 bits read as TF32 (low 13 mantissa bits dropped: the TMA-fed kernel) or operands
    # rounded to nearest TF32 on the way to shared memory (the register-staged kernel), wide accumulation
    trunc = lambda x: (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
    exact_rn = _tf32(a).double() @ _tf32(b).double().t()
    exact_tr = trunc(a).double() @ trunc(b).double().t()
    full = a.double() @ b.double().t()

    def err(ref):
        r = ref + (bias.double() if bias is not None else 0)
        if kw.get("relu"):
            r = r.clamp_min(0)
        if base is not None:
            r = r + base.double()
        return float((res.double() - r).abs().max()) / float(ref.abs().max())

    assert min(err(exact_rn), err(exact_tr)) <= 2e-5, (M, N, K, err(exact_rn), err(exact_tr))
    assert err(full) <= 4e-3, (M, N, K, err(full))


def test_gemm_fused_mask_and_transposed_output():
    import torch

    from rlgymppo_cpp_b200 import gemm as G

    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(1000, 92, device="cuda", generator=g); b = torch.randn(256, 92, device="cuda", generator=g)
    mask = torch.randn(1000, 256, device="cuda", generator=g).clamp_min(0)
    out_t = torch.full((256, 1000), float("nan"), device="cuda")
    res = G.gemm(a, b, mask=mask, out_t=out_t)
    ref = G.gemm(a, b) * (mask > 0)
    assert torch.equal(res, ref) and torch.equal(out_t, res.t())
    # ReLU + transposed output (the forward of a hidden layer)
    out_t = torch.empty((256, 1000), device="cuda")
    res = G.gemm(a, b, relu=True, out_t=out_t)
    assert torch.equal(res, G.gemm(a, b).clamp_min(0)) and torch.equal(out_t, res.t())


def test_gemm_argument_errors():
    import torch

    from rlgymppo_cpp_b200 import gemm as G
    from rlgymppo_cpp_b200.engine import EngineError

    a = torch.zeros(8, 6, device="cuda"); b = torch.zeros(8, 6, device="cuda")
    with pytest.raises(EngineError):
        G.gemm(a, b)  # K not a multiple of 4
    a = torch.zeros(8, 8, device="cuda"); b = torch.zeros(8, 8, device="cuda")
    with pytest.raises(EngineError):
        G.gemm(a, b, out=torch.zeros(8, 8, device="cuda"), split_k=2)  # split-K without the atomic accumulate


def test_mlp_autograd_matches_torch():
    """Forward values and every parameter gradient of the tensor-core MLP against torch's fp32 autograd: as close as torch's own
    TF32 (cuBLAS) path is — the precision the update runs in either way (learner.py sets allow_tf32)."""
    import torch

    from rlgymppo_cpp_b200 import gemm as G, learner as L

    torch.manual_seed(3)
    for out_dim in (90, 1):
        seq = L.make_mlp(89, [256, 256, 256], out_dim).cuda()
        x = torch.randn(2048, 89, device="cuda")
        tgt = torch.randn(2048, out_dim, device="cuda")

        def grads(fwd, tf32):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for p in seq.parameters():
                p.grad = None
            y = fwd(x)
            ((y - tgt) ** 2).mean().backward()
            return y.detach(), [p.grad.clone() for p in seq.parameters()]

        y32, g32 = grads(seq, False)
        ytf, gtf = grads(seq, True)
        ymine, gmine = grads(G.MLPTF32(seq), False)
        torch.backends.cuda.matmul.allow_tf32 = False
        # (the TMA-fed kernel reads fp32 bits as TF32, i.e. truncates where cuBLAS rounds: up to ~2x its error per operand)
        assert float((ymine - y32).abs().max()) <= 8 * float((ytf - y32).abs().max()) + 1e-6
        for p, a, b, c in zip(seq.parameters(), g32, gtf, gmine):
            assert c.shape == a.shape
            e_mine, e_torch = float((c - a).abs().max()), float((b - a).abs().max())
            assert e_mine <= 8 * e_torch + 2e-3 * float(a.abs().max()), (tuple(p.shape), e_mine, e_torch, float(a.abs().max()))
