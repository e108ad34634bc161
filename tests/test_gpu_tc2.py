"""GPU: the persistent grouped tcgen05 GEMM (abn_gemm_bf16_group) -- forward, dgrad and
wgrad forms on natural row-major bf16 operands (K-major and MN-major UMMA descriptors),
against float64 arithmetic on the same bf16 values."""
import numpy as np
import pytest
import torch

from abnet3_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bf(rows, cols, seed, scale=1.0, pad_val=7.0):
    """bf16 [rows, pad8(cols + 1)] with the padding poisoned: it must never be read."""
    g = torch.Generator().manual_seed(seed)
    t = torch.full((rows, ops.pad8(cols + 1)), pad_val, dtype=torch.bfloat16)
    t[:, :cols] = (torch.randn(rows, cols, generator=g) * scale).bfloat16()
    return t


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 128, 64), (256, 500, 280), (1000, 100, 500),
                                   (16384, 500, 500), (300, 100, 24), (130, 36, 8), (2048, 280, 500),
                                   (333, 512, 128)])
def test_forward_form_k_major(M, N, K):
    A, B = _bf(M, K, 1, K ** -0.5), _bf(N, K, 2)
    ref = A[:, :K].double() @ B[:, :K].double().T
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm_group([ops.gemm_problem(A.to(DEV), B.to(DEV), M, N, K, ops.GE_BIAS_ACT, out)])
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=2e-4, atol=2e-4)


def test_forward_bias_activation_bf16_and_ones_column():
    M, N, K = 700, 500, 280
    A, B = _bf(M, K, 3, K ** -0.5), _bf(N, K, 4)
    ref = A[:, :K].double() @ B[:, :K].double().T
    bias = torch.randn(N) * 0.1
    for act, fn in (("sigmoid", torch.sigmoid), ("tanh", torch.tanh), ("relu", torch.relu),
                    ("none", lambda v: v)):
        want = fn(ref + bias.double())
        o16 = torch.zeros((M, ops.pad8(N + 1)), dtype=torch.bfloat16, device=DEV)
        ops.gemm_group([ops.gemm_problem(A.to(DEV), B.to(DEV), M, N, K, ops.GE_BIAS_ACT, o16, act=act,
                                         bias=bias.to(DEV), ones_col=True)])
        torch.cuda.synchronize()
        np.testing.assert_allclose(o16[:, :N].float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=4e-3)
        assert bool((o16[:, N] == 1).all())
        o32 = torch.zeros((M, N), device=DEV)
        ops.gemm_group([ops.gemm_problem(A.to(DEV), B.to(DEV), M, N, K, ops.GE_BIAS_ACT, o32, act=act,
                                         bias=bias.to(DEV))])
        torch.cuda.synchronize()
        np.testing.assert_allclose(o32.cpu().numpy(), want.numpy(), rtol=2e-3, atol=1e-3)


@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (700, 500, 100), (16384, 280, 500), (130, 500, 200)])
def test_dgrad_form_b_mn_major_with_act_derivative(M, N, K):
    """dz_below = (dz W) * act'(y_below): A = dz [M, K], B = W [K, N] as stored (N contiguous)."""
    A = _bf(M, K, 5, K ** -0.5)
    W = _bf(K, N, 6)                       # [n_out = K, n_in = N]
    ref = A[:, :K].double() @ W[:, :N].double()
    g = torch.Generator().manual_seed(7)
    for act, dfn in (("sigmoid", lambda y: y * (1 - y)), ("none", lambda y: torch.ones_like(y))):
        yprev = _bf(M, N, 8)
        yprev[:, :N] = torch.sigmoid(torch.randn(M, N, generator=g)).bfloat16()
        want = ref * dfn(yprev[:, :N].double())
        o16 = torch.zeros((M, ops.pad8(N + 1)), dtype=torch.bfloat16, device=DEV)
        ops.gemm_group([ops.gemm_problem(A.to(DEV), W.to(DEV), M, N, K, ops.GE_DACT, o16, b_mn=True,
                                         act=act, yprev=yprev.to(DEV))])
        torch.cuda.synchronize()
        np.testing.assert_allclose(o16[:, :N].float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=3e-3)


@pytest.mark.parametrize("rows,n_out,n_in,split", [(128, 128, 64, 1), (1000, 500, 280, 3),
                                                    (16384, 500, 500, 9), (16384, 100, 500, 16),
                                                    (777, 36, 24, 2)])
def test_wgrad_form_both_mn_major_split_k_and_bias_gradient(rows, n_out, n_in, split):
    """dW += dz^T x and db += colsum(dz) through the ones column of x."""
    dz = _bf(rows, n_out, 9, rows ** -0.5)
    x = _bf(rows, n_in, 10)
    x[:, n_in] = 1.0
    want_w = dz[:, :n_out].double().T @ x[:, :n_in].double()
    want_b = dz[:, :n_out].double().sum(0)
    gw = torch.ones((n_out, n_in), device=DEV)
    gb = torch.full((n_out,), 2.0, device=DEV)
    ops.gemm_group([ops.gemm_problem(dz.to(DEV), x.to(DEV), n_out, n_in, rows, ops.GE_ATOMIC, gw,
                                     a_mn=True, b_mn=True, split_k=split, ones_out=gb)])
    torch.cuda.synchronize()
    np.testing.assert_allclose(gw.cpu().numpy(), 1.0 + want_w.numpy(), rtol=3e-4, atol=3e-4)
    np.testing.assert_allclose(gb.cpu().numpy(), 2.0 + want_b.numpy(), rtol=3e-4, atol=3e-4)


def test_group_of_four_problems_in_one_launch():
    rows = 4096
    shapes = [(500, 280), (500, 500), (500, 500), (100, 500)]
    probs, wants, outs = [], [], []
    for i, (n_out, n_in) in enumerate(shapes):
        dz = _bf(rows, n_out, 20 + i, rows ** -0.5)
        x = _bf(rows, n_in, 30 + i)
        x[:, n_in] = 1.0
        gw = torch.zeros((n_out, n_in), device=DEV)
        gb = torch.zeros((n_out,), device=DEV)
        probs.append(ops.gemm_problem(dz.to(DEV), x.to(DEV), n_out, n_in, rows, ops.GE_ATOMIC, gw,
                                      a_mn=True, b_mn=True, split_k=4, ones_out=gb))
        wants.append((dz[:, :n_out].double().T @ x[:, :n_in].double(), dz[:, :n_out].double().sum(0)))
        outs.append((gw, gb))
    ops.gemm_group(probs)
    torch.cuda.synchronize()
    for (gw, gb), (ww, wb) in zip(outs, wants):
        np.testing.assert_allclose(gw.cpu().numpy(), ww.numpy(), rtol=3e-4, atol=3e-4)
        np.testing.assert_allclose(gb.cpu().numpy(), wb.numpy(), rtol=3e-4, atol=3e-4)
