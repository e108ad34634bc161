"""GPU: kernel (3) on the tensor cores -- the tcgen05/TMEM/TMA GEMM behind the
embedder layers -- against float64 arithmetic on the same bf16 operands, and the
bf16 training step against the fp32 step.  Stated bf16 tolerance: operands are
rounded to bf16 (2^-9 relative), products are exact and accumulate in fp32."""
import numpy as np
import pytest
import torch

from abnet3_b200 import ops
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.model import SiameseNetwork, SiameseMultitaskNetwork

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _operands(M, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.zeros((M, ops.pad8(K)), dtype=torch.bfloat16)
    B = torch.zeros((N, ops.pad8(K)), dtype=torch.bfloat16)
    A[:, :K] = (torch.randn(M, K, generator=g) / np.sqrt(K)).bfloat16()
    B[:, :K] = torch.randn(N, K, generator=g).bfloat16()
    A[:, K:] = 7.0            # padding columns must never be read (TMA bounds = true K)
    B[:, K:] = 7.0
    ref = A[:, :K].double() @ B[:, :K].double().T
    return A.to(DEV), B.to(DEV), ref


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 64, 64), (256, 500, 280), (1000, 100, 500),
                                   (16384, 500, 500), (300, 100, 24), (130, 36, 8), (2048, 280, 500)])
def test_tcgen05_gemm_store(M, N, K):
    A, B, ref = _operands(M, N, K, M + N + K)
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm_bf16_tn(A, B, M, N, K, ops.EPI_STORE, out_f32=out)
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), ref.numpy(), rtol=2e-4, atol=2e-4)


def test_tcgen05_gemm_bias_act_and_all_outputs():
    M, N, K = 700, 500, 280
    A, B, ref = _operands(M, N, K, 3)
    bias = torch.randn(N) * 0.1
    for act, fn in (("sigmoid", torch.sigmoid), ("tanh", torch.tanh), ("relu", torch.relu),
                    ("none", lambda v: v)):
        want = fn(ref + bias.double())
        o32 = torch.zeros((M, N), device=DEV)
        o16 = torch.zeros((M, ops.pad8(N)), dtype=torch.bfloat16, device=DEV)
        oT = torch.zeros((N, ops.pad8(M)), dtype=torch.bfloat16, device=DEV)
        ops.gemm_bf16_tn(A, B, M, N, K, ops.EPI_BIAS_ACT, bias.to(DEV), act, out_f32=o32,
                         out_bf16=o16, outT_bf16=oT)
        torch.cuda.synchronize()
        np.testing.assert_allclose(o32.cpu().numpy(), want.numpy(), rtol=3e-4, atol=3e-4)
        assert torch.equal(o16[:, :N].cpu(), o32.cpu().bfloat16())            # same values, rounded once
        assert torch.equal(oT[:, :M].cpu(), o32.cpu().bfloat16().T)
        assert float(o16[:, N:].abs().max()) == 0 and float(oT[:, M:].abs().max()) == 0


def test_tcgen05_gemm_dgrad_epilogue_applies_act_derivative():
    """epilogue 3: dz_below = (dz W) * act'(y_below), plus its transpose and column sums."""
    M, N, K = 700, 500, 100
    A, B, ref = _operands(M, N, K, 5)
    g = torch.Generator().manual_seed(1)
    for act, dfn in (("sigmoid", lambda y: y * (1 - y)), ("tanh", lambda y: 1 - y * y),
                     ("relu", lambda y: (y > 0).double())):
        yprev = torch.zeros((M, ops.pad8(N)), dtype=torch.bfloat16)
        raw = torch.randn(M, N, generator=g)
        yprev[:, :N] = {"sigmoid": torch.sigmoid, "tanh": torch.tanh, "relu": torch.relu}[act](raw).bfloat16()
        want = ref * dfn(yprev[:, :N].double())
        o16 = torch.zeros((M, ops.pad8(N)), dtype=torch.bfloat16, device=DEV)
        oT = torch.zeros((N, ops.pad8(M)), dtype=torch.bfloat16, device=DEV)
        db = torch.full((N,), 0.5, device=DEV)
        ops.gemm_bf16_tn(A, B, M, N, K, ops.EPI_DGRAD_ACT, act=act, yprev=yprev.to(DEV),
                         out_bf16=o16, outT_bf16=oT, db=db)
        torch.cuda.synchronize()
        np.testing.assert_allclose(o16[:, :N].float().cpu().numpy(), want.numpy(), rtol=1e-2, atol=2e-3)
        assert torch.equal(oT[:, :M].cpu(), o16[:, :N].cpu().T)
        np.testing.assert_allclose(db.cpu().numpy(), 0.5 + want.sum(0).numpy(), rtol=2e-3, atol=2e-2)


def test_tcgen05_gemm_split_k_atomic_accumulates():
    M, N, K = 500, 280, 16384          # the wgrad shape: contraction over the batch
    A, B, ref = _operands(M, N, K, 9)
    out = torch.ones((M, N), device=DEV)
    ops.gemm_bf16_tn(A, B, M, N, K, ops.EPI_ATOMIC, out_f32=out, split_k=8)
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), 1.0 + ref.numpy(), rtol=3e-4, atol=3e-4)


def test_cast_and_act_backward_bf16():
    torch.manual_seed(0)
    x = torch.randn(333, 280, device=DEV)
    xb = torch.zeros((333, 280), dtype=torch.bfloat16, device=DEV)
    xT = torch.zeros((280, ops.pad8(333)), dtype=torch.bfloat16, device=DEV)
    ops.cast_bf16(x, xb, xT)
    assert torch.equal(xb, x.bfloat16()) and torch.equal(xT[:, :333], x.bfloat16().T)
    y = torch.sigmoid(torch.randn(333, 100, device=DEV))
    dy = torch.randn(333, 100, device=DEV)
    dz = torch.zeros((333, ops.pad8(100)), dtype=torch.bfloat16, device=DEV)
    dzT = torch.zeros((100, ops.pad8(333)), dtype=torch.bfloat16, device=DEV)
    db = torch.zeros(100, device=DEV)
    ops.act_backward_bf16(y, dy, "sigmoid", dz, dzT, db)
    want = dy * y * (1 - y)
    assert torch.equal(dz[:, :100], want.bfloat16()) and torch.equal(dzT[:, :333], want.bfloat16().T)
    np.testing.assert_allclose(db.cpu().numpy(), want.sum(0).cpu().numpy(), rtol=1e-4, atol=1e-5)


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def test_bf16_training_step_tracks_fp32_step():
    """Stated tolerance of the tensor-core path: embeddings / loss within 1e-2
    relative, per-tensor gradients (hence the SGD update) within 6e-2 relative in
    norm (bf16 rounding of activations and of dz compounds over the 4 layers)."""
    torch.manual_seed(0)
    cfg = dict(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
               activation_layer="sigmoid")
    n32 = SiameseNetwork(**cfg).to(DEV)
    n16 = SiameseNetwork(precision="bf16", **cfg).to(DEV)
    n16.load_state_dict(n32.state_dict())
    s32 = SiameseTrainStep(n32, ("coscos2", 0.0, False), "sgd", lr=0.05, momentum=0.0)
    s16 = SiameseTrainStep(n16, ("coscos2", 0.0, False), "sgd", lr=0.05, momentum=0.0)
    before = {k: v.clone() for k, v in n32.state_dict().items()}
    n = 4096
    x = torch.randn(2 * n, 280, device=DEV)
    y = torch.where(torch.rand(n, device=DEV) < 0.5, 1.0, -1.0)
    l32 = float(s32.step(x, n, y).item())
    l16 = float(s16.step(x, n, y).item())
    assert abs(l16 - l32) <= 1e-2 * abs(l32)
    assert _rel(s16.acts[-1], s32.acts[-1]) < 1e-2
    for (k, a), (_, b) in zip(n16.state_dict().items(), n32.state_dict().items()):
        upd16, upd32 = a - before[k], b - before[k]
        assert _rel(upd16, upd32) < 6e-2, k


def test_bf16_multitask_step_tracks_fp32_step():
    torch.manual_seed(2)
    cfg = dict(input_dim=280, num_hidden_layers_shared=2, num_hidden_layers_spk=1,
               num_hidden_layers_phn=1, hidden_dim=500, output_dim=100, p_dropout=0.0,
               activation_layer="sigmoid")
    n32 = SiameseMultitaskNetwork(**cfg).to(DEV)
    n16 = SiameseMultitaskNetwork(precision="bf16", **cfg).to(DEV)
    n16.load_state_dict(n32.state_dict())
    spec = (("coscos2", 0.0, False), ("coscos2", 0.0, False), 0.3)
    s32 = SiameseTrainStep(n32, spec, "sgd", lr=0.05, momentum=0.0)
    s16 = SiameseTrainStep(n16, spec, "sgd", lr=0.05, momentum=0.0)
    before = {k: v.clone() for k, v in n32.state_dict().items()}
    n = 1024
    x = torch.randn(2 * n, 280, device=DEV)
    ys = torch.where(torch.rand(n, device=DEV) < 0.5, 1.0, -1.0)
    yp = torch.where(torch.rand(n, device=DEV) < 0.5, 1.0, -1.0)
    l32 = float(s32.step(x, n, ys, yp).item())
    l16 = float(s16.step(x, n, ys, yp).item())
    assert abs(l16 - l32) <= 1e-2 * abs(l32)
    for (k, a), (_, b) in zip(n16.state_dict().items(), n32.state_dict().items()):
        upd16, upd32 = a - before[k], b - before[k]
        if float(upd32.norm()) == 0:
            assert float(upd16.norm()) == 0
        else:
            assert _rel(upd16, upd32) < 6e-2, k
