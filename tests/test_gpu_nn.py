"""GPU parity: kernel (4) pair loss + gradient, kernel (3) embedder layers
(fp32 SIMT path) and the optimizer step, through the C ABI, against the golden
vectors of the LIVE reference and against the oracle restatement (oracle/nets.py)
at the canonical sizes.  Tolerance: 1e-4 relative in fp32 (BASELINE north_star)."""
import os

import numpy as np
import pytest
import torch

from abnet3_b200 import ops
from oracle import nets as onets

pytestmark = pytest.mark.gpu
DEV = "cuda"
RTOL = 1e-4


def _close(a, b, rtol=RTOL, atol=1e-6):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "nets.npz"))


@pytest.mark.parametrize("lname,kind,margin", [("coscos2", "coscos2", 0.5),
                                               ("cosmargin", "cosmargin", 0.5),
                                               ("cosmargin02", "cosmargin", 0.2)])
def test_pair_loss_against_reference_golden(gold, lname, kind, margin):
    e1 = torch.from_numpy(gold["loss/e1"]).to(DEV)
    e2 = torch.from_numpy(gold["loss/e2"]).to(DEV)
    y = torch.from_numpy(gold["loss/y"]).float().to(DEV)
    n = e1.shape[0]
    for avg in (True, False):
        loss, de1, de2 = ops.pair_loss(e1, e2, y, kind, margin, scale=1.0 / n if avg else 1.0)
        tag = "loss/%s_avg%d" % (lname, int(avg))
        _close(loss[0], gold[tag + "/loss"])
        _close(de1, gold[tag + "/de1"], atol=1e-7)
        _close(de2, gold[tag + "/de2"], atol=1e-7)


def test_pair_loss_canonical_size_vs_oracle():
    torch.manual_seed(0)
    n, dim = 8192, 100
    e1 = torch.sigmoid(torch.randn(n, dim))
    e2 = torch.sigmoid(torch.randn(n, dim))
    y = torch.where(torch.rand(n) < 0.5, 1.0, -1.0)
    for kind, fn in (("coscos2", onets.coscos2), ("cosmargin", onets.cosmargin)):
        a = e1.clone().requires_grad_(True)
        b = e2.clone().requires_grad_(True)
        ref = fn(a, b, y, avg=False)
        ref.backward()
        loss, de1, de2 = ops.pair_loss(e1.to(DEV), e2.to(DEV), y.to(DEV), kind)
        _close(loss[0], ref.item())
        _close(de1, a.grad, atol=1e-7)
        _close(de2, b.grad, atol=1e-7)
        # forward-only mode
        loss2, n1, n2 = ops.pair_loss(e1.to(DEV), e2.to(DEV), y.to(DEV), kind, need_grad=False)
        assert n1 is None and n2 is None
        _close(loss2[0], ref.item())


@pytest.mark.parametrize("m,n_in,n_out,act", [(24, 40, 48, "sigmoid"), (1000, 280, 500, "sigmoid"),
                                              (777, 500, 100, "tanh"), (130, 500, 500, "relu"),
                                              (65, 36, 20, "none"), (16384, 280, 500, "sigmoid")])
def test_linear_forward_backward_vs_torch(m, n_in, n_out, act):
    torch.manual_seed(m)
    x = torch.randn(m, n_in)
    W = torch.randn(n_out, n_in) / np.sqrt(n_in)
    b = torch.randn(n_out) * 0.1
    dy = torch.randn(m, n_out) / m
    fn = {"sigmoid": torch.sigmoid, "tanh": torch.tanh, "relu": torch.relu, "none": lambda v: v}[act]
    xr = x.double().requires_grad_(True)
    Wr = W.double().requires_grad_(True)
    br = b.double().requires_grad_(True)
    yr = fn(torch.nn.functional.linear(xr, Wr, br))
    yr.backward(dy.double())
    xd, Wd, bd = x.to(DEV), W.to(DEV), b.to(DEV)
    y = ops.linear_forward(xd, Wd, bd, act)
    _close(y, yr.float(), atol=1e-5)
    dyd = dy.to(DEV).clone()
    dx, dW, db = ops.linear_backward(xd, Wd, y, dyd, act)
    scale = float(xr.grad.abs().max())
    _close(dx, xr.grad.float(), atol=1e-4 * scale)
    _close(dW, Wr.grad.float(), atol=1e-4 * float(Wr.grad.abs().max()))
    _close(db, br.grad.float(), atol=1e-4 * float(br.grad.abs().max()))
    # accumulate mode adds on top
    dy2 = dy.to(DEV).clone()
    ops.linear_backward(xd, Wd, y, dy2, act, need_dx=False, dW=dW, db=db, accumulate=True)
    _close(dW, 2 * Wr.grad.float(), atol=2e-4 * float(Wr.grad.abs().max()))


@pytest.mark.parametrize("kind", ["sgd", "adadelta", "adam"])
def test_optimizer_step_vs_torch_optim(kind):
    torch.manual_seed(1)
    n = 100003
    p0 = torch.randn(n)
    ref = p0.clone().requires_grad_(True)
    opt = {"sgd": lambda: torch.optim.SGD([ref], lr=0.01, momentum=0.9),
           "adadelta": lambda: torch.optim.Adadelta([ref], lr=0.1),
           "adam": lambda: torch.optim.Adam([ref], lr=0.001)}[kind]()
    lr = {"sgd": 0.01, "adadelta": 0.1, "adam": 0.001}[kind]
    p = p0.to(DEV).clone()
    s0 = torch.zeros(n, device=DEV)
    s1 = torch.zeros(n, device=DEV)
    for step in range(1, 6):
        g = torch.randn(n)
        ref.grad = g.clone()
        opt.step()
        ops.optimizer_step(p, g.to(DEV), s0, s1, kind, lr, momentum=0.9, step=step)
    _close(p, ref.detach(), rtol=1e-5, atol=1e-6)
