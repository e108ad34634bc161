"""GPU: dropout between Linear and activation (abnet3/model.py:111, :136-141; the reference's
default is p = 0.1) inside the kernels' epilogues, checked by MASK REPLAY: the keep masks the
kernels evaluate (a counter-based hash of seed, step, layer, row, col -- abn_dropout_mask exports
them) are handed to the oracle (oracle.nets.siamese_forward_once_dropout restating
Linear -> Dropout -> act with torch autograd), and embeddings, loss and the parameter update must
agree -- for the fp32 SIMT path (1e-4), the bf16 tensor-core chain kernels (bf16 tolerance), and
the autograd surface of SiameseNetwork."""
import numpy as np
import pytest
import torch

from oracle import nets as onets
from abnet3_b200 import ops
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.loss import coscos2
from abnet3_b200.model import SiameseNetwork

pytestmark = pytest.mark.gpu
DEV = "cuda"
P = 0.1
CFG = dict(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=P,
           activation_layer="sigmoid")


def _oracle_step(sd0, x, y, masks, lr):
    sd = {k: v.detach().cpu().double().requires_grad_() for k, v in sd0.items()}
    n = y.shape[0]
    e = onets.siamese_forward_once_dropout(sd, x.cpu().double(), [m.cpu() for m in masks], P)
    loss = onets.coscos2(e[:n], e[n:], y.cpu().double(), avg=False)
    loss.backward()
    return e.detach(), float(loss.detach()), {k: -lr * v.grad for k, v in sd.items()}


@pytest.mark.parametrize("precision,tol_e,tol_l,tol_u", [("fp32", 1e-4, 1e-4, 2e-3), ("bf16", 6e-3, 1e-2, 1e-1)])
def test_training_step_with_dropout_matches_the_oracle_under_the_same_masks(precision, tol_e, tol_l, tol_u):
    torch.manual_seed(0)
    net = SiameseNetwork(precision=precision, **CFG).to(DEV)
    net.train()
    lr = 0.05
    eng = SiameseTrainStep(net, ("coscos2", 0.0, False), "sgd", lr=lr, momentum=0.0)
    n = 2048
    x = torch.randn(2 * n, 280, device=DEV)
    x[n:] = 0.6 * x[:n] + 0.8 * x[n:]
    y = torch.where(torch.rand(n, device=DEV) < 0.5, 1.0, -1.0)
    for step in range(2):                                   # two steps: the masks move on
        sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
        masks = eng.dropout_masks(2 * n)                    # of the step about to run
        kept = float(torch.cat([m.float().flatten() for m in masks]).mean())
        assert abs(kept - (1 - P)) < 3e-3, kept
        loss = float(eng.step(x, n, y).item())
        e_ref, l_ref, upd_ref = _oracle_step(sd0, x, y, masks, lr)
        e_gpu = eng.acts[-1][:2 * n].detach().cpu().double()
        assert float((e_gpu - e_ref).abs().max()) < tol_e
        assert abs(loss - l_ref) <= tol_l * abs(l_ref), (loss, l_ref)
        for k, v in net.state_dict().items():
            upd = (v.detach() - sd0[k]).cpu().double()
            rel = float((upd - upd_ref[k]).norm() / upd_ref[k].norm())
            assert rel < tol_u, (k, rel)
        nxt = eng.dropout_masks(2 * n)
        assert not torch.equal(nxt[0], masks[0])            # a new step, new masks
        assert not torch.equal(masks[1], masks[2])          # layers have their own masks
    # eval mode: dropout is the identity
    net.eval()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    l_eval = float(eng.step(x, n, y, do_training=False).item())
    e = onets.siamese_forward_once(sd, x.cpu())
    l_ref = float(onets.coscos2(e[:n], e[n:], y.cpu(), avg=False))
    assert abs(l_eval - l_ref) <= (1e-4 if precision == "fp32" else 1e-2) * abs(l_ref)


def test_autograd_surface_with_dropout_matches_the_oracle_under_the_same_masks():
    torch.manual_seed(1)
    net = SiameseNetwork(precision="bf16", **CFG).to(DEV)       # (dropout takes the fp32 kernels here)
    net.train()
    n = 512
    x1, x2 = torch.randn(n, 280, device=DEV), torch.randn(n, 280, device=DEV)
    y = torch.where(torch.rand(n, device=DEV) < 0.5, 1.0, -1.0)
    net(x1, x2)                                                 # creates the {seed, step} state
    state = net._drop_state
    masks = [ops.dropout_mask(state, P, l, 2 * n, w) for l, w in enumerate([500, 500, 500, 100])]
    e1, e2 = net(x1, x2)
    loss = coscos2(avg=False)(e1, e2, y)
    loss.backward()
    sd = {k: v.detach().cpu().double().requires_grad_() for k, v in net.state_dict().items()}
    e = onets.siamese_forward_once_dropout(sd, torch.cat([x1, x2]).cpu().double(),
                                           [m.cpu() for m in masks], P)
    l_ref = onets.coscos2(e[:n], e[n:], y.cpu().double(), avg=False)
    l_ref.backward()
    assert abs(float(loss) - float(l_ref)) <= 1e-4 * abs(float(l_ref))
    assert float((torch.cat([e1, e2]).detach().cpu().double() - e.detach()).abs().max()) < 1e-4
    for k, p in net.named_parameters():
        g, g_ref = p.grad.detach().cpu().double(), sd[k].grad
        assert float((g - g_ref).norm() / g_ref.norm()) < 2e-3, k
    net.eval()
    a, _ = net(x1, x2)
    b, _ = net(x1, x2)
    assert torch.equal(a, b)                                    # eval: deterministic, no dropout


def test_trainer_epoch_with_the_reference_default_dropout():
    """p_dropout = 0.1 through FramesDataLoader + TrainerSiamese (pipelined graph sweeps): the
    epoch runs, the training loss falls, the dev sweep (eval mode) is deterministic."""
    from abnet3_b200 import synth, utils
    from abnet3_b200.dataloader import FramesDataLoader
    from abnet3_b200.trainer import TrainerSiamese
    c = synth.make_corpus(400, cluster_size=8, tokens_per_file=50, seed=2, device=DEV)
    same = synth.make_same_pairs(c, 300, seed=3)
    diff = synth.make_diff_pairs(c, 300, seed=4)
    tokens = {"train": (same, diff), "dev": (same[:60], diff[:60])}
    dl = FramesDataLoader.from_tokens(utils.FeatureTable.from_device(c.feat, c.file_off), tokens,
                                      batch_size=2048, randomize_dataset=False)
    torch.manual_seed(0)
    net = SiameseNetwork(**CFG).to(DEV)
    tr = TrainerSiamese(network=net, loss=coscos2(avg=False), optimizer_type="adadelta", lr=0.1,
                        momentum=None, cuda=True, dataloader=dl, log_dir="/tmp/abn_test_runs")
    assert tr.engine.p_drop == P
    d0 = tr.optimize_model(do_training=True)
    for _ in range(3):
        d1 = tr.optimize_model(do_training=True)
    assert tr.train_losses[-1] < tr.train_losses[0] and d1 < d0
    step0 = int(tr.engine._drop_state[1].item())
    assert step0 == 4 * tr.last_sweep["train_batches"]          # one mask set per training step
    net.eval()
    a = float(tr.engine.sweep_table(c.feat, dl.frame_pairs["dev"], 2048, 1, do_training=False).item())
    b = float(tr.engine.sweep_table(c.feat, dl.frame_pairs["dev"], 2048, 1, do_training=False).item())
    assert abs(a - b) <= 1e-6 * abs(a) and int(tr.engine._drop_state[1].item()) == step0
