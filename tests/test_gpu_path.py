"""GPU: the whole path through the reference-facing surface -- FramesDataLoader (device
frame-pair table) -> TrainerSiamese.optimize_model -> engine.sweep_table (one CUDA graph
replay per batch, batch position on the device) -- against the same batches stepped one by
one through the eager engine, for SiameseNetwork and SiameseMultitaskNetwork."""
import numpy as np
import pytest
import torch

from abnet3_b200 import ops, synth, utils
from abnet3_b200.dataloader import FramesDataLoader, MultiTaskFramesDataLoader
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.loss import coscos2, weighted_loss_multi
from abnet3_b200.model import SiameseNetwork, SiameseMultitaskNetwork
from abnet3_b200.trainer import TrainerSiamese, TrainerSiameseMultitask, _loss_spec

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    """max |a - b| against the tensor's scale (near-zero biases: the fp32 reductions of the
    weight-gradient GEMM add their partial sums in a run-dependent order)."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-2))


def _net(multitask, seed=0, hidden=500):
    torch.manual_seed(seed)
    if multitask:
        net = SiameseMultitaskNetwork(input_dim=280, num_hidden_layers_shared=2, num_hidden_layers_spk=1,
                                      num_hidden_layers_phn=1, hidden_dim=hidden, output_dim=100,
                                      p_dropout=0.0, activation_layer="sigmoid").to(DEV)
        loss = weighted_loss_multi(avg=False, loss_phn=coscos2(avg=False), loss_spk=coscos2(avg=False),
                                   weight=0.3)
    else:
        net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=hidden, output_dim=100,
                             p_dropout=0.0, activation_layer="sigmoid").to(DEV)
        loss = coscos2(avg=False)
    assert net.precision == "bf16"                   # the default on CUDA
    return net, loss


def _table(n_fp, n_rows, multitask, seed=3):
    g = torch.Generator(device=DEV).manual_seed(seed)
    feat = torch.randn(n_rows, 280, device=DEV, generator=g)
    idx1 = torch.randint(0, n_rows, (n_fp,), device=DEV, dtype=torch.int32, generator=g)
    idx2 = torch.randint(0, n_rows, (n_fp,), device=DEV, dtype=torch.int32, generator=g)
    cols = [(torch.randint(0, 2, (n_fp,), device=DEV, generator=g) * 2 - 1).to(torch.int8)
            for _ in range(2 if multitask else 1)]
    return feat, (idx1, idx2) + tuple(cols)


@pytest.mark.parametrize("pipeline", ["1", "0"])
@pytest.mark.parametrize("multitask", [False, True])
def test_sweep_table_equals_batches_stepped_one_by_one(multitask, pipeline, monkeypatch):
    """pipeline=1 (default): the next batch is gathered on a side stream beside the current step
    (second operand set, fork / join inside the CUDA graph); 0: one stream, one operand set."""
    monkeypatch.setenv("ABN_PIPELINE", pipeline)
    B, nb, start = 1000, 7, 3000
    feat, table = _table(12000, 5000, multitask)
    net_a, loss = _net(multitask)
    net_b, _ = _net(multitask)
    net_b.load_state_dict(net_a.state_dict())
    ea = SiameseTrainStep(net_a, _loss_spec(loss), "adadelta", lr=0.1, momentum=None)
    eb = SiameseTrainStep(net_b, _loss_spec(loss), "adadelta", lr=0.1, momentum=None)
    total = float(ea.sweep_table(feat, table, B, nb, start=start, do_training=True).item())
    ref_total = 0.0
    for k in range(nb):
        sl = slice(start + k * B, start + (k + 1) * B)
        x = torch.cat([feat[table[0][sl].long()], feat[table[1][sl].long()]])
        labels = [c[sl].float() for c in table[2:]]
        ref_total += float(eb.step(x, B, *labels).item())
    assert abs(total - ref_total) <= 1e-5 * abs(ref_total), (total, ref_total)
    # the device-side batch position (a pipelined sweep has prefetched one batch more)
    assert int(ea._cursor[0].item()) == start + (nb + (pipeline == "1")) * B
    assert (ea._pipe is not None) and ea._can_pipeline(True) == (pipeline == "1")
    for (k, a), (_, b) in zip(net_a.state_dict().items(), net_b.state_dict().items()):
        assert _rel(a, b) < 2e-3, k
    # a sweep that ends exactly at the end of the table (the prefetch past it gathers nothing)
    last = float(ea.sweep_table(feat, table, B, 2, start=table[0].numel() - 2 * B).item())
    ref_last = 0.0
    for k in range(2):
        sl = slice(table[0].numel() - (2 - k) * B, table[0].numel() - (1 - k) * B)
        x = torch.cat([feat[table[0][sl].long()], feat[table[1][sl].long()]])
        ref_last += float(eb.step(x, B, *[c[sl].float() for c in table[2:]]).item())
    assert abs(last - ref_last) <= 1e-4 * abs(ref_last), (last, ref_last)
    # evaluation sweep: no weight changes, same loss as eager forward + loss
    before = ea.bucket.param.clone()
    ev = float(ea.sweep_table(feat, table, B, 3, start=0, do_training=False).item())
    assert torch.equal(before, ea.bucket.param)
    ref_ev = 0.0
    for k in range(3):
        sl = slice(k * B, (k + 1) * B)
        x = torch.cat([feat[table[0][sl].long()], feat[table[1][sl].long()]])
        ref_ev += float(eb.step(x, B, *[c[sl].float() for c in table[2:]], do_training=False).item())
    assert abs(ev - ref_ev) <= 1e-3 * abs(ref_ev), (ev, ref_ev)


@pytest.mark.parametrize("multitask", [False, True])
def test_trainer_epoch_over_frames_dataloader(multitask):
    """optimize_model over FramesDataLoader.from_tokens == the same shuffled table swept by a
    second engine; dev loss reported like the reference (summed, trainer.py:256)."""
    c = synth.make_corpus(400, cluster_size=8, tokens_per_file=50, seed=2, device=DEV)
    same = synth.make_same_pairs(c, 300, seed=3)
    diff = synth.make_diff_pairs(c, 300, seed=4)

    def spk(tok):
        return torch.where((tok[:, 0] // 3000) == (tok[:, 2] // 3000), 1, -1).to(torch.int8)

    def toks(s, d):
        return (s, d, [(spk(s), spk(d))]) if multitask else (s, d)

    tokens = {"train": toks(same, diff), "dev": toks(same[:60], diff[:60])}
    table = utils.FeatureTable.from_device(c.feat, c.file_off)
    assert table.stack == 7
    Loader = MultiTaskFramesDataLoader if multitask else FramesDataLoader
    Trainer = TrainerSiameseMultitask if multitask else TrainerSiamese
    dl = Loader.from_tokens(table, tokens, batch_size=2048, randomize_dataset=False,
                            exact_numpy_shuffle=True)
    net, loss = _net(multitask)
    ref_net, _ = _net(multitask)
    ref_net.load_state_dict(net.state_dict())
    tr = Trainer(network=net, loss=loss, optimizer_type="adadelta", lr=0.1, momentum=None,
                 cuda=True, dataloader=dl, log_dir="/tmp/abn_test_runs")
    np.random.seed(0)
    dev_loss = tr.optimize_model(do_training=True)
    tab = dl.frame_pairs["train"]
    assert len(tab) == (4 if multitask else 3)
    n = tab[0].numel()
    nb = n // 2048
    assert nb >= 5 and tr.last_sweep["train_batches"] == nb
    # every same pair contributed its whole DTW path, every diff pair min(n1, n2) rows
    res = ops.align_pairs(c.feat, same, stack=7)
    assert n == int(res.path_len.sum().item()) + int(torch.minimum(diff[:, 1], diff[:, 3]).sum().item())
    ref = SiameseTrainStep(ref_net, _loss_spec(loss), "adadelta", lr=0.1, momentum=None)
    ref_total = float(ref.sweep_table(c.feat, tab, 2048, nb, start=0, graph=False).item())
    assert abs(tr.train_losses[-1] * nb - ref_total) <= 1e-4 * abs(ref_total)
    for (k, a), (_, b) in zip(net.state_dict().items(), ref_net.state_dict().items()):
        assert _rel(a, b) < 2e-3, k
    dtab = dl.frame_pairs["dev"]
    ndb = max(dtab[0].numel() // 2048, 1)
    ref_dev = float(ref.sweep_table(c.feat, dtab, min(2048, dtab[0].numel()), ndb, do_training=False,
                                    graph=False).item())
    assert abs(dev_loss - ref_dev) <= 1e-3 * abs(ref_dev)
    # a second epoch continues to reduce the training loss
    tr.optimize_model(do_training=True)
    assert tr.train_losses[-1] < tr.train_losses[-2]
    # frame batches of the generator surface carry the same rows / labels
    batch = next(dl.batch_iterator(train_mode=True))
    assert len(batch) == (4 if multitask else 3) and batch[0].shape == (2048, 280)
    assert torch.equal(batch[0], c.feat[tab[0][:2048].long()])
    assert torch.equal(batch[-1], tab[-1][:2048].float())
    if multitask:
        assert torch.equal(batch[2], tab[2][:2048].float())


def test_engine_refreshes_bf16_weights_after_load_state_dict():
    feat, table = _table(4000, 3000, False)
    net, loss = _net(False, seed=1)
    other, _ = _net(False, seed=2)
    eng = SiameseTrainStep(net, _loss_spec(loss), "sgd", lr=0.01, momentum=0.0)
    fresh = SiameseTrainStep(other, _loss_spec(loss), "sgd", lr=0.01, momentum=0.0)
    want = float(fresh.sweep_table(feat, table, 1000, 2, do_training=False).item())
    net.load_state_dict(other.state_dict())              # copies into the fp32 masters in place
    got = float(eng.sweep_table(feat, table, 1000, 2, do_training=False).item())
    assert abs(got - want) <= 1e-6 * abs(want)     # (block partial sums are added atomically)


def test_trainer_train_loop_saves_best_checkpoints_like_the_reference(tmp_path):
    """TrainerBuilder.train() (abnet3/trainer.py:117-173): epoch-0 evaluation pass, epochs with the
    train + dev sweeps, best-model checkpoints under the reference's file names, early stopping on
    the summed dev loss; the saved state_dict reloads into a fresh network (reference key names)."""
    c = synth.make_corpus(300, cluster_size=8, tokens_per_file=50, seed=5, device=DEV)
    same = synth.make_same_pairs(c, 200, seed=6)
    diff = synth.make_diff_pairs(c, 200, seed=7)
    tokens = {"train": (same, diff), "dev": (same[:50], diff[:50])}
    dl = FramesDataLoader.from_tokens(utils.FeatureTable.from_device(c.feat, c.file_off), tokens,
                                      batch_size=1024)
    torch.manual_seed(0)
    net = SiameseNetwork(input_dim=280, num_hidden_layers=1, hidden_dim=500, output_dim=100,
                         p_dropout=0.1, activation_layer="sigmoid",
                         output_path=str(tmp_path / "model")).to(DEV)
    tr = TrainerSiamese(network=net, loss=coscos2(avg=False), optimizer_type="adadelta", lr=0.1,
                        momentum=None, cuda=True, dataloader=dl, num_epochs=3, patience=1,
                        checkpoints=True, log_dir=str(tmp_path / "runs"))
    tr.train()
    assert len(tr.train_losses) == len(tr.dev_losses) and 2 <= len(tr.train_losses) <= 4
    assert tr.train_losses[-1] < tr.train_losses[0]              # epoch 0 is the untrained evaluation pass
    assert (tmp_path / "model.pth").exists() and (tmp_path / "model0.pth").exists()
    assert (tmp_path / "model.params").exists()
    sd = torch.load(str(tmp_path / "model.pth"))
    assert sorted(sd) == ["hidden_layers.0.bias", "hidden_layers.0.weight", "input_emb.0.bias",
                          "input_emb.0.weight", "output_layer.0.bias", "output_layer.0.weight"]
    fresh = SiameseNetwork(input_dim=280, num_hidden_layers=1, hidden_dim=500, output_dim=100,
                           p_dropout=0.1, activation_layer="sigmoid").to(DEV)
    fresh.load_network(str(tmp_path / "model.pth"))
    fresh.eval()
    x = torch.randn(64, 280, device=DEV)
    best = SiameseNetwork(input_dim=280, num_hidden_layers=1, hidden_dim=500, output_dim=100,
                          p_dropout=0.1, activation_layer="sigmoid").to(DEV)
    best.load_state_dict(sd)
    best.eval()
    assert torch.equal(fresh.forward_once(x), best.forward_once(x))
