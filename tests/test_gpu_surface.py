"""GPU parity of the reference-facing Python surface: SiameseNetwork /
SiameseMultitaskNetwork / losses against the golden vectors of the LIVE
reference, the fused training step against torch autograd + torch.optim on the
oracle restatement, and the dataloaders against the oracle's restatement of
abnet3/dataloader.py on the same files."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import nets as onets
from oracle.align import FeaturesAccessor, load_frames_from_pairs, load_all_frames
from abnet3_b200 import synth
from abnet3_b200.dataloader import (OriginalDataLoader, FramesDataLoader, MultiTaskDataLoader,
                                    PairsDataLoader)
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.loss import coscos2, cosmargin, weighted_loss_multi
from abnet3_b200.model import SiameseNetwork, SiameseMultitaskNetwork
from abnet3_b200.utils import group_pairs, get_dtw_alignment, cosine_distance, DTW

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(a, b, rtol=1e-4, atol=1e-6):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "nets.npz"))


def _load_sd(net, gold, name):
    pre = name + "/sd/"
    sd = {k[len(pre):]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith(pre)}
    assert sorted(sd) == sorted(net.state_dict())        # same key layout as the reference
    net.load_state_dict(sd)
    return sd


CFGS = {
    "sia_sig": dict(input_dim=40, num_hidden_layers=2, hidden_dim=48, output_dim=20,
                    p_dropout=0.0, activation_layer="sigmoid"),
    "sia_tanh1": dict(input_dim=24, num_hidden_layers=1, hidden_dim=32, output_dim=16,
                      p_dropout=0.0, activation_layer="tanh"),
    "sia_relu0": dict(input_dim=24, num_hidden_layers=0, hidden_dim=32, output_dim=16,
                      p_dropout=0.0, activation_layer="relu"),
}


@pytest.mark.parametrize("name", sorted(CFGS))
def test_siamese_network_matches_reference_golden(gold, name):
    net = SiameseNetwork(precision="fp32", **CFGS[name]).to(DEV)
    _load_sd(net, gold, name)
    net.train()
    x1 = torch.from_numpy(gold[name + "/x1"]).to(DEV)
    x2 = torch.from_numpy(gold[name + "/x2"]).to(DEV)
    y = torch.from_numpy(gold[name + "/y"]).to(DEV)
    e1, e2 = net(x1, x2)
    _close(e1, gold[name + "/e1"])
    _close(e2, gold[name + "/e2"])
    for lname, crit in (("coscos2", coscos2), ("cosmargin", cosmargin)):
        for avg in (True, False):
            net.zero_grad()
            e1, e2 = net(x1, x2)
            loss = crit(avg=avg)(e1, e2, y)
            loss.backward()
            tag = "%s/%s_avg%d" % (name, lname, int(avg))
            _close(loss, gold[tag + "/loss"])
            for k, p in net.named_parameters():
                g = gold["%s/grad/%s" % (tag, k)]
                _close(p.grad, g, atol=1e-4 * max(float(np.abs(g).max()), 1e-3))


def test_multitask_network_matches_reference_golden(gold):
    cfg = dict(input_dim=40, num_hidden_layers_shared=2, num_hidden_layers_spk=1,
               num_hidden_layers_phn=1, hidden_dim=48, output_dim=20, p_dropout=0.0,
               activation_layer="sigmoid")
    net = SiameseMultitaskNetwork(precision="fp32", **cfg).to(DEV)
    _load_sd(net, gold, "multi")
    net.train()
    x1, x2 = torch.from_numpy(gold["multi/x1"]).to(DEV), torch.from_numpy(gold["multi/x2"]).to(DEV)
    ys = torch.from_numpy(gold["multi/y_spk"]).to(DEV)
    yp = torch.from_numpy(gold["multi/y_phn"]).to(DEV)
    spk1, phn1, spk2, phn2 = net(x1, x2)
    for nm, v in (("spk1", spk1), ("phn1", phn1), ("spk2", spk2), ("phn2", phn2)):
        _close(v, gold["multi/" + nm])
    crit = weighted_loss_multi(loss_phn=coscos2(avg=False), loss_spk=coscos2(avg=False), weight=0.3)
    loss = crit(spk1, phn1, spk2, phn2, ys, yp)
    loss.backward()
    _close(loss, gold["multi/loss"])
    for k, p in net.named_parameters():
        g = gold["multi/grad/%s" % k]
        if g.size == 0:
            assert p.grad is None          # hidden_layers_spk/phn never get a gradient
        else:
            _close(p.grad, g, atol=1e-4 * max(float(np.abs(g).max()), 1e-3))


def test_unsupported_configurations_fail_loudly():
    net = SiameseNetwork(input_dim=8, num_hidden_layers=1, hidden_dim=8, output_dim=4,
                         p_dropout=0.1, activation_layer="relu").to(DEV)
    x = torch.randn(4, 8, device=DEV)
    net.eval()
    net(x, x)                                    # dropout is the identity in eval mode
    bn = SiameseNetwork(input_dim=8, num_hidden_layers=1, hidden_dim=8, output_dim=4,
                        p_dropout=0.0, batch_norm=True, activation_layer="relu").to(DEV)
    bn.train()
    with pytest.raises(NotImplementedError):
        bn(x, x)                                 # batch statistics: not implemented
    with pytest.raises(RuntimeError):
        net(x.cpu(), x.cpu())                    # no CPU path


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_batch_norm_network_in_eval_mode_matches_torch_modules(precision, tol):
    """batch_norm=True (abnet3/model.py:136-141: Linear -> Dropout -> BatchNorm1d -> act) in eval
    mode: the running statistics are folded into W and b; checked against torch's own modules
    of the same tree (which is what the reference executes), also through the embedder."""
    from abnet3_b200.embedder import EmbedderSiamese
    torch.manual_seed(3)
    net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100,
                         p_dropout=0.1, batch_norm=True, activation_layer="sigmoid",
                         precision=precision).to(DEV)
    for m in net.modules():                      # non-trivial statistics and affine parameters
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.3)
            m.running_var.uniform_(0.5, 2.0)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    net.eval()
    x = torch.randn(700, 280, device=DEV)
    with torch.no_grad():
        want = net.output_layer(net.hidden_layers(net.input_emb(x)))      # plain torch modules
        got = net.forward_once(x)
    assert float((got - want).abs().max()) < tol
    emb = EmbedderSiamese(network=net, feature_path={"a": x.cpu().numpy()}, output_path=None).embed()
    assert float(np.abs(emb["a"] - want.cpu().numpy()).max()) < tol


@pytest.mark.parametrize("opt,lr", [("sgd", 0.05), ("adadelta", 0.1), ("adam", 0.002)])
def test_fused_train_step_matches_autograd_and_torch_optim(opt, lr):
    """trainer.py:237-240 on the canonical 280-500-500-500-100 network."""
    torch.manual_seed(0)
    net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100,
                         p_dropout=0.0, activation_layer="sigmoid", precision="fp32").to(DEV)
    sd0 = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    step = SiameseTrainStep(net, ("coscos2", 0.0, False), opt, lr=lr, momentum=0.9)
    ref = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    ropt = {"sgd": lambda p: torch.optim.SGD(p, lr=lr, momentum=0.9),
            "adadelta": lambda p: torch.optim.Adadelta(p, lr=lr),
            "adam": lambda p: torch.optim.Adam(p, lr=lr)}[opt](list(ref.values()))
    n = 512
    for it in range(3):
        x = torch.randn(2 * n, 280)
        y = torch.where(torch.rand(n) < 0.5, 1.0, -1.0)
        ropt.zero_grad()
        e = onets.siamese_forward_once(ref, x, "sigmoid")
        rloss = onets.coscos2(e[:n], e[n:], y, avg=False)
        rloss.backward()
        ropt.step()
        loss = step.step(x.to(DEV), n, y.to(DEV))
        _close(loss, rloss.item())
    for k, v in net.state_dict().items():
        if opt == "adam":
            # Adam divides by sqrt(v): entries whose gradient is at the fp32 noise floor get
            # O(lr) steps of either sign, so compare in units of the step size
            diff = (v.cpu() - ref[k].detach()).abs()
            assert float(diff.max()) <= 2 * 3 * lr
            assert float(diff.mean()) <= 0.01 * lr
            assert float((diff > 0.1 * lr).float().mean()) < 1e-3
        else:
            _close(v, ref[k].detach(), rtol=2e-4, atol=2e-6)


def test_graphed_step_equals_eager_step():
    """CUDA-graph replay of the step must give the eager step's parameters."""
    torch.manual_seed(3)
    cfg = dict(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
               activation_layer="sigmoid")
    a, b = SiameseNetwork(precision="fp32", **cfg).to(DEV), SiameseNetwork(precision="fp32", **cfg).to(DEV)
    b.load_state_dict(a.state_dict())
    sa = SiameseTrainStep(a, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
    sb = SiameseTrainStep(b, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
    n = 1024
    for it in range(5):
        x = torch.randn(2 * n, 280, device=DEV)
        y = torch.where(torch.rand(n, device=DEV) < 0.5, 1.0, -1.0)
        la = sa.step(x, n, y).clone()
        lb = sb.step(x, n, y, graph=True).clone()
        _close(lb, la, rtol=1e-5)
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        _close(w, v, rtol=1e-4, atol=1e-6)      # split-K wgrad uses fp32 atomics: order noise only


def test_fused_multitask_step_matches_autograd():
    torch.manual_seed(1)
    net = SiameseMultitaskNetwork(input_dim=280, num_hidden_layers_shared=2,
                                  num_hidden_layers_spk=1, num_hidden_layers_phn=1,
                                  hidden_dim=500, output_dim=100, p_dropout=0.0,
                                  activation_layer="sigmoid", precision="fp32").to(DEV)
    sd0 = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    spec = (("coscos2", 0.0, False), ("coscos2", 0.0, False), 0.3)
    step = SiameseTrainStep(net, spec, "sgd", lr=0.05, momentum=0.0)
    ref = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    n = 256
    x = torch.randn(2 * n, 280)
    ys = torch.where(torch.rand(n) < 0.5, 1.0, -1.0)
    yp = torch.where(torch.rand(n) < 0.5, 1.0, -1.0)
    spk, phn = onets.multitask_forward_once(ref, x)
    lf = lambda a, b, y: onets.coscos2(a, b, y, avg=False)
    rloss = onets.weighted_loss_multi(spk[:n], phn[:n], spk[n:], phn[n:], ys, yp, lf, lf, 0.3)
    rloss.backward()
    loss = step.step(x.to(DEV), n, ys.to(DEV), yp.to(DEV))
    _close(loss, rloss.item())
    for k, v in net.state_dict().items():
        g = ref[k].grad
        want = sd0[k] if g is None else sd0[k] - 0.05 * g
        _close(v, want, rtol=2e-4, atol=2e-6)     # untouched when the reference gives no grad


def test_per_pair_api_matches_oracle():
    from oracle.make_golden import smooth_tokens
    rng = np.random.default_rng(3)
    x = smooth_tokens(rng, 37, 280)
    y = (0.8 * smooth_tokens(rng, 52, 280) + 0.2 * x[np.minimum(np.arange(52), 36)]).astype(np.float32)
    d = cosine_distance(x, y)
    dref = oracle.cosine_distance(x, y)
    assert d.dtype == np.float64 and d.shape == dref.shape
    np.testing.assert_allclose(d, dref, atol=2e-6, rtol=0)
    p1, p2 = get_dtw_alignment(x, y)
    q1, q2 = oracle.get_dtw_alignment(x, y)
    np.testing.assert_array_equal(p1, q1)
    np.testing.assert_array_equal(p2, q2)
    cost, _, (L, a, b) = DTW(x, y, return_alignment=True, dist_array=dref)
    c, r1, r2 = oracle.dtw(dref)
    assert cost == c and L == len(r1)
    np.testing.assert_array_equal(a, r1)
    np.testing.assert_array_equal(b, r2)
    # float64 all-ones frames (the reference's own MockFeaturesAccessor,
    # test/test_dataloader.py:5-8): the dtype assert fails exactly like utils.py:41-42
    with pytest.raises(AssertionError):
        cosine_distance(np.ones((10, 3)), np.ones((10, 3), np.float32))


# ----------------------------------------------------------------- dataloaders
@pytest.fixture(scope="module")
def corpus_files(tmp_path_factory):
    """A small synthetic corpus written the way the reference's tools lay data out:
    a features archive + <pairs>/train_pairs/dataset and dev_pairs/dataset."""
    root = tmp_path_factory.mktemp("corpus")
    c = synth.make_corpus(240, cluster_size=8, tokens_per_file=40, seed=11)
    feat = c.feat.numpy()
    file_off = c.file_off.numpy()
    feats = {"file%d" % i: feat[file_off[i]:file_off[i + 1]] for i in range(len(file_off) - 1)}
    np.savez(root / "features.npz", **feats)
    rng = np.random.default_rng(5)
    same = synth.make_same_pairs(c, 60, seed=6).numpy()
    diff = synth.make_diff_pairs(c, 60, seed=7).numpy()
    tok_file = c.tok_file.numpy()
    starts = c.tok_start.numpy()

    def line(tk, kind):
        out = []
        for s, n in ((tk[0], tk[1]), (tk[2], tk[3])):
            f = int(np.searchsorted(file_off, s, side="right") - 1)
            k0 = s - file_off[f]
            # frame centres are 0.0025 + 0.01 k; window [on, off] picks frames k0 .. k0+n-1
            out += ["file%d" % f, "%.4f" % (0.01 * k0), "%.4f" % (0.01 * (k0 + n - 1) + 0.005)]
        return " ".join(out + [kind])

    lines = [line(t, "same") for t in same] + [line(t, "diff") for t in diff]
    lines.append("file0 0.5000 0.4000 file1 0.1000 0.3000 same")       # s > e: skipped
    order = rng.permutation(len(lines))
    lines = [lines[i] for i in order]
    for mode, sl in (("train_pairs", slice(0, 90)), ("dev_pairs", slice(90, None))):
        os.makedirs(root / mode)
        (root / mode / "dataset").write_text("\n".join(lines[sl]) + "\n")
    spk = root / "spk.txt"
    spk.write_text("".join("file%d spk%d\n" % (i, i // 2) for i in range(len(feats))))
    times = {k: 0.0025 + 0.01 * np.arange(v.shape[0]) for k, v in feats.items()}
    return root, FeaturesAccessor(times, feats)


def _compare_batches(mine, ref):
    assert len(mine) == len(ref)
    for a, b in zip(mine, ref):
        a = a.cpu().numpy()
        assert a.shape == b.shape, (a.shape, b.shape)
        np.testing.assert_array_equal(a, b.astype(a.dtype))


def test_original_dataloader_batches_equal_reference_semantics(corpus_files):
    root, acc = corpus_files
    dl = OriginalDataLoader(str(root), str(root / "features.npz"), num_max_minibatches=5,
                            batch_size=8)
    for train_mode in (True, False):
        np.random.seed(123)
        got = list(dl.batch_iterator(train_mode=train_mode))
        # replay the reference's batch selection (dataloader.py:286-300) with the oracle
        np.random.seed(123)
        pairs = dl.pairs['train' if train_mode else 'dev']
        starts = list(range(0, len(pairs), 8))
        if 5 < len(starts):
            sel = np.random.choice(range(len(starts)), 5, replace=False)
        else:
            sel = np.random.permutation(range(len(starts)))
        assert len(got) == len(sel)
        for batch, b in zip(got, sel):
            ref = load_frames_from_pairs(acc, group_pairs(pairs[starts[b]:starts[b] + 8]))
            _compare_batches(batch, ref)
            assert batch[0].is_cuda and batch[0].dtype == torch.float32
    assert dl.statistics_training['SameType'] > 0 and dl.statistics_training['DiffType'] > 0


def test_original_dataloader_align_different_words_quirk(corpus_files):
    root, acc = corpus_files
    dl = OriginalDataLoader(str(root), str(root / "features.npz"), num_max_minibatches=1000,
                            batch_size=8, align_different_words=True)
    np.random.seed(7)
    got = list(dl.batch_iterator(train_mode=False))
    np.random.seed(7)
    pairs = dl.pairs['dev']
    starts = list(range(0, len(pairs), 8))
    sel = np.random.permutation(range(len(starts)))
    for batch, b in zip(got, sel):
        ref = load_frames_from_pairs(acc, group_pairs(pairs[starts[b]:starts[b] + 8]),
                                     align_different_words=True)
        _compare_batches(batch, ref)


def test_temporal_coherence_batches_match_reference_semantics(corpus_files):
    """tcl > 0 (abnet3/dataloader.py:303-352): every batch gets tcl / (1 - tcl) x its frame
    pairs of (t, t+1) 'same' and (t, t+15..30) 'different' pairs of random files, appended
    after the shuffled batch, with the reference's own ``random`` draws."""
    import random
    from oracle.align import add_tcl_to_batch
    root, acc = corpus_files
    dl = OriginalDataLoader(str(root), str(root / "features.npz"), num_max_minibatches=3,
                            batch_size=8, tcl=0.3)
    np.random.seed(31)
    random.seed(32)
    got = list(dl.batch_iterator(train_mode=True))
    np.random.seed(31)
    random.seed(32)
    pairs = dl.pairs['train']
    starts = list(range(0, len(pairs), 8))
    sel = np.random.choice(range(len(starts)), 3, replace=False)
    assert len(got) == 3
    for batch, b in zip(got, sel):
        ref = load_frames_from_pairs(acc, group_pairs(pairs[starts[b]:starts[b] + 8]))
        n0 = len(ref[2])
        ref = add_tcl_to_batch(acc, ref, 0.3, dl.train_files)
        assert len(ref[2]) > n0                       # pairs were added
        _compare_batches(batch, ref)


def test_multitask_dataloader_matches_reference_semantics(corpus_files):
    root, acc = corpus_files
    from abnet3_b200.utils import read_spkid_file
    dl = MultiTaskDataLoader(str(root), str(root / "features.npz"), fid2spk_file=str(root / "spk.txt"),
                             num_max_minibatches=4, batch_size=8)
    np.random.seed(99)
    got = list(dl.batch_iterator(train_mode=True))
    np.random.seed(99)
    pairs = dl.pairs['train']
    starts = list(range(0, len(pairs), 8))
    sel = np.random.choice(range(len(starts)), 4, replace=False)
    fid2spk = read_spkid_file(str(root / "spk.txt"))
    for batch, b in zip(got, sel):
        ref = load_frames_from_pairs(acc, group_pairs(pairs[starts[b]:starts[b] + 8]),
                                     fid2spk=fid2spk)
        assert len(batch) == 4
        _compare_batches(batch, ref)


def test_frames_dataloader_table_and_batches(corpus_files):
    root, acc = corpus_files
    dl = FramesDataLoader(str(root), str(root / "features.npz"), batch_size=256,
                          randomize_dataset=False, exact_numpy_shuffle=True)
    np.random.seed(5)
    dl.load_data()
    # oracle: same construction + the same np.random.shuffle draws (train first, then dev)
    np.random.seed(5)
    feat_tab = dl.table.host
    for mode in ('train', 'dev'):
        token_feats, frames = load_all_frames(acc, group_pairs(dl.pairs[mode]), shuffle=True)
        idx1, idx2, y = (t.cpu().numpy() for t in dl.frame_pairs[mode])
        assert len(frames) == len(idx1)
        for k in list(range(0, len(frames), max(1, len(frames) // 200))):
            f1, s1, e1, i1, f2, s2, e2, i2, lab = frames[k]
            np.testing.assert_array_equal(feat_tab[idx1[k]], token_feats[f1, s1, e1][i1])
            np.testing.assert_array_equal(feat_tab[idx2[k]], token_feats[f2, s2, e2][i2])
            assert y[k] == lab
    batches = list(dl.batch_iterator(train_mode=True))
    n = dl.frame_pairs['train'][0].numel()
    assert len(batches) == max(n // 256, 1)                  # tail dropped (:708)
    X1, X2, yb = batches[0]
    idx1, idx2, y = dl.frame_pairs['train']
    assert torch.equal(X1, dl.table.feat[idx1[:256].long()])
    assert torch.equal(X2, dl.table.feat[idx2[:256].long()])
    assert torch.equal(yb, y[:256].float())


def test_pairs_dataloader_on_reference_pair_file(golden_dir):
    """Config C1: the reference's own frame-indexed pair file over synthetic files."""
    rng = np.random.default_rng(0)
    feats = {}
    from oracle.make_golden import smooth_tokens
    for i in range(5):
        feats["file%d" % i] = smooth_tokens(rng, 80000, 40, rho=0.95)
    import random
    dl = PairsDataLoader(os.path.join(golden_dir, "pairs_knn.txt"), feats,
                         os.path.join(golden_dir, "id_to_file.txt"), ratio_split_train_test=0.7,
                         batch_size=4, train_iterations=3, test_iterations=2)
    random.seed(4)
    got = list(dl.batch_iterator(train_mode=True))
    assert len(got) == 3
    # replay the sampling (dataloader.py:517-535) and the batches with the oracle
    random.seed(4)
    acc = FeaturesAccessor({k: None for k in feats}, feats)
    allpos, tokens = dl.pairs['train'], dl.tokens['train']
    n_pairs = 3 * 4
    n_pos = min(int(n_pairs * 0.5), len(allpos))
    pos = [p + ['same'] for p in random.sample(allpos, n_pos)]
    toks = random.choices(tokens, k=2 * (n_pairs - n_pos))
    neg = [list(toks[i]) + list(toks[i + 1]) + ["diff"] for i in range(0, len(toks), 2)]
    pairs = pos + neg
    random.shuffle(pairs)
    for i, batch in enumerate(got):
        ref = load_frames_from_pairs(acc, group_pairs(pairs[i * 4:(i + 1) * 4]), frames=True,
                                     align_different_words=True)
        _compare_batches(batch, ref)


def test_c1_one_epoch_on_the_reference_pair_file_gpu_vs_oracle(golden_dir):
    """Config C1 (BASELINE.json configs[0]): the reference's own pair set
    (test/data/dataloader/pairs_knn.txt) over synthetic 280-dim files, SiameseNetwork
    280-500-500-100 sigmoid, coscos2, ONE epoch (train sweep + test sweep) through
    PairsDataLoader + TrainerSiamese on the GPU (fp32 path) against the same epoch replayed on
    the CPU: the oracle's batches (abnet3/dataloader.py:510-546, :166-261), the oracle network
    and loss under torch autograd, torch.optim.SGD(lr, momentum) as abnet3/trainer.py:70-72."""
    import random
    from oracle.make_golden import smooth_tokens
    from abnet3_b200.trainer import TrainerSiamese
    rng = np.random.default_rng(1)
    feats = {"file%d" % i: smooth_tokens(rng, 80000, 280, rho=0.95) for i in range(5)}
    kw = dict(ratio_split_train_test=1.0, batch_size=4, train_iterations=4, test_iterations=2)
    dl = PairsDataLoader(os.path.join(golden_dir, "pairs_knn.txt"), feats,
                         os.path.join(golden_dir, "id_to_file.txt"), **kw)
    # the fixture's 30 pairs all train (its 70 % split leaves one pair, SURVEY 8d "bypass the stale
    # split"); the test sweep runs over the same pairs
    dl.load_pairs()
    assert len(dl.pairs['train']) == 30
    dl.pairs['test'], dl.tokens['test'] = dl.pairs['train'], dl.tokens['train']
    torch.manual_seed(11)
    net = SiameseNetwork(input_dim=280, num_hidden_layers=1, hidden_dim=500, output_dim=100,
                         p_dropout=0.0, activation_layer="sigmoid", precision="fp32").to(DEV)
    sd0 = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    tr = TrainerSiamese(network=net, loss=coscos2(avg=True), optimizer_type="sgd", lr=0.05,
                        momentum=0.9, cuda=True, dataloader=dl, log_dir="/tmp/abn_test_runs")
    random.seed(21)
    dev_loss = tr.optimize_model(do_training=True)

    # ---- the same epoch on the CPU
    acc = FeaturesAccessor({k: None for k in feats}, feats)
    sd = {k: v.clone().requires_grad_() for k, v in sd0.items()}
    opt = torch.optim.SGD(list(sd.values()), lr=0.05, momentum=0.9)
    random.seed(21)

    def epoch_batches(mode, iterations):
        allpos, tokens = dl.pairs[mode], dl.tokens[mode]
        n_pairs = iterations * 4
        n_pos = min(int(n_pairs * 0.5), len(allpos))
        pos = [p + ['same'] for p in random.sample(allpos, n_pos)]
        toks = random.choices(tokens, k=2 * (n_pairs - n_pos))
        neg = [list(toks[i]) + list(toks[i + 1]) + ["diff"] for i in range(0, len(toks), 2)]
        pairs = pos + neg
        random.shuffle(pairs)
        for i in range(iterations):
            batch = pairs[i * 4:(i + 1) * 4]
            if batch:
                yield load_frames_from_pairs(acc, group_pairs(batch), frames=True,
                                             align_different_words=True)

    def loss_of(batch):
        X1, X2, y = (torch.from_numpy(np.ascontiguousarray(a)) for a in batch)
        n = X1.shape[0]
        e = onets.siamese_forward_once(sd, torch.cat([X1, X2]).float())
        return onets.coscos2(e[:n], e[n:], y.float(), avg=True)

    train_total, nb = 0.0, 0
    for batch in epoch_batches('train', 4):
        loss = loss_of(batch)
        opt.zero_grad()
        loss.backward()
        opt.step()
        train_total += float(loss.detach())
        nb += 1
    dev_total, ndb = 0.0, 0
    with torch.no_grad():
        for batch in epoch_batches('test', 2):
            dev_total += float(loss_of(batch))
            ndb += 1
    assert nb == 4 and tr.last_sweep == {'train_batches': nb, 'dev_batches': ndb}
    assert abs(tr.train_losses[-1] - train_total / nb) <= 1e-4 * abs(train_total / nb)
    assert abs(dev_loss - dev_total) <= 1e-4 * abs(dev_total)
    for k, v in net.state_dict().items():
        upd_gpu, upd_cpu = v.detach().cpu() - sd0[k], sd[k].detach() - sd0[k]
        assert float((upd_gpu - upd_cpu).norm()) <= 1e-3 * float(upd_cpu.norm()), k
