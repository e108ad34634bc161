"""The oracle against the golden vectors produced by the LIVE reference
(oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import cosine_distance, dtw, path_cost
from oracle import nets as onets


@pytest.fixture(scope="module")
def cos_gold(golden_dir):
    return np.load(os.path.join(golden_dir, "cosine.npz"))


@pytest.fixture(scope="module")
def nets_gold(golden_dir):
    return np.load(os.path.join(golden_dir, "nets.npz"))


@pytest.fixture(scope="module")
def dtw_gold(golden_dir):
    return np.load(os.path.join(golden_dir, "dtw.npz"))


def test_cosine_distance_matches_reference_bitwise(cos_gold):
    # same numpy, same expression order -> identical bits
    for k in range(int(cos_gold["n_cases"])):
        d = cosine_distance(cos_gold["x%d" % k], cos_gold["y%d" % k])
        ref = cos_gold["d%d" % k]
        assert d.dtype == np.float64 and d.shape == ref.shape
        np.testing.assert_array_equal(d, ref)


def test_cosine_distance_identical_frames_raise_like_reference(cos_gold):
    x = cos_gold["x_identical"]
    if int(cos_gold["identical_raises"]):
        with pytest.raises(AssertionError):
            cosine_distance(x, x.copy())
    else:
        cosine_distance(x, x.copy())


def test_cosine_distance_values_are_float32_representable(cos_gold):
    # the reference computes in float32 and only then widens (utils.py:43-53)
    d = cos_gold["d7"]
    np.testing.assert_array_equal(d, d.astype(np.float32).astype(np.float64))


def test_dtw_regression_and_invariants(dtw_gold):
    for k in range(int(dtw_gold["n_cases"])):
        d = dtw_gold["d%d" % k]
        cost, p1, p2, acc = dtw(d, return_acc=True)
        assert cost == float(dtw_gold["cost%d" % k])
        np.testing.assert_array_equal(p1, dtw_gold["p1_%d" % k])
        np.testing.assert_array_equal(p2, dtw_gold["p2_%d" % k])
        n1, n2 = d.shape
        assert p1[0] == 0 and p2[0] == 0 and p1[-1] == n1 - 1 and p2[-1] == n2 - 1
        assert max(n1, n2) <= len(p1) <= n1 + n2 - 1
        steps = np.stack([np.diff(p1), np.diff(p2)], 1)
        assert ((steps >= 0) & (steps <= 1)).all() and (steps.sum(1) >= 1).all()
        assert path_cost(d, p1, p2) == cost == acc[-1, -1]


def test_dtw_matches_bruteforce_recurrence():
    rng = np.random.default_rng(5)
    for n1, n2 in [(1, 1), (2, 5), (6, 3), (7, 7)]:
        d = rng.random((n1, n2))
        C = np.full((n1 + 1, n2 + 1), np.inf)
        C[0, 0] = 0
        for i in range(n1):
            for j in range(n2):
                C[i + 1, j + 1] = d[i, j] + min(C[i, j + 1], C[i, j], C[i + 1, j])
        cost, p1, p2 = dtw(d)
        assert cost == C[n1, n2]
        # numpy-argmin traceback on the inf-bordered matrix (diag, up, left)
        i, j = n1 - 1, n2 - 1
        q1, q2 = [i], [j]
        while i > 0 or j > 0:
            tb = int(np.argmin((C[i, j], C[i, j + 1], C[i + 1, j])))
            if tb == 0:
                i, j = i - 1, j - 1
            elif tb == 1:
                i -= 1
            else:
                j -= 1
            q1.insert(0, i)
            q2.insert(0, j)
        np.testing.assert_array_equal(p1, q1)
        np.testing.assert_array_equal(p2, q2)


def test_dtw_c_oracle_agrees_with_the_independent_numpy_dtw_on_10k_matrices():
    """The unpinned DTW oracle's second opinion: oracle/dtw_numpy.py (inf-bordered matrix,
    argmin traceback, batched numpy) against oracle/dtw_oracle.c -- costs bit for bit, paths
    step for step -- on 10 000 matrices: continuous random ones, float32-valued ones (what the
    GPU hands over), and coarse grids where a large share of the traceback steps are exact
    ties."""
    from oracle.dtw_numpy import dtw_batch
    rng = np.random.default_rng(2024)
    total, tie_steps = 0, 0
    cases = []
    for shape, n, kind in [((12, 12), 2500, "uniform"), ((7, 19), 1500, "uniform"),
                           ((23, 9), 1500, "f32"), ((16, 16), 2500, "grid4"),
                           ((9, 14), 1500, "grid2"), ((1, 11), 250, "grid2"), ((13, 1), 250, "uniform")]:
        if kind == "uniform":
            D = rng.random((n,) + shape)
        elif kind == "f32":
            D = rng.random((n,) + shape).astype(np.float32).astype(np.float64)
        else:
            D = rng.integers(0, int(kind[4:]), (n,) + shape).astype(np.float64) / 4.0
        cases.append(D)
    for D in cases:
        cost, paths = dtw_batch(D)
        for b in range(D.shape[0]):
            c, p1, p2, ties = dtw(D[b], return_ties=True)
            assert c == cost[b]
            np.testing.assert_array_equal(p1, paths[b][0])
            np.testing.assert_array_equal(p2, paths[b][1])
            tie_steps += ties
            total += 1
    assert total == 10000 and tie_steps > 15000        # the tie rule was exercised, a lot


def test_dtw_tie_rule_is_diag_then_up_then_left(dtw_gold):
    for k in range(int(dtw_gold["n_tie_cases"])):
        d = dtw_gold["tie_d%d" % k]
        cost, p1, p2, ties = dtw(d, return_ties=True)
        np.testing.assert_array_equal(p1, dtw_gold["tie_p1_%d" % k])
        np.testing.assert_array_equal(p2, dtw_gold["tie_p2_%d" % k])
        n1, n2 = d.shape
        # constant grid: diagonal while possible from the END, then the border
        assert len(p1) == max(n1, n2)
    cost, p1, p2 = dtw(np.full((3, 5), 1.0))
    np.testing.assert_array_equal(p1, [0, 0, 0, 1, 2])
    np.testing.assert_array_equal(p2, [0, 1, 2, 3, 4])


def test_dtw_rejects_nan_and_negative():
    d = np.ones((3, 3))
    d[1, 1] = np.nan
    with pytest.raises(ValueError):
        dtw(d)
    d[1, 1] = -0.5
    with pytest.raises(ValueError):
        dtw(d)


def _sd(gold, name):
    pre = name + "/sd/"
    return {k[len(pre):]: torch.from_numpy(gold[k]) for k in gold.files
            if k.startswith(pre)}


@pytest.mark.parametrize("name,act", [("sia_sig", "sigmoid"),
                                      ("sia_tanh1", "tanh"),
                                      ("sia_relu0", "relu")])
def test_siamese_restatement_matches_reference(nets_gold, name, act):
    g = nets_gold
    sd = {k: v.clone().requires_grad_(True) for k, v in _sd(g, name).items()}
    x1, x2 = torch.from_numpy(g[name + "/x1"]), torch.from_numpy(g[name + "/x2"])
    y = torch.from_numpy(g[name + "/y"])
    e1 = onets.siamese_forward_once(sd, x1, act)
    e2 = onets.siamese_forward_once(sd, x2, act)
    np.testing.assert_allclose(e1.detach().numpy(), g[name + "/e1"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(e2.detach().numpy(), g[name + "/e2"], rtol=1e-6, atol=1e-7)
    for lname, fn in (("coscos2", onets.coscos2), ("cosmargin", onets.cosmargin)):
        for avg in (True, False):
            for v in sd.values():
                v.grad = None
            loss = fn(onets.siamese_forward_once(sd, x1, act),
                      onets.siamese_forward_once(sd, x2, act), y, avg=avg)
            loss.backward()
            tag = "%s/%s_avg%d" % (name, lname, int(avg))
            np.testing.assert_allclose(loss.item(), g[tag + "/loss"], rtol=1e-6)
            for k, v in sd.items():
                np.testing.assert_allclose(v.grad.numpy(), g["%s/grad/%s" % (tag, k)],
                                           rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("lname,kw", [("coscos2", {}), ("cosmargin", {"margin": 0.5}),
                                      ("cosmargin02", {"margin": 0.2})])
def test_loss_restatement_matches_reference(nets_gold, lname, kw):
    g = nets_gold
    fn = onets.coscos2 if lname == "coscos2" else onets.cosmargin
    y = torch.from_numpy(g["loss/y"])
    for avg in (True, False):
        a = torch.from_numpy(g["loss/e1"]).clone().requires_grad_(True)
        b = torch.from_numpy(g["loss/e2"]).clone().requires_grad_(True)
        loss = fn(a, b, y, avg=avg, **kw)
        loss.backward()
        tag = "loss/%s_avg%d" % (lname, int(avg))
        np.testing.assert_allclose(loss.item(), g[tag + "/loss"], rtol=1e-6)
        np.testing.assert_allclose(a.grad.numpy(), g[tag + "/de1"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(b.grad.numpy(), g[tag + "/de2"], rtol=1e-5, atol=1e-7)


def test_loss_float64_labels(nets_gold):
    g = nets_gold
    loss = onets.coscos2(torch.from_numpy(g["loss/e1"]), torch.from_numpy(g["loss/e2"]),
                         torch.from_numpy(g["loss/y"].astype(np.float64)), avg=False)
    np.testing.assert_allclose(loss.item(), g["loss/coscos2_f64labels/loss"], rtol=1e-6)


def test_multitask_restatement_matches_reference(nets_gold):
    g = nets_gold
    sd = {k: v.clone().requires_grad_(True) for k, v in _sd(g, "multi").items()}
    x1, x2 = torch.from_numpy(g["multi/x1"]), torch.from_numpy(g["multi/x2"])
    ys, yp = torch.from_numpy(g["multi/y_spk"]), torch.from_numpy(g["multi/y_phn"])
    spk1, phn1 = onets.multitask_forward_once(sd, x1)
    spk2, phn2 = onets.multitask_forward_once(sd, x2)
    for nm, v in (("spk1", spk1), ("phn1", phn1), ("spk2", spk2), ("phn2", phn2)):
        np.testing.assert_allclose(v.detach().numpy(), g["multi/" + nm], rtol=1e-6, atol=1e-7)
    lf = lambda a, b, y: onets.coscos2(a, b, y, avg=False)
    loss = onets.weighted_loss_multi(spk1, phn1, spk2, phn2, ys, yp, lf, lf, weight=0.3)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["multi/loss"], rtol=1e-6)
    for k, v in sd.items():
        ref = g["multi/grad/%s" % k]
        if ref.size == 0:       # hidden_layers_spk/phn never receive gradient
            assert v.grad is None
        else:
            np.testing.assert_allclose(v.grad.numpy(), ref, rtol=2e-5, atol=1e-7)


def test_training_loop_restatement_reproduces_the_reference_trajectory(golden_dir):
    """oracle/nets.py + torch.optim.Adadelta (what bench.py's CPU arm times) against the first
    40 losses of the live reference's trajectory fixture, and the fixture's own conditioning."""
    from oracle import trajectory as tj
    gold = np.load(os.path.join(golden_dir, "trajectory.npz"))
    ref, pert = gold["losses"], gold["losses_perturbed"]
    assert len(ref) == tj.STEPS and ref[-1] < 0.1 * ref[0]
    assert np.max(np.abs(pert - ref) / ref) < 1e-4          # well conditioned in the reference itself
    feat = torch.from_numpy(tj.features())
    sd = {k: torch.from_numpy(v).requires_grad_() for k, v in tj.state_dict().items()}
    opt = torch.optim.Adadelta(list(sd.values()), lr=0.1)
    for k, (i1, i2, y) in enumerate(tj.batches()[:40]):
        x = torch.cat([feat[torch.from_numpy(i1).long()], feat[torch.from_numpy(i2).long()]])
        e = onets.siamese_forward_once(sd, x)
        loss = onets.coscos2(e[:tj.BATCH], e[tj.BATCH:], torch.from_numpy(y), avg=False)
        opt.zero_grad()
        loss.backward()
        opt.step()
        assert abs(float(loss.detach()) - ref[k]) <= 1e-4 * ref[k], k
