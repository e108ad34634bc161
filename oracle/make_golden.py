"""Generate tests/golden/*.npz by running the LIVE reference in this container.

TEST INFRASTRUCTURE ONLY.  Run here (where /root/reference exists):

    python -m oracle.make_golden

/root/reference does not exist on the GPU box, so only the committed .npz
fixtures travel.  The reference is imported UNMODIFIED; three of its imports
are absent from this image and irrelevant to the functions called, so empty
stub modules stand in for them (`h5features`, `dtw`, `tensorboardX`), and
`scipy.arccos` (removed from modern scipy; it was an alias of numpy.arccos) is
restored as that alias.

Fixtures written
----------------
cosine.npz   abnet3.utils.cosine_distance on seeded float32 inputs incl. the
             (1,1) branch, zero-norm rows, the pairs_knn.txt token lengths.
nets.npz     abnet3.model.SiameseNetwork / SiameseMultitaskNetwork forward,
             abnet3.loss.coscos2 / cosmargin / weighted_loss_multi values and
             autograd gradients, for fixed state_dicts and inputs.
trajectory.npz  the losses of 300 training steps of the live reference (network + coscos2 +
             torch.optim.Adadelta) on the seeded inputs of oracle/trajectory.py.
dtw.npz      OUR oracle's DTW (oracle/dtw_oracle.c) on seeded matrices, kept
             as a regression fixture.  Not a reference output: PARITY UNPINNED.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = "/root/reference"


def import_reference():
    for name in ("h5features", "dtw", "tensorboardX"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "dtw":
                m.DTW = None
            if name == "tensorboardX":
                m.SummaryWriter = object
            sys.modules[name] = m
    import scipy
    if not hasattr(scipy, "arccos"):
        scipy.arccos = np.arccos
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import abnet3.utils as ref_utils
    import abnet3.model as ref_model
    import abnet3.loss as ref_loss
    return ref_utils, ref_model, ref_loss


def smooth_tokens(rng, n, dim, rho=0.9):
    """AR(1)-smoothed float32 frames (speech-like: neighbouring frames are
    correlated so DTW paths are not trivial)."""
    x = rng.standard_normal((n, dim)).astype(np.float32)
    for t in range(1, n):
        x[t] = rho * x[t - 1] + np.float32(np.sqrt(1 - rho * rho)) * x[t]
    return x


def make_cosine(ref_utils):
    rng = np.random.default_rng(20251018)
    out = {}
    # token lengths of test/data/dataloader/pairs_knn.txt: 62/9, 68/3(!), 74/92,
    # 56/62, 50/62, 50/56, 50/50, 98/56, 76/98 (+ the (1,1) branch and D=40)
    shapes = [(1, 1, 280), (1, 7, 280), (9, 62, 280), (3, 68, 280),
              (20, 20, 280), (50, 56, 280), (74, 92, 280), (76, 98, 280),
              (33, 47, 40), (80, 21, 280)]
    for k, (n1, n2, dim) in enumerate(shapes):
        x = smooth_tokens(rng, n1, dim)
        y = (smooth_tokens(rng, n2, dim) * np.float32(0.7)
             + np.float32(0.3) * x[np.minimum(np.arange(n2), n1 - 1)])
        y = y.astype(np.float32)
        if k == 4:                       # zero-norm rows on both sides
            x[3] = 0
            x[11] = 0
            y[5] = 0
        if k == 5:                       # zero-norm row on one side only
            y[0] = 0
        with np.errstate(divide="ignore", invalid="ignore"):
            d = ref_utils.cosine_distance(x, y)
        out["x%d" % k], out["y%d" % k], out["d%d" % k] = x, y, d
    out["n_cases"] = np.int64(len(shapes))
    # identical frames: the reference NaN-asserts on a fraction of them
    x = smooth_tokens(rng, 40, 280)
    try:
        with np.errstate(divide="ignore", invalid="ignore"):
            ref_utils.cosine_distance(x, x.copy())
        out["identical_raises"] = np.int64(0)
    except AssertionError:
        out["identical_raises"] = np.int64(1)
    out["x_identical"] = x
    np.savez_compressed(os.path.join(GOLD, "cosine.npz"), **out)
    print("cosine.npz:", len(shapes), "cases; identical frames raise =",
          int(out["identical_raises"]))


def make_nets(ref_model, ref_loss):
    import torch
    out = {}
    torch.manual_seed(1234)
    rng = np.random.default_rng(7)

    def t(a):
        return torch.from_numpy(a)

    # --- SiameseNetwork, two activations / depths -------------------------
    cfgs = {
        "sia_sig": dict(input_dim=40, num_hidden_layers=2, hidden_dim=48,
                        output_dim=20, p_dropout=0.0,
                        activation_layer="sigmoid"),
        "sia_tanh1": dict(input_dim=24, num_hidden_layers=1, hidden_dim=32,
                          output_dim=16, p_dropout=0.0,
                          activation_layer="tanh"),
        "sia_relu0": dict(input_dim=24, num_hidden_layers=0, hidden_dim=32,
                          output_dim=16, p_dropout=0.0,
                          activation_layer="relu"),
    }
    B = 24
    for name, cfg in cfgs.items():
        net = ref_model.SiameseNetwork(**cfg)
        net.train()
        sd = net.state_dict()
        x1 = rng.standard_normal((B, cfg["input_dim"])).astype(np.float32)
        x2 = rng.standard_normal((B, cfg["input_dim"])).astype(np.float32)
        y = rng.choice([-1, 1], B).astype(np.int64)
        y[0], y[1] = 1, -1
        for k, v in sd.items():
            out["%s/sd/%s" % (name, k)] = v.numpy().copy()
        out[name + "/x1"], out[name + "/x2"], out[name + "/y"] = x1, x2, y
        for lname, lcls, kw in (("coscos2", ref_loss.coscos2, {}),
                                ("cosmargin", ref_loss.cosmargin,
                                 {"margin": 0.5})):
            for avg in (True, False):
                net.zero_grad()
                e1, e2 = net(t(x1), t(x2))
                loss = lcls(avg=avg, **kw)(e1, e2, t(y))
                loss.backward()
                tag = "%s/%s_avg%d" % (name, lname, int(avg))
                out[tag + "/loss"] = loss.detach().numpy().copy()
                for k, p in net.named_parameters():
                    out["%s/grad/%s" % (tag, k)] = p.grad.numpy().copy()
        out[name + "/e1"] = e1.detach().numpy().copy()
        out[name + "/e2"] = e2.detach().numpy().copy()

    # --- loss alone, with gradients w.r.t. the embeddings -----------------
    e1 = rng.standard_normal((B, 20)).astype(np.float32)
    e2 = rng.standard_normal((B, 20)).astype(np.float32)
    e2[2] = e1[2] * 2.0            # cos = +1
    e2[3] = -e1[3]                 # cos = -1
    e1[4] = 0.0                    # zero vector: eps clamp
    y = rng.choice([-1, 1], B).astype(np.int64)
    y[5] = 0                       # "other" label keeps the raw cosine
    out["loss/e1"], out["loss/e2"], out["loss/y"] = e1, e2, y
    for lname, lcls, kw in (("coscos2", ref_loss.coscos2, {}),
                            ("cosmargin", ref_loss.cosmargin, {"margin": 0.5}),
                            ("cosmargin02", ref_loss.cosmargin,
                             {"margin": 0.2})):
        for avg in (True, False):
            a = t(e1).clone().requires_grad_(True)
            b = t(e2).clone().requires_grad_(True)
            loss = lcls(avg=avg, **kw)(a, b, t(y))
            loss.backward()
            tag = "loss/%s_avg%d" % (lname, int(avg))
            out[tag + "/loss"] = loss.detach().numpy().copy()
            out[tag + "/de1"] = a.grad.numpy().copy()
            out[tag + "/de2"] = b.grad.numpy().copy()
    # float64 labels, as OriginalDataLoader yields them (np.ones -> float64)
    a = t(e1).clone().requires_grad_(True)
    b = t(e2).clone().requires_grad_(True)
    loss = ref_loss.coscos2(avg=False)(a, b, t(y.astype(np.float64)))
    out["loss/coscos2_f64labels/loss"] = loss.detach().numpy().copy()

    # --- SiameseMultitaskNetwork + weighted_loss_multi ---------------------
    cfg = dict(input_dim=40, num_hidden_layers_shared=2,
               num_hidden_layers_spk=1, num_hidden_layers_phn=1,
               hidden_dim=48, output_dim=20, p_dropout=0.0,
               activation_layer="sigmoid")
    net = ref_model.SiameseMultitaskNetwork(**cfg)
    net.train()
    for k, v in net.state_dict().items():
        out["multi/sd/%s" % k] = v.numpy().copy()
    x1 = rng.standard_normal((B, 40)).astype(np.float32)
    x2 = rng.standard_normal((B, 40)).astype(np.float32)
    y_spk = rng.choice([-1, 1], B).astype(np.int64)
    y_phn = rng.choice([-1, 1], B).astype(np.int64)
    out["multi/x1"], out["multi/x2"] = x1, x2
    out["multi/y_spk"], out["multi/y_phn"] = y_spk, y_phn
    crit = ref_loss.weighted_loss_multi(
        loss_phn=ref_loss.coscos2(avg=False),
        loss_spk=ref_loss.coscos2(avg=False), weight=0.3)
    net.zero_grad()
    spk1, phn1, spk2, phn2 = net(t(x1), t(x2))
    loss = crit(spk1, phn1, spk2, phn2, t(y_spk), t(y_phn))
    loss.backward()
    out["multi/loss"] = loss.detach().numpy().copy()
    for nm, v in (("spk1", spk1), ("phn1", phn1), ("spk2", spk2),
                  ("phn2", phn2)):
        out["multi/" + nm] = v.detach().numpy().copy()
    for k, p in net.named_parameters():
        out["multi/grad/%s" % k] = (p.grad.numpy().copy() if p.grad is not None
                                    else np.zeros(0, np.float32))
    np.savez_compressed(os.path.join(GOLD, "nets.npz"), **out)
    print("nets.npz:", len(out), "arrays")


def make_dtw():
    from oracle.cosine import cosine_distance
    from oracle.dtw import dtw
    rng = np.random.default_rng(99)
    out = {}
    cases = [(1, 1), (1, 9), (9, 1), (9, 62), (20, 20), (50, 56), (80, 21),
             (33, 47)]
    for k, (n1, n2) in enumerate(cases):
        x = smooth_tokens(rng, n1, 40)
        y = smooth_tokens(rng, n2, 40)
        d = cosine_distance(x, y)
        cost, p1, p2 = dtw(d)
        out["d%d" % k], out["cost%d" % k] = d, np.float64(cost)
        out["p1_%d" % k], out["p2_%d" % k] = p1, p2
    # exact-tie grids: the documented tie rule (diag, up, left) decides
    for k, (n1, n2) in enumerate([(4, 4), (3, 7), (7, 3), (5, 5)]):
        d = np.full((n1, n2), 0.25)
        cost, p1, p2 = dtw(d)
        out["tie_d%d" % k], out["tie_cost%d" % k] = d, np.float64(cost)
        out["tie_p1_%d" % k], out["tie_p2_%d" % k] = p1, p2
    out["n_cases"], out["n_tie_cases"] = np.int64(len(cases)), np.int64(4)
    np.savez_compressed(os.path.join(GOLD, "dtw.npz"), **out)
    print("dtw.npz:", len(cases), "+ 4 tie cases (oracle regression, unpinned)")


def make_trajectory(ref_model, ref_loss):
    """300 training steps of the LIVE reference (abnet3.model.SiameseNetwork 280-500-500-500-100
    sigmoid + abnet3.loss.coscos2(avg=False) + torch.optim.Adadelta(lr=0.1), the loop of
    abnet3/trainer.py:231-243) on the deterministic inputs of oracle/trajectory.py, fp32 on the
    CPU.  Only the losses are stored."""
    import torch
    from oracle import trajectory as tj
    torch.set_num_threads(os.cpu_count() or 1)
    feat = torch.from_numpy(tj.features())
    batches = tj.batches()

    def run(perturb):
        net = ref_model.SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500,
                                       output_dim=100, p_dropout=0.0, activation_layer="sigmoid",
                                       batch_norm=False, output_path=None)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in tj.state_dict(perturb=perturb).items()})
        net.train()
        loss_fn = ref_loss.coscos2(avg=False)
        opt = torch.optim.Adadelta(net.parameters(), lr=0.1)
        losses = []
        for i1, i2, y in batches:
            e1, e2 = net(feat[torch.from_numpy(i1).long()], feat[torch.from_numpy(i2).long()])
            loss = loss_fn(e1, e2, torch.from_numpy(y))
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        return np.asarray(losses, dtype=np.float64), net

    losses, net = run(0.0)
    perturbed, _ = run(1e-6)          # how well conditioned the trajectory is, in the reference itself
    norms = {k.replace(".", "_"): float(v.detach().double().norm()) for k, v in net.state_dict().items()}
    np.savez(os.path.join(GOLD, "trajectory.npz"), losses=losses, losses_perturbed=perturbed,
             **{"norm_" + k: np.float64(v) for k, v in norms.items()})
    print("trajectory.npz: loss %.3f -> %.3f over %d steps; 1e-6 perturbation moves a loss by at "
          "most %.2e relative" % (losses[0], losses[-1], len(losses),
                                  float(np.max(np.abs(perturbed - losses) / losses))))


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref_utils, ref_model, ref_loss = import_reference()
    make_cosine(ref_utils)
    make_nets(ref_model, ref_loss)
    make_dtw()
    make_trajectory(ref_model, ref_loss)


if __name__ == "__main__":
    main()
