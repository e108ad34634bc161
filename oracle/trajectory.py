"""Deterministic inputs of the 300-step loss-trajectory fixture -- TEST INFRASTRUCTURE ONLY.

Shared by oracle/make_golden.py (which runs the LIVE reference network, loss and
torch.optim.Adadelta on them, abnet3/trainer.py:73-75, :231-243) and by
tests/test_gpu_trajectory.py (which runs the sm_100a training step on the same inputs).
Everything comes from numpy's default_rng (stable across versions); nothing is stored but
the reference's losses.
"""
import numpy as np

N_ROWS, DIM, BATCH, STEPS = 40000, 280, 2048, 300
LAYERS = [(500, 280), (500, 500), (500, 500), (100, 500)]           # 280-500-500-500-100
KEYS = ["input_emb.0", "hidden_layers.0", "hidden_layers.3", "output_layer.0"]


def features(seed=7):
    """[N_ROWS, 280] float32, AR(1) along the rows: neighbouring frames are similar."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N_ROWS, DIM)).astype(np.float32)
    rho, c = np.float32(0.95), np.float32(np.sqrt(1 - 0.95 ** 2))
    for t in range(1, N_ROWS):
        x[t] = rho * x[t - 1] + c * x[t]
    return x


def batches(seed=8):
    """STEPS batches of (idx1, idx2, y): 'same' = two frames at most 2 apart, 'different' =
    two unrelated frames; half and half, shuffled."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(STEPS):
        i = rng.integers(2, N_ROWS - 2, BATCH)
        same = rng.random(BATCH) < 0.5
        j = np.where(same, i + rng.integers(1, 3, BATCH), rng.integers(0, N_ROWS, BATCH))
        out.append((i.astype(np.int32), j.astype(np.int32), np.where(same, 1.0, -1.0).astype(np.float32)))
    return out


W_SCALE = 4.0


def state_dict(seed=9, perturb=0.0):
    """Uniform weights with W_SCALE x the Xavier bound of abnet3/model.py:172-177, zero biases.

    Why not the plain Xavier bound: from it, coscos2 on sigmoid outputs sits on a symmetric
    plateau (loss ~1000 for ~100 steps) and then breaks away at a moment that is chaotic in the
    initial weights -- a 1e-6 relative perturbation of the weights moves the losses after the
    break by 25-45 % in the REFERENCE ITSELF (measured with the live modules), so no
    implementation can be compared with a stored trajectory there.  From 4 x the bound the loss
    descends at once (988 -> 68 in 200 steps) and the trajectory is well conditioned: the same
    perturbation moves no loss by more than 3e-6 (``losses_perturbed`` in the fixture).
    ``perturb``: relative N(0, perturb) noise on every weight, for that measurement."""
    rng = np.random.default_rng(seed)
    prng = np.random.default_rng(seed + 1000)
    sd = {}
    for key, (n_out, n_in) in zip(KEYS, LAYERS):
        bound = W_SCALE * np.sqrt(6.0 / (n_in + n_out))
        w = (rng.random((n_out, n_in)) * 2 - 1) * bound
        if perturb:
            w = w * (1.0 + perturb * prng.standard_normal(w.shape))
        sd[key + ".weight"] = w.astype(np.float32)
        sd[key + ".bias"] = np.zeros(n_out, dtype=np.float32)
    return sd
