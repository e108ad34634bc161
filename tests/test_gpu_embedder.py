"""GPU: EmbedderSiamese / EmbedderSiameseMultitask (abnet3/embedder.py:54-148) against the
oracle's restatement of forward_once (pinned to the live reference by tests/golden/nets.npz) on the
same weights and frames: fp32 networks within 1e-4, bf16 (tensor-core) networks within the stated
1e-2; the reference's chunking (np.array_split into len // batch_size + 1 pieces), the first
output of network(feat, feat) == forward_once(feat), the item / times bookkeeping of the written
archive."""
import os

import numpy as np
import pytest
import torch

from oracle import nets as onets
from abnet3_b200.embedder import EmbedderSiamese, EmbedderSiameseMultitask
from abnet3_b200.model import SiameseNetwork, SiameseMultitaskNetwork

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _corpus(seed, lengths, dim=280):
    rng = np.random.default_rng(seed)
    return {"item%02d" % i: rng.standard_normal((n, dim)).astype(np.float32) for i, n in enumerate(lengths)}


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_siamese_embedder_matches_the_oracle_forward(tmp_path, precision, tol):
    torch.manual_seed(3)
    net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
                         activation_layer="sigmoid", precision=precision)
    feats = _corpus(0, [1, 37, 5000, 5001, 12345, 0 + 256])
    out_path = str(tmp_path / "emb.npz")
    emb = EmbedderSiamese(network=net, feature_path=feats, output_path=out_path, batch_size=5000).embed()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    for name, x in feats.items():
        want = onets.siamese_forward_once(sd, torch.from_numpy(x)).numpy()
        assert emb[name].shape == want.shape and emb[name].dtype == np.float32
        np.testing.assert_allclose(emb[name], want, rtol=tol, atol=tol)
    # what the reference computes, network(feat, feat)[0], is the same rows
    x = torch.from_numpy(feats["item01"]).to(DEV)
    with torch.no_grad():
        e1, e2 = net(x, x)
    np.testing.assert_allclose(e1.cpu().numpy(), emb["item01"], rtol=tol, atol=tol)
    assert torch.equal(e1, e2)
    # archive: items in file order with their times
    z = np.load(out_path)
    assert [k for k in z.files if not k.startswith("times/")] == list(feats)
    np.testing.assert_array_equal(z["item02"], emb["item02"])
    np.testing.assert_allclose(z["times/item01"], 0.0025 + 0.01 * np.arange(37))


def test_multitask_embedder_matches_the_oracle_forward(tmp_path):
    torch.manual_seed(4)
    net = SiameseMultitaskNetwork(input_dim=280, num_hidden_layers_shared=2, num_hidden_layers_spk=1,
                                  num_hidden_layers_phn=1, hidden_dim=500, output_dim=100, p_dropout=0.0,
                                  activation_layer="sigmoid", precision="bf16")
    feats = _corpus(1, [700, 3, 9000])
    out_path = str(tmp_path / "emb")
    spk, phn = EmbedderSiameseMultitask(network=net, feature_path=feats, output_path=out_path,
                                        batch_size=4096).embed()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    for name, x in feats.items():
        w_spk, w_phn = onets.multitask_forward_once(sd, torch.from_numpy(x))
        np.testing.assert_allclose(spk[name], w_spk.numpy(), rtol=1e-2, atol=1e-2)
        np.testing.assert_allclose(phn[name], w_phn.numpy(), rtol=1e-2, atol=1e-2)
    assert os.path.exists(out_path + ".spk.npz") and os.path.exists(out_path + ".phn.npz")


def test_wide_network_takes_the_layer_by_layer_path():
    torch.manual_seed(5)
    net = SiameseNetwork(input_dim=40, num_hidden_layers=1, hidden_dim=600, output_dim=64, p_dropout=0.0,
                         activation_layer="tanh", precision="bf16")
    feats = {k: 0.5 * v for k, v in _corpus(2, [1000], dim=40).items()}
    emb = EmbedderSiamese(network=net, feature_path=feats, output_path=None).embed()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    want = onets.siamese_forward_once(sd, torch.from_numpy(feats["item00"]), activation="tanh").numpy()
    np.testing.assert_allclose(emb["item00"], want, rtol=3e-2, atol=3e-2)      # tanh, K = 600, bf16 operands
