"""GPU: the forward pass with the activations resident in shared memory (abn_mlp_forward_fused)
against the same layers run one by one through the grouped GEMM (abn_gemm_bf16_group): same
arithmetic, same k order -- the outputs must be IDENTICAL, hidden activations (bf16, with their
column of ones) and embeddings (fp32) alike; and against float64 arithmetic within the bf16
tolerance."""
import numpy as np
import pytest
import torch

from abnet3_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bf(rows, cols, seed, scale=1.0, pad_val=7.0):
    g = torch.Generator().manual_seed(seed)
    t = torch.full((rows, ops.pad_row(cols + 1)), pad_val, dtype=torch.bfloat16)
    t[:, :cols] = (torch.randn(rows, cols, generator=g) * scale).bfloat16()
    return t.to(DEV)


def _run_both(rows, dims, act, last_act=None):
    last_act = last_act or act
    n_layers = len(dims) - 1
    x = _bf(rows, dims[0], 1)
    Ws = [_bf(dims[l + 1], dims[l], 10 + l, dims[l] ** -0.5) for l in range(n_layers)]
    bs = [(torch.randn(dims[l + 1], generator=torch.Generator().manual_seed(20 + l)) * 0.2).to(DEV)
          for l in range(n_layers)]
    acts = [act] * (n_layers - 1) + [last_act]

    def buffers():
        hid = [torch.zeros((rows, ops.pad_row(dims[l + 1] + 1)), dtype=torch.bfloat16, device=DEV)
               for l in range(n_layers - 1)]
        return hid + [torch.full((rows, dims[-1]), float("nan"), device=DEV)]

    ref = buffers()
    h = x
    for l in range(n_layers):
        ops.gemm_group([ops.gemm_problem(h, Ws[l], rows, dims[l + 1], dims[l], ops.GE_BIAS_ACT, ref[l],
                                         act=acts[l], bias=bs[l], ones_col=(l < n_layers - 1))])
        h = ref[l]
    got = buffers()
    layers = ops.mlp_layers([(Ws[l], dims[l], bs[l], acts[l], got[l], l < n_layers - 1)
                             for l in range(n_layers)])
    ops.mlp_forward_fused(x, rows, layers)
    torch.cuda.synchronize()
    return x, Ws, bs, acts, ref, got


@pytest.mark.parametrize("rows,dims,act", [
    (16384, [280, 500, 500, 500, 100], "sigmoid"),       # the canonical embedder, one training batch
    (256, [280, 500, 500, 500, 100], "sigmoid"),
    (1000, [280, 500, 500, 100], "tanh"),                # ragged last row block
    (777, [40, 64, 100], "relu"),
    (5000, [280, 511, 256, 255, 200], "sigmoid"),        # widths at the slab / tile edges
    (130, [512, 100, 36], "none"),
    (40000, [280, 500, 100], "sigmoid"),                 # more row blocks than CTA pairs
    (300, [280, 100], "sigmoid"),                        # a single layer
])
def test_fused_forward_is_bit_identical_to_the_layer_by_layer_gemms(rows, dims, act):
    x, Ws, bs, acts, ref, got = _run_both(rows, dims, act)
    for l in range(len(dims) - 1):
        n = dims[l + 1]
        if l < len(dims) - 2:
            assert torch.equal(got[l][:, :n + 1], ref[l][:, :n + 1]), "hidden layer %d" % l
            assert bool((got[l][:, n] == 1).all())
        else:
            assert torch.equal(got[l], ref[l]), "embeddings"


def test_fused_forward_against_float64():
    rows, dims = 2048, [280, 500, 500, 500, 100]
    x, Ws, bs, acts, ref, got = _run_both(rows, dims, "sigmoid")
    h = x[:, :280].double().cpu()
    for l in range(4):
        h = torch.sigmoid(h @ Ws[l][:, :dims[l]].double().cpu().T + bs[l].double().cpu())
        if l < 3:
            h = h.bfloat16().double()      # the hidden activations are stored (and re-read) as bf16
    np.testing.assert_allclose(got[-1].cpu().numpy(), h.numpy(), rtol=1e-2, atol=1e-2)


def test_fused_forward_refuses_layers_wider_than_the_slab():
    from abnet3_b200._lib import AbnError
    rows, dims = 256, [280, 600, 100]
    x = _bf(rows, 280, 1)
    W0, W1 = _bf(600, 280, 2), _bf(100, 600, 3)
    h = torch.zeros((rows, ops.pad_row(601)), dtype=torch.bfloat16, device=DEV)
    out = torch.zeros((rows, 100), device=DEV)
    layers = ops.mlp_layers([(W0, 280, None, "sigmoid", h, True), (W1, 600, None, "sigmoid", out, False)])
    with pytest.raises(AbnError):
        ops.mlp_forward_fused(x, rows, layers)


@pytest.mark.parametrize("rows,dims,act", [
    (16384, [280, 500, 500, 500, 100], "sigmoid"),       # the canonical embedder
    (1000, [280, 500, 500, 100], "tanh"),
    (777, [40, 64, 100], "relu"),
    (5000, [280, 511, 256, 255, 200], "sigmoid"),
    (40000, [280, 500, 100], "sigmoid"),
    (130, [64, 512, 36], "none"),
])
def test_fused_dgrad_chain_is_bit_identical_to_the_layer_by_layer_gemms(rows, dims, act):
    """dz_below = (dz W) * act'(y_below) from the top layer down to the first hidden layer."""
    n_layers = len(dims) - 1
    g = torch.Generator().manual_seed(5)
    Ws = [_bf(dims[l + 1], dims[l], 40 + l, dims[l + 1] ** -0.5) for l in range(n_layers)]
    ys = [None] + [_bf(rows, dims[l], 50 + l) for l in range(1, n_layers)]      # outputs of layers 0..n-2
    for l in range(1, n_layers):
        ys[l][:, :dims[l]] = torch.sigmoid(torch.randn(rows, dims[l], generator=g)).bfloat16().to(DEV)
    dz_top = _bf(rows, dims[-1], 60, 0.5)

    def buffers():
        return [None] + [torch.zeros((rows, ops.pad_row(dims[l])), dtype=torch.bfloat16, device=DEV)
                         for l in range(1, n_layers)]

    ref = buffers()
    dz = dz_top
    for l in range(n_layers - 1, 0, -1):        # layer l: dz [rows, dims[l+1]] -> dz_below [rows, dims[l]]
        ops.gemm_group([ops.gemm_problem(dz, Ws[l], rows, dims[l], dims[l + 1], ops.GE_DACT, ref[l], b_mn=True,
                                         act=act, yprev=ys[l])])
        dz = ref[l]
    got = buffers()
    layers = ops.mlp_dlayers([(Ws[l], dims[l], act, ys[l], got[l]) for l in range(n_layers - 1, 0, -1)])
    ops.mlp_dgrad_fused(dz_top, rows, layers)
    torch.cuda.synchronize()
    for l in range(1, n_layers):
        assert torch.equal(got[l][:, :dims[l]], ref[l][:, :dims[l]]), "dz of layer %d" % l


# ---- the pair loss inside the forward chain's last epilogue (abn_mlp_forward_loss_fused) ----------
@pytest.mark.parametrize("n_pairs,dims,act,kind", [
    (8192, [280, 500, 500, 500, 100], "sigmoid", "coscos2"),     # the canonical training batch
    (8192, [280, 500, 500, 500, 100], "sigmoid", "cosmargin"),
    (501, [280, 500, 100], "tanh", "coscos2"),                   # ragged last row block
    (333, [40, 64, 36], "relu", "cosmargin"),                    # one 64-column block
    (70, [280, 128], "none", "coscos2"),                         # single layer, both blocks full
    (40000, [280, 500, 64], "sigmoid", "coscos2"),               # more row blocks than CTA pairs
])
def test_loss_fused_into_the_forward_chain_matches_forward_then_loss_kernel(n_pairs, dims, act, kind):
    """Interleaved rows (2k, 2k + 1 = the two frames of pair k) through the fused launch against
    the stacked rows (k, n + k) through abn_mlp_forward_fused + abn_pair_loss_dz: the embeddings
    are bit-identical (same MMAs per row), so dz is, and the loss agrees to fp32 summation order."""
    rows = 2 * n_pairs
    n_layers = len(dims) - 1
    x = _bf(rows, dims[0], 3)                              # stacked: x1 rows, then x2 rows
    xi = torch.empty_like(x)
    xi[0::2], xi[1::2] = x[:n_pairs], x[n_pairs:]
    Ws = [_bf(dims[l + 1], dims[l], 30 + l, dims[l] ** -0.5) for l in range(n_layers)]
    bs = [(torch.randn(dims[l + 1], generator=torch.Generator().manual_seed(40 + l)) * 0.2).to(DEV)
          for l in range(n_layers)]
    acts = [act] * n_layers
    g = torch.Generator().manual_seed(5)
    y = torch.randint(0, 3, (n_pairs,), generator=g).float().sub(1).to(DEV)     # -1 / 0 / +1
    ld_dz = ops.pad_row(dims[-1])

    def run(xin, fused):
        hid = [torch.zeros((rows, ops.pad_row(dims[l + 1] + 1)), dtype=torch.bfloat16, device=DEV)
               for l in range(n_layers - 1)]
        emb = torch.full((rows, dims[-1]), float("nan"), device=DEV)
        dz = torch.zeros((rows, ld_dz), dtype=torch.bfloat16, device=DEV)
        loss = torch.zeros(1, device=DEV)
        layers = ops.mlp_layers([(Ws[l], dims[l], bs[l], acts[l], (hid + [emb])[l], l < n_layers - 1)
                                 for l in range(n_layers)])
        if fused:
            ops.mlp_forward_loss_fused(xin, rows, layers, y, dz, kind, 0.4, 1.0 / n_pairs, loss_out=loss,
                                       write_embeddings=True)
        else:
            ops.mlp_forward_fused(xin, rows, layers)
            ops.pair_loss_dz(emb[:n_pairs], emb[n_pairs:], y, dz[:n_pairs], dz[n_pairs:], kind, 0.4,
                             1.0 / n_pairs, act, loss_out=loss)
        torch.cuda.synchronize()
        return emb, dz, loss

    emb_r, dz_r, loss_r = run(x, False)
    emb_f, dz_f, loss_f = run(xi, True)
    d = dims[-1]
    assert torch.equal(emb_f[0::2], emb_r[:n_pairs]) and torch.equal(emb_f[1::2], emb_r[n_pairs:])
    # dot / norms are summed in another order (per 64-column block instead of 8 lanes x 4): the
    # cosine differs in its last bits, dz (bf16) in at most one unit of ITS last place
    a = torch.cat([dz_f[0::2, :d], dz_f[1::2, :d]]).float()
    b = dz_r[:, :d].float()
    assert float((a - b).abs().max()) <= 2 ** -7 * float(b.abs().max()) + 1e-12
    assert float((a != b).float().mean()) < 0.05
    np.testing.assert_allclose(loss_f.item(), loss_r.item(), rtol=2e-5, atol=1e-7)


def test_loss_fused_without_the_embeddings_leaves_their_buffer_alone():
    rows, dims = 512, [280, 500, 100]
    x = _bf(rows, dims[0], 3)
    Ws = [_bf(dims[l + 1], dims[l], 30 + l, dims[l] ** -0.5) for l in range(2)]
    bs = [torch.zeros(dims[l + 1], device=DEV) for l in range(2)]
    hid = torch.zeros((rows, ops.pad_row(501)), dtype=torch.bfloat16, device=DEV)
    emb = torch.full((rows, 100), 3.0, device=DEV)
    dz = torch.zeros((rows, ops.pad_row(100)), dtype=torch.bfloat16, device=DEV)
    y = torch.ones(rows // 2, device=DEV)
    layers = ops.mlp_layers([(Ws[0], 280, bs[0], "sigmoid", hid, True), (Ws[1], 500, bs[1], "sigmoid", emb, False)])
    loss = ops.mlp_forward_loss_fused(x, rows, layers, y, dz, "coscos2", 0.5, 1.0)
    torch.cuda.synchronize()
    assert bool((emb == 3.0).all()) and float(loss) > 0 and bool((dz[:, :100] != 0).any())
    with pytest.raises(Exception):
        ops.mlp_forward_loss_fused(x, rows - 1, layers, y, dz, "coscos2", 0.5, 1.0)     # odd row count
