"""GPU parity: kernels (1) cosine distance and (2) DTW + traceback through the
C ABI against the CPU oracle on identical seeded inputs."""
import os

import numpy as np
import pytest
import torch

import oracle
from abnet3_b200 import ops, synth, utils

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _flat_dist(mats):
    off = np.zeros(len(mats) + 1, dtype=np.int64)
    off[1:] = np.cumsum([m.size for m in mats])
    flat = np.concatenate([m.ravel() for m in mats]) if mats else np.zeros(0)
    shape = np.array([m.shape for m in mats], dtype=np.int32).reshape(-1, 2)
    return flat.astype(np.float64), off, shape


def _run_dtw(mats):
    flat, off, shape = _flat_dist(mats)
    p1, p2, poff, plen, cost, valid = ops.dtw_from_dist(
        torch.from_numpy(flat).to(DEV), torch.from_numpy(off).to(DEV),
        torch.from_numpy(shape).to(DEV))
    torch.cuda.synchronize()
    return (p1.cpu().numpy(), p2.cpu().numpy(), poff.cpu().numpy(), plen.cpu().numpy(),
            cost.cpu().numpy(), valid.cpu().numpy())


def test_dtw_bit_exact_on_golden_matrices(golden_dir):
    g = np.load(os.path.join(golden_dir, "dtw.npz"))
    mats = [g["d%d" % k] for k in range(int(g["n_cases"]))]
    mats += [g["tie_d%d" % k] for k in range(int(g["n_tie_cases"]))]
    p1, p2, poff, plen, cost, valid = _run_dtw(mats)
    names = ["%d" % k for k in range(int(g["n_cases"]))] + \
            ["tie_%d" % k for k in range(int(g["n_tie_cases"]))]
    for i, nm in enumerate(names):
        pre = "tie_" if nm.startswith("tie_") else ""
        k = nm.replace("tie_", "")
        assert valid[i] == 1
        assert cost[i] == float(g["%scost%s" % (pre, k)])       # bit-exact float64
        s = slice(poff[i], poff[i] + plen[i])
        np.testing.assert_array_equal(p1[s], g["%sp1_%s" % (pre, k)])
        np.testing.assert_array_equal(p2[s], g["%sp2_%s" % (pre, k)])


def test_dtw_bit_exact_on_random_float64_matrices():
    # genuine float64 D (sums round): the wavefront must still match the
    # sequential recurrence bit for bit, and exact ties follow the oracle's rule
    rng = np.random.default_rng(11)
    mats = []
    for _ in range(300):
        n1, n2 = rng.integers(1, 97, 2)
        mats.append(rng.random((n1, n2)))
    for n1, n2 in [(1, 1), (1, 96), (96, 1), (96, 96), (33, 64), (64, 33), (32, 32), (65, 2)]:
        mats.append(rng.random((n1, n2)))
    # coarse grids: many exact ties
    for _ in range(100):
        n1, n2 = rng.integers(2, 60, 2)
        mats.append(rng.integers(0, 3, (n1, n2)).astype(np.float64) * 0.25)
    p1, p2, poff, plen, cost, valid = _run_dtw(mats)
    n_ties = 0
    for i, d in enumerate(mats):
        c, q1, q2, ties = oracle.dtw(d, return_ties=True)
        n_ties += ties
        assert valid[i] == 1 and cost[i] == c
        s = slice(poff[i], poff[i] + plen[i])
        np.testing.assert_array_equal(p1[s], q1)
        np.testing.assert_array_equal(p2[s], q2)
    assert n_ties > 100          # the tie rule really was exercised


def test_dtw_invalid_matrices_are_flagged():
    a = np.full((5, 7), 0.5)
    b = a.copy(); b[2, 3] = np.nan
    c = a.copy(); c[0, 0] = -1.0
    _, _, _, plen, cost, valid = _run_dtw([a, b, c])
    assert list(valid) == [1, 0, 0] and plen[1] == 0 and plen[2] == 0
    assert np.isnan(cost[1]) and np.isnan(cost[2])


def _gpu_distance(feat, pairs):
    dist, off, valid = ops.cosine_distance(torch.from_numpy(feat).to(DEV),
                                           torch.from_numpy(pairs).to(DEV))
    torch.cuda.synchronize()
    return dist.cpu().numpy(), off.cpu().numpy(), valid.cpu().numpy()


def test_cosine_distance_against_reference_golden(golden_dir):
    # golden = the LIVE reference's cosine_distance (float32 arithmetic).  A GPU
    # cannot reproduce BLAS sgemm's summation order, so parity is "float32
    # rounding noise": |d_gpu - d_ref| <= 2e-6 absolute (d in [0, 1]).
    g = np.load(os.path.join(golden_dir, "cosine.npz"))
    for k in range(int(g["n_cases"])):
        x, y, ref = g["x%d" % k], g["y%d" % k], g["d%d" % k]
        n1, n2 = x.shape[0], y.shape[0]
        if max(n1, n2) > ops.MAX_TOKEN_FRAMES:
            continue
        feat = np.concatenate([x, y]).astype(np.float32)
        pairs = np.array([[0, n1, n1, n2]], dtype=np.int32)
        dist, off, valid = _gpu_distance(feat, pairs)
        d = dist[:n1 * n2].reshape(n1, n2)
        assert valid[0] == 1
        # zero-norm rules (utils.py:55-58) are exact
        np.testing.assert_array_equal(d[(ref == 1.0) | (ref == 0.0)],
                                      ref[(ref == 1.0) | (ref == 0.0)])
        np.testing.assert_allclose(d, ref, rtol=0, atol=2e-6)


def test_cosine_distance_identical_frames_follow_nan_rule(golden_dir):
    # cos > 1 by rounding -> arccos NaN -> the reference asserts and the pair is
    # dropped (utils.py:59, dataloader.py:190-191).  Which identical frames
    # round above 1 depends on the summation order, so only the RULE is
    # checked: valid == 0 iff the GPU matrix holds a NaN.
    g = np.load(os.path.join(golden_dir, "cosine.npz"))
    x = g["x_identical"]
    n = x.shape[0]
    feat = np.concatenate([x, x]).astype(np.float32)
    dist, off, valid = _gpu_distance(feat, np.array([[0, n, n, n]], dtype=np.int32))
    d = dist[:n * n]
    assert bool(valid[0]) == (not np.isnan(d).any())
    off_diag = d.reshape(n, n)[~np.eye(n, dtype=bool)]
    assert not np.isnan(off_diag).any()


@pytest.fixture(scope="module")
def small_corpus():
    c = synth.make_corpus(600, cluster_size=8, tokens_per_file=150, seed=3)
    pairs = synth.make_same_pairs(c, 400, seed=4)
    return c, pairs


def test_align_pairs_end_to_end_vs_oracle(small_corpus):
    """get_dtw_alignment (utils.py:147-153) batched.  Bar (BASELINE north_star):
    costs within 1e-6 relative; paths bit-exact except where float32 rounding of
    the distance flips a near-tie -- every differing path is re-scored under the
    ORACLE's distance matrix and must be optimal to 1e-6 relative."""
    c, pairs = small_corpus
    feat = c.feat.numpy()
    res = ops.align_pairs(c.feat.to(DEV), pairs.to(DEV))
    torch.cuda.synchronize()
    idx1, idx2 = res.idx1.cpu().numpy(), res.idx2.cpu().numpy()
    off, plen = res.path_off.cpu().numpy(), res.path_len.cpu().numpy()
    cost, valid = res.cost.cpu().numpy(), res.valid.cpu().numpy()
    recs = oracle.align_pairs(feat, pairs.numpy())
    mismatched = 0
    for p, (r, tk) in enumerate(zip(recs, pairs.numpy().tolist())):
        assert bool(valid[p]) == r["valid"]
        s1, n1, s2, n2 = tk
        g1 = idx1[off[p]:off[p] + plen[p]] - s1
        g2 = idx2[off[p]:off[p] + plen[p]] - s2
        assert abs(cost[p] - r["cost"]) <= 1e-6 * r["cost"]
        if len(g1) == len(r["path1"]) and (g1 == r["path1"]).all() and (g2 == r["path2"]).all():
            continue
        mismatched += 1
        rescored = oracle.path_cost(r["dist"], g1, g2)
        assert abs(rescored - r["cost"]) <= 1e-6 * r["cost"]
    assert mismatched <= max(2, len(recs) // 50)


def test_align_pairs_is_bit_exact_given_its_own_distances(small_corpus):
    """Isolates the DTW stage of the FUSED kernel: feed the GPU's own float32
    distance matrices to the oracle DTW -> paths and costs must be identical."""
    c, pairs = small_corpus
    feat_d, pairs_d = c.feat.to(DEV), pairs.to(DEV)
    res = ops.align_pairs(feat_d, pairs_d)
    dist, doff, dvalid = ops.cosine_distance(feat_d, pairs_d)
    torch.cuda.synchronize()
    idx1, idx2 = res.idx1.cpu().numpy(), res.idx2.cpu().numpy()
    off, plen = res.path_off.cpu().numpy(), res.path_len.cpu().numpy()
    cost = res.cost.cpu().numpy()
    dist, doff = dist.cpu().numpy(), doff.cpu().numpy()
    for p, (s1, n1, s2, n2) in enumerate(pairs.numpy().tolist()):
        d = dist[doff[p]:doff[p + 1]].reshape(n1, n2).astype(np.float64)
        cst, q1, q2 = oracle.dtw(d)
        assert cost[p] == cst
        np.testing.assert_array_equal(idx1[off[p]:off[p] + plen[p]] - s1, q1)
        np.testing.assert_array_equal(idx2[off[p]:off[p] + plen[p]] - s2, q2)


def test_align_pairs_edge_shapes():
    rng = np.random.default_rng(2)
    feat = rng.standard_normal((400, 280)).astype(np.float32)
    feat[7] = 0.0                                       # zero-norm frame
    pairs = np.array([[0, 1, 1, 1], [0, 1, 10, 96], [10, 96, 0, 1], [100, 96, 200, 96],
                      [0, 20, 30, 20], [5, 0, 30, 20], [390, 20, 0, 20], [40, 17, 90, 33]],
                     dtype=np.int32)
    res = ops.align_pairs(torch.from_numpy(feat).to(DEV), torch.from_numpy(pairs).to(DEV))
    torch.cuda.synchronize()
    valid = res.valid.cpu().numpy()
    plen = res.path_len.cpu().numpy()
    assert valid[5] == 0 and plen[5] == 0               # empty token (s > e, dataloader.py:184)
    assert valid[6] == 0                                # token runs past the table
    recs = oracle.align_pairs(feat, pairs[[0, 1, 2, 3, 4, 7]])
    idx1, idx2, off = res.idx1.cpu().numpy(), res.idx2.cpu().numpy(), res.path_off.cpu().numpy()
    for p, r in zip([0, 1, 2, 3, 4, 7], recs):
        assert valid[p] == 1 and r["valid"]
        s1, n1, s2, n2 = pairs[p]
        g1 = idx1[off[p]:off[p] + plen[p]] - s1
        g2 = idx2[off[p]:off[p] + plen[p]] - s2
        assert g1[0] == 0 and g2[0] == 0 and g1[-1] == n1 - 1 and g2[-1] == n2 - 1
        rescored = oracle.path_cost(r["dist"], g1, g2)
        assert abs(rescored - r["cost"]) <= 1e-6 * max(r["cost"], 1e-30)


def test_token_longer_than_limit_is_refused():
    feat = torch.zeros((1300, 280), device=DEV)
    pairs = torch.tensor([[0, 600, 650, 50]], dtype=torch.int32, device=DEV)
    with pytest.raises(Exception):
        ops.align_pairs(feat, pairs)


def test_long_tokens_tiled_path_vs_oracle():
    """Tokens above 96 frames (config C5 sweeps 50-400; the reference's own pair file
    holds 98-frame tokens) take the tiled kernel: 96 x 96 distance tiles, DTW carried
    across tiles through boundary rows / columns.  Same bar as the fused path, plus:
    bit-exact against the oracle DTW run on the GPU's own distances."""
    from oracle.make_golden import smooth_tokens
    rng = np.random.default_rng(21)
    shapes = [(98, 56), (97, 97), (200, 130), (96, 193), (400, 400), (101, 30), (50, 50)]
    feats, pairs, row = [], [], 0
    for n1, n2 in shapes:
        x = smooth_tokens(rng, n1, 280)
        warp = np.minimum((np.arange(n2) * n1) // n2, n1 - 1)
        y = (0.75 * x[warp] + 0.5 * smooth_tokens(rng, n2, 280)).astype(np.float32)
        feats += [x, y]
        pairs.append([row, n1, row + n1, n2])
        row += n1 + n2
    feat = np.concatenate(feats)
    pairs = np.array(pairs, dtype=np.int32)
    fd, pd = torch.from_numpy(feat).to(DEV), torch.from_numpy(pairs).to(DEV)
    res = ops.align_pairs(fd, pd)
    dist, doff, dvalid = ops.cosine_distance(fd, pd)
    torch.cuda.synchronize()
    idx1, idx2 = res.idx1.cpu().numpy(), res.idx2.cpu().numpy()
    off, plen = res.path_off.cpu().numpy(), res.path_len.cpu().numpy()
    cost, valid = res.cost.cpu().numpy(), res.valid.cpu().numpy()
    dist, doff = dist.cpu().numpy(), doff.cpu().numpy()
    recs = oracle.align_pairs(feat, pairs)
    for p, (r, (s1, n1, s2, n2)) in enumerate(zip(recs, pairs.tolist())):
        assert valid[p] == 1 and r["valid"] and dvalid[p].item() == 1
        g1 = idx1[off[p]:off[p] + plen[p]] - s1
        g2 = idx2[off[p]:off[p] + plen[p]] - s2
        d_gpu = dist[doff[p]:doff[p + 1]].reshape(n1, n2).astype(np.float64)
        np.testing.assert_allclose(d_gpu, r["dist"], rtol=0, atol=2e-6)
        c, q1, q2 = oracle.dtw(d_gpu)                    # DTW stage alone: bit-exact
        assert cost[p] == c
        np.testing.assert_array_equal(g1, q1)
        np.testing.assert_array_equal(g2, q2)
        assert abs(cost[p] - r["cost"]) <= 1e-6 * r["cost"]     # end to end vs the oracle
        assert abs(oracle.path_cost(r["dist"], g1, g2) - r["cost"]) <= 1e-6 * r["cost"]


def test_long_tokens_on_a_stacked_table_vs_generic_kernels_and_oracle():
    """Config C5 shapes on a 7 x 40 stacked table: the long path's stacked tiles (90 x 90,
    40-deep Gram + 7-tap sums) return the same bits as its generic tiles (96 x 96, 280-deep),
    the band DTW is bit-exact against the oracle DTW on the GPU's own distances, and the whole
    thing meets the end-to-end bar -- in several windows of a small workspace, mixed with short
    pairs."""
    c = synth.make_corpus(60, cluster_size=4, tokens_per_file=7, len_range=(30, 420), seed=17)
    feat_d = c.feat.to(DEV)
    last = torch.zeros(c.feat.shape[0], dtype=torch.uint8)
    last[(c.file_off[1:] - 1).long()] = 1
    assert ops.stack_violations(feat_d, 7, last.to(DEV)) == 0
    pairs = synth.make_same_pairs(c, 40, seed=18).to(DEV)
    lens = pairs[:, [1, 3]].cpu().numpy()
    assert (lens > 256).any() and (lens > 96).sum() >= 20 and (lens.max(1) <= 90).any()
    g = ops.align_pairs(feat_d, pairs, stack=0)
    s = ops.align_pairs(feat_d, pairs, stack=7)
    dg, og, vg = ops.cosine_distance(feat_d, pairs, stack=0)
    ds, _, vs = ops.cosine_distance(feat_d, pairs, stack=7)
    torch.cuda.synchronize()
    assert torch.equal(dg.view(torch.int32), ds.view(torch.int32)) and torch.equal(vg, vs)
    assert int(g.valid.sum()) == pairs.shape[0] and torch.equal(g.valid, s.valid)
    assert torch.equal(g.path_len, s.path_len)
    assert torch.equal(g.cost.view(torch.int64), s.cost.view(torch.int64))
    plen, off = s.path_len.cpu().numpy(), s.path_off.cpu().numpy()
    g1, g2, s1, s2 = (t.cpu().numpy() for t in (g.idx1, g.idx2, s.idx1, s.idx2))
    dist, doff, cost = ds.cpu().numpy(), og.cpu().numpy(), s.cost.cpu().numpy()
    pn = pairs.cpu().numpy()
    feat = c.feat.numpy()
    for p, (r1, n1, r2, n2) in enumerate(pn.tolist()):
        sl = slice(off[p], off[p] + plen[p])
        np.testing.assert_array_equal(g1[sl], s1[sl])
        np.testing.assert_array_equal(g2[sl], s2[sl])
        d_gpu = dist[doff[p]:doff[p + 1]].reshape(n1, n2).astype(np.float64)
        cst, q1, q2 = oracle.dtw(d_gpu)                              # DTW stage alone: bit-exact
        assert cost[p] == cst
        np.testing.assert_array_equal(s1[sl] - r1, q1)
        np.testing.assert_array_equal(s2[sl] - r2, q2)
        if p < 12:                                                   # end to end vs the oracle
            d_ref = oracle.cosine_distance(feat[r1:r1 + n1], feat[r2:r2 + n2])
            np.testing.assert_allclose(d_gpu, d_ref, rtol=0, atol=2e-6)
            c_ref, _, _ = oracle.dtw(d_ref)
            assert abs(cost[p] - c_ref) <= 1e-6 * c_ref
    # a workspace with room for 3 long pairs at a time: many windows, same results
    P = pairs.shape[0]
    small = torch.empty(ops._lib.lib().abn_align_workspace_bytes(P, 420, (P + 2) // 3),
                        dtype=torch.uint8, device=DEV)
    path_len = torch.zeros(P, dtype=torch.int32, device=DEV)
    cost2 = torch.zeros(P, dtype=torch.float64, device=DEV)
    valid = torch.zeros(P, dtype=torch.uint8, device=DEV)
    i1, i2 = torch.empty_like(s.idx1), torch.empty_like(s.idx2)
    ops.check(ops._lib.lib().abn_align_pairs(
        ops.ptr(feat_d), feat_d.shape[0], 280, ops.ptr(pairs), P, 420, 7, ops.ptr(s.path_off),
        ops.ptr(i1), ops.ptr(i2), ops.ptr(path_len), ops.ptr(cost2), ops.ptr(valid), ops.ptr(small),
        small.numel(), ops.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(path_len, s.path_len) and torch.equal(cost2.view(torch.int64), s.cost.view(torch.int64))
    for p in range(P):
        sl = slice(off[p], off[p] + plen[p])
        assert torch.equal(i1[sl], s.idx1[sl]) and torch.equal(i2[sl], s.idx2[sl])


def test_diff_pairs_match_reference_row_selection():
    pairs = np.array([[100, 20, 500, 35], [100, 35, 500, 20], [10, 30, 700, 30],
                      [0, 1, 50, 9], [0, 0, 50, 9]], dtype=np.int32)
    for stretch in (False, True):
        i1, i2, off = ops.diff_pairs(torch.from_numpy(pairs).to(DEV), stretch=stretch)
        torch.cuda.synchronize()
        i1, i2, off = i1.cpu().numpy(), i2.cpu().numpy(), off.cpu().numpy()
        for p, (s1, n1, s2, n2) in enumerate(pairs.tolist()):
            if n1 <= 0 or n2 <= 0:
                assert off[p + 1] == off[p]
                continue
            (srcA, srcB), a, b, _ = oracle.diff_pair_indices(n1, n2, stretch)
            sa = s1 if srcA == 1 else s2
            sb = s1 if srcB == 1 else s2
            np.testing.assert_array_equal(i1[off[p]:off[p + 1]], sa + a)
            np.testing.assert_array_equal(i2[off[p]:off[p + 1]], sb + b)


def test_compact_and_gather(small_corpus):
    c, pairs = small_corpus
    feat_d = c.feat.to(DEV)
    res = ops.align_pairs(feat_d, pairs.to(DEV))
    d1, d2, doff = ops.compact_paths(res)
    torch.cuda.synchronize()
    plen, off = res.path_len.cpu().numpy(), res.path_off.cpu().numpy()
    ref1 = np.concatenate([res.idx1.cpu().numpy()[off[p]:off[p] + plen[p]] for p in range(len(plen))])
    np.testing.assert_array_equal(d1.cpu().numpy(), ref1)
    n = d1.numel()
    y = torch.ones(n, dtype=torch.int8, device=DEV)
    y[::3] = -1
    sel = torch.randperm(n, device=DEV)[:1000]
    x1, x2, yo = ops.gather_batch(feat_d, d1, d2, y, sel)
    torch.cuda.synchronize()
    assert torch.equal(x1, feat_d[d1[sel].long()]) and torch.equal(x2, feat_d[d2[sel].long()])
    assert torch.equal(yo, y[sel].float())


def test_stacked_fast_path_is_bit_identical_to_generic_kernels():
    """on a 7x40 stacked table the fast path (40-deep Gram tile + 7-tap
    diagonal sums) must return the SAME BITS as the generic 280-deep kernels:
    distances, costs, paths -- including tokens at file edges (zero-padded context)
    and tokens of 91-96 frames, which the stacked mode routes to the tiled kernel."""
    c = synth.make_corpus(500, cluster_size=8, tokens_per_file=60, len_range=(1, 96), seed=13)
    feat_d = c.feat.to(DEV)
    last = torch.zeros(c.feat.shape[0], dtype=torch.uint8)
    last[(c.file_off[1:] - 1).long()] = 1
    assert ops.stack_violations(feat_d, 7, last.to(DEV)) == 0
    assert ops.stack_violations(feat_d, 7) > 0            # file boundaries break the overlap
    assert ops.stack_violations(torch.randn(100, 280, device=DEV), 7) == 99
    pairs = synth.make_same_pairs(c, 600, seed=14)
    # force file-edge tokens and long tokens into the list
    starts, lens = c.tok_start, c.tok_len
    edge = [0, 59, 60, 119, 120, int(lens.numel()) - 1]
    extra = torch.tensor([[int(starts[a]), int(lens[a]), int(starts[b]), int(lens[b])]
                          for a, b in zip(edge, reversed(edge))], dtype=torch.int32)
    longs = torch.nonzero(lens > 90).squeeze(1)[:6]
    assert longs.numel() >= 2
    extra2 = torch.tensor([[int(starts[a]), int(lens[a]), int(starts[b]), int(lens[b])]
                           for a, b in zip(longs.tolist(), reversed(longs.tolist()))],
                          dtype=torch.int32)
    pairs = torch.cat([pairs, extra, extra2]).contiguous().to(DEV)
    g = ops.align_pairs(feat_d, pairs, stack=0)
    s = ops.align_pairs(feat_d, pairs, stack=7)
    dg, og, vg = ops.cosine_distance(feat_d, pairs, stack=0)
    ds, os_, vs = ops.cosine_distance(feat_d, pairs, stack=7)
    torch.cuda.synchronize()
    assert torch.equal(vg, vs) and torch.equal(og, os_)
    assert torch.equal(dg.view(torch.int32), ds.view(torch.int32))          # same bits
    assert torch.equal(g.valid, s.valid) and torch.equal(g.path_len, s.path_len)
    assert torch.equal(g.cost.view(torch.int64), s.cost.view(torch.int64))
    plen, off = g.path_len.cpu().numpy(), g.path_off.cpu().numpy()
    g1, g2, s1, s2 = (t.cpu().numpy() for t in (g.idx1, g.idx2, s.idx1, s.idx2))
    for p in range(len(plen)):
        sl = slice(off[p], off[p] + plen[p])
        np.testing.assert_array_equal(g1[sl], s1[sl])
        np.testing.assert_array_equal(g2[sl], s2[sl])
    assert int(g.valid.sum()) == pairs.shape[0]


def test_stacked_mode_refuses_other_shapes():
    feat = torch.randn(200, 40, device=DEV)
    pairs = torch.tensor([[0, 30, 50, 40]], dtype=torch.int32, device=DEV)
    with pytest.raises(Exception):
        ops.align_pairs(feat, pairs, stack=7)


def test_stack_upload_rebuilds_the_table_from_its_middle_blocks():
    """abn_stack_upload: 7x fewer PCIe bytes, device table identical to the host table."""
    corpus = synth.make_corpus(120, cluster_size=8, tokens_per_file=40, seed=3)
    host = corpus.feat.contiguous().pin_memory()
    last = torch.zeros(host.shape[0], dtype=torch.uint8)
    last[(corpus.file_off[1:] - 1).long()] = 1
    dev_t = ops.stack_upload(host, 7, last.to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(dev_t.cpu(), host)
    # and end to end through the host-buffer call
    pairs = synth.make_same_pairs(corpus, 50, seed=4)
    a = utils.align_pairs_host(host, pairs, stack=7, last_row_of_file=last)
    b = utils.align_pairs_host(host, pairs, stack=0)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    # chunked (copy-back of chunk i overlaps the alignment of chunk i + 1): same result
    c = utils.align_pairs_host(host, pairs, stack=7, last_row_of_file=last, chunks=3)
    a = [t.clone() for t in utils.align_pairs_host(host, pairs, stack=7, last_row_of_file=last, chunks=1)]
    for x, y in zip(a, c):
        assert torch.equal(x, y)


def test_alignment_invariants_at_scale():
    """Size-independent properties on a large pair list (BASELINE C2 shapes, 200 k pairs):
    every path starts at (0, 0), ends at (n1-1, n2-1), moves by (1,0) / (0,1) / (1,1), has
    max(n1, n2) <= L <= n1 + n2 - 1; the cost is the float64 sum of the kernel's own
    distances along the path; the stacked and the generic kernels agree bit for bit; a
    second run returns the same bits (no run-to-run variation)."""
    corpus = synth.make_corpus(8000, seed=11, device=DEV)
    pairs = synth.make_same_pairs(corpus, 200_000, seed=12)
    feat = corpus.feat
    res = ops.align_pairs(feat, pairs, stack=0)
    d1, d2, off = ops.compact_paths(res)
    torch.cuda.synchronize()
    assert bool((res.valid == 1).all())
    L = res.path_len.long()
    n1, n2 = pairs[:, 1].long(), pairs[:, 3].long()
    assert bool((L >= torch.maximum(n1, n2)).all()) and bool((L <= n1 + n2 - 1).all())
    first, last = off[:-1], off[1:] - 1
    assert torch.equal(d1[first].long(), pairs[:, 0].long()) and torch.equal(d2[first].long(), pairs[:, 2].long())
    assert torch.equal(d1[last].long(), pairs[:, 0].long() + n1 - 1)
    assert torch.equal(d2[last].long(), pairs[:, 2].long() + n2 - 1)
    di, dj = d1[1:] - d1[:-1], d2[1:] - d2[:-1]
    inner = torch.ones(d1.numel() - 1, dtype=torch.bool, device=DEV)
    inner[last[:-1]] = False                      # steps across pair boundaries do not count
    assert bool((((di == 0) | (di == 1)) & ((dj == 0) | (dj == 1)) & (di + dj >= 1))[inner].all())
    # stacked fast path and a second run: same bits
    res7 = ops.align_pairs(feat, pairs, stack=7)
    e1, e2, _ = ops.compact_paths(res7)
    assert torch.equal(res7.cost.view(torch.int64), res.cost.view(torch.int64))
    assert torch.equal(e1, d1) and torch.equal(e2, d2)
    res_b = ops.align_pairs(feat, pairs, stack=7)
    assert torch.equal(res_b.cost.view(torch.int64), res7.cost.view(torch.int64))
    # cost == float64 sum of the kernel's own distances along the path (first 500 pairs)
    sub = pairs[:500].contiguous()
    dist, doff, _ = ops.cosine_distance(feat, sub)
    dist, doff = dist.double().cpu().numpy(), doff.cpu().numpy()
    offc, d1c, d2c = off.cpu().numpy(), d1.cpu().numpy(), d2.cpu().numpy()
    cost = res.cost.cpu().numpy()
    for p, (s1, a, s2, b) in enumerate(sub.cpu().numpy().tolist()):
        i = d1c[offc[p]:offc[p + 1]] - s1
        j = d2c[offc[p]:offc[p + 1]] - s2
        acc = 0.0
        for v in dist[doff[p] + i * b + j]:
            acc += v
        assert acc == cost[p], (p, acc, cost[p])


def test_unstacked_input_and_direction_stream_through_the_host_call():
    """align_pairs_host(frames_host=[N, 40] frames, last_row_of_file, paths='directions'): the
    un-stacked frames are stacked on the device exactly like abnet3/features.py:135-159 (the
    table equals the stacked one bit for bit), and the 2-bit direction stream decodes to the
    index pairs of the plain call."""
    from abnet3_b200 import utils
    c = synth.make_corpus(300, cluster_size=8, tokens_per_file=40, seed=23)
    last = torch.zeros(c.feat.shape[0], dtype=torch.uint8)
    last[(c.file_off[1:] - 1).long()] = 1
    frames = c.feat[:, 120:160].contiguous()
    built = ops.stack_from_frames(frames, 7, last.to(DEV))
    assert torch.equal(built.cpu(), c.feat)
    tab = utils.FeatureTable.from_frames(frames.pin_memory(), c.file_off)
    assert tab.stack == 7 and torch.equal(tab.feat.cpu(), c.feat)
    pairs = synth.make_same_pairs(c, 500, seed=24)
    pairs[7, 1] = 0                                            # a skipped pair (s > e)
    ref = utils.align_pairs_host(c.feat, pairs, chunks=3)
    ref = [t.clone() for t in ref]
    got = utils.align_pairs_host(None, pairs, frames_host=frames, last_row_of_file=last,
                                 paths="directions", chunks=3)
    dirs, dir_off, plen, cost, valid = got
    assert torch.equal(plen, ref[3]) and torch.equal(valid, ref[5])
    assert torch.equal(cost[valid.bool()], ref[4][valid.bool()])
    assert dirs.numel() * 16 < ref[0].numel() * 8 + 16 * 500      # 2 bits instead of 64 per step
    i1, i2, off = utils.decode_directions(dirs.numpy(), dir_off.numpy(), plen.numpy(), pairs.numpy())
    np.testing.assert_array_equal(off, ref[2].numpy())
    np.testing.assert_array_equal(i1, ref[0].numpy())
    np.testing.assert_array_equal(i2, ref[1].numpy())
