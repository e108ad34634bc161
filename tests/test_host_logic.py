"""CPU: host-side logic of the reference surface (pair parsing, token slicing,
splits, sharding, step agreement over gloo) -- no kernel is called."""
import os

import numpy as np
import pytest
import torch

from abnet3_b200 import utils as U
from abnet3_b200 import dataloader as DL
from abnet3_b200 import trainer as TR
from oracle.align import FeaturesAccessor, diff_pair_indices


def test_read_dataset_group_pairs(tmp_path):
    p = tmp_path / "dataset"
    p.write_text("f1 0.10 0.55 f2 1.00 1.42 same\nf1 0.10 0.55 f3 2.00 2.30 diff\n")
    pairs = U.read_dataset(str(p))
    assert pairs == [("f1", 0.10, 0.55, "f2", 1.00, 1.42, "same"),
                     ("f1", 0.10, 0.55, "f3", 2.00, 2.30, "diff")]
    g = U.group_pairs(pairs)
    assert g["same"] == [("f1", 0.10, 0.55, "f2", 1.00, 1.42)] and len(g["diff"]) == 1
    (tmp_path / "bad").write_text("f1 0 1 f2 0 1 other\n")
    with pytest.raises(AssertionError):
        U.read_dataset(str(tmp_path / "bad"))


def test_spkid_file(tmp_path):
    p = tmp_path / "spk"
    p.write_text("f1 A\nf2 A\nf3 B\n")
    spk = U.read_spkid_file(str(p))
    assert spk == {"f1": "A", "f2": "A", "f3": "B"}


class _HostTable(U.FeatureTable):
    """FeatureTable without the device copy (CPU test only)."""

    def __init__(self, features, times=None):
        super().__init__(features, times, device=torch.device("cpu"))


def test_token_slicing_matches_reference_accessor():
    rng = np.random.default_rng(0)
    feats = {"a": rng.standard_normal((500, 8)).astype(np.float32),
             "b": rng.standard_normal((300, 8)).astype(np.float32)}
    times = {k: 0.0025 + 0.01 * np.arange(v.shape[0]) for k, v in feats.items()}
    table = _HostTable(feats, times)
    acc = FeaturesAccessor(times, feats)
    for f, on, off in [("a", 0.0, 0.5), ("a", 1.0025, 1.5025), ("b", 2.9, 9.0), ("b", 0.5, 0.4),
                       ("a", 0.0125, 0.0125), ("b", 1.234, 1.777)]:
        s, n = table.token_by_time(f, on, off)
        ref = acc.get(f, on, off)
        assert n == ref.shape[0]
        np.testing.assert_array_equal(table.host[s:s + n], ref)
    vec = table.tokens_by_time(["a", "b", "a"], [0.0, 0.5, 1.0025], [0.5, 0.4, 1.5025])
    assert vec[:, 1].tolist() == [table.token_by_time("a", 0.0, 0.5)[1], 0,
                                  table.token_by_time("a", 1.0025, 1.5025)[1]]
    for f, a, b in [("a", 10, 60), ("b", 290, 400), ("a", 50, 40), ("b", -5, 7)]:
        s, n = table.token_by_frames(f, a, b)
        np.testing.assert_array_equal(table.host[s:s + n], acc.get_between_frames(f, a, b))


def test_diff_rows_match_reference_selection():
    for n1, n2 in [(20, 35), (35, 20), (30, 30), (1, 9), (9, 1)]:
        for stretch in (False, True):
            (srcA, srcB), a, b, nlab = diff_pair_indices(n1, n2, stretch)
            r1, r2, nl = DL._diff_rows((100, n1, 500, n2), stretch)
            sa = 100 if srcA == 1 else 500
            sb = 100 if srcB == 1 else 500
            np.testing.assert_array_equal(r1, sa + a)
            np.testing.assert_array_equal(r2, sb + b)
            assert nl == nlab == min(n1, n2)


def test_pairs_loader_split_on_reference_fixture(golden_dir):
    # the reference's own pair file (test/data/dataloader/pairs_knn.txt)
    loader = DL.PairsDataLoader(os.path.join(golden_dir, "pairs_knn.txt"), None,
                                os.path.join(golden_dir, "id_to_file.txt"),
                                ratio_split_train_test=0.7, train_iterations=2,
                                test_iterations=2)
    loader.load_pairs()
    allp = loader.pairs["train"] + loader.pairs["test"]
    assert all(len(p) == 6 for p in allp)
    for p in allp:
        assert p[0] in ["file%d" % i for i in range(5)] and p[3] in ["file%d" % i for i in range(5)]
    # split_each_file (dataloader.py:484-508): thresholds at 70 % of each file's last frame
    last = {}
    for line in open(os.path.join(golden_dir, "pairs_knn.txt")):
        f1, f2, b1, e1, b2, e2, _ = line.split(" ")
        last["file" + f1] = max(last.get("file" + f1, 0), int(e1))
        last["file" + f2] = max(last.get("file" + f2, 0), int(e2))
    for f1, s1, e1, f2, s2, e2 in loader.pairs["train"]:
        assert s1 < 0.7 * last[f1] and s2 <= 0.7 * last[f2]
    for f1, s1, e1, f2, s2, e2 in loader.pairs["test"]:
        assert s1 > 0.7 * last[f1] and s2 > 0.7 * last[f2]
    files_split = DL.PairsDataLoader(os.path.join(golden_dir, "pairs_knn.txt"), None, None,
                                     split_method=DL.PairsDataLoader.SPLIT_FILES)
    files_split.load_pairs()
    dev_files = {p[0] for p in files_split.pairs["test"]} | {p[3] for p in files_split.pairs["test"]}
    trn_files = {p[0] for p in files_split.pairs["train"]} | {p[3] for p in files_split.pairs["train"]}
    assert not (dev_files & trn_files)


def test_shard_pairs_partitions_the_list():
    pairs = list(range(103))
    shards = [TR.shard_pairs(pairs, r, 4) for r in range(4)]
    assert sorted(sum(shards, [])) == pairs
    assert max(map(len, shards)) - min(map(len, shards)) <= 1


def _gloo_worker(rank, world, port, n_batches, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank owns a shard with a different number of batches: everybody must
    # stop after the SHORTEST shard, and a flat "gradient" all-reduce sums up
    steps = 0
    it = iter(range(n_batches[rank]))
    bucket_sum = 0.0
    while True:
        b = next(it, None)
        if not TR.all_ranks_have_batch(b is not None, "cpu", world):
            break
        g = torch.full((8,), float(rank + 1))
        dist.all_reduce(g)
        bucket_sum += float(g[0])
        steps += 1
    # per-epoch agreement (what the FramesDataLoader sweep uses) and the torch.optim fallback's
    # gradient reduction: mean for an averaged loss, sum otherwise
    agreed = TR.agree_on_batches(n_batches[rank], "cpu", world)
    w = torch.nn.Parameter(torch.zeros(4))
    w.grad = torch.full((4,), float(rank + 1))
    TR.reduce_grads([w], True, world)
    mean_g = float(w.grad[0])
    w.grad = torch.full((4,), float(rank + 1))
    TR.reduce_grads([w], False, world)
    out.put((rank, steps, bucket_sum, agreed, mean_g, float(w.grad[0])))
    dist.destroy_process_group()


def test_step_agreement_and_gradient_allreduce_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, n_batches = 2, [5, 3]
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n_batches, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [3, 3]                   # shortest shard decides
    assert [r[2] for r in res] == [9.0, 9.0]               # (1 + 2) summed over 3 steps
    assert [r[3] for r in res] == [3, 3]                   # one MIN reduction per epoch
    assert [r[4] for r in res] == [1.5, 1.5] and [r[5] for r in res] == [3.0, 3.0]


def test_dataloader_shards_partition_the_pair_lists(tmp_path):
    """dataloader.shard(rank, world): every rank keeps its own part of the train and dev lists
    (what the trainer applies under torchrun)."""
    from abnet3_b200.dataloader import OriginalDataLoader
    lines = ["f%d 0.10 0.30 g%d 0.20 0.50 %s" % (k, k, "same" if k % 2 else "diff") for k in range(11)]
    for mode in ("train_pairs", "dev_pairs"):
        os.makedirs(tmp_path / mode)
        (tmp_path / mode / "dataset").write_text("\n".join(lines) + "\n")
    seen = []
    for rank in range(3):
        dl = OriginalDataLoader(str(tmp_path), None)
        dl.features = object()                      # (only the pair lists are loaded here)
        dl.shard(rank, 3)
        dl.load_data()
        seen += [p[0] for p in dl.pairs["train"]]
        assert len(dl.pairs["dev"]) in (3, 4)
        with pytest.raises(RuntimeError):
            dl.shard((rank + 1) % 3, 3)             # too late: the lists are loaded
    assert sorted(seen) == sorted("f%d" % k for k in range(11))


def test_embedder_surface_and_feature_archives(tmp_path):
    """abnet3/embedder.py:37-51 constructor contract, the .npz feature container round trip and the
    loud failure without a GPU (no CPU path)."""
    import numpy as np
    import pytest
    import torch
    from abnet3_b200.embedder import (EmbedderBuilder, EmbedderSiamese, _load_features, _write_features)
    from abnet3_b200.model import SiameseNetwork
    with pytest.raises(ValueError):
        EmbedderSiamese(network=None)
    net = SiameseNetwork(input_dim=8, num_hidden_layers=1, hidden_dim=16, output_dim=4, p_dropout=0.0,
                         activation_layer="sigmoid")
    e = EmbedderSiamese(network=net, feature_path={"a": np.zeros((3, 8), np.float32)}, cuda=False)
    assert e.batch_size == 5000 and e.output_path is None
    with pytest.raises(NotImplementedError):
        EmbedderBuilder(network=net).embed()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            e.embed()
    items = ["s01_a", "s02_b"]
    embs = [np.arange(12, dtype=np.float32).reshape(3, 4), np.ones((2, 4), np.float32)]
    times = [0.0025 + 0.01 * np.arange(3), 0.0025 + 0.01 * np.arange(2)]
    path = str(tmp_path / "out.npz")
    _write_features(path, items, times, embs)
    got_items, got_times, got_feats = _load_features(path)
    assert got_items == items
    for a, b in zip(got_feats, embs):
        np.testing.assert_array_equal(a, b)
    for a, b in zip(got_times, times):
        np.testing.assert_allclose(a, b)


def test_direction_stream_decoder_restores_index_pairs():
    """utils.decode_directions: the host side of align_pairs_host(paths='directions')."""
    from abnet3_b200.utils import decode_directions
    rng = np.random.default_rng(0)
    P = 60
    tok = np.zeros((P, 4), np.int64)
    paths, dirs, dir_off, L = [], [], [0], []
    for p in range(P):
        n1, n2 = rng.integers(1, 30, 2)
        tok[p] = [rng.integers(0, 1000), n1, rng.integers(0, 1000), n2]
        if p % 7 == 3:                               # a dropped pair: empty path
            L.append(0)
            dir_off.append(dir_off[-1])
            paths.append((np.zeros(0, int), np.zeros(0, int)))
            continue
        i = j = 0
        pi, pj, ds = [0], [0], []
        while (i, j) != (n1 - 1, n2 - 1):
            opts = ([0] if i < n1 - 1 and j < n2 - 1 else []) + ([1] if i < n1 - 1 else []) + \
                   ([2] if j < n2 - 1 else [])
            d = rng.choice(opts)
            ds.append(d)
            i += d != 2
            j += d != 1
            pi.append(i)
            pj.append(j)
        L.append(len(pi))
        paths.append((np.array(pi) + tok[p, 0], np.array(pj) + tok[p, 2]))
        b = np.zeros((len(ds) + 3) // 4, np.uint8)
        for k, d in enumerate(ds):
            b[k // 4] |= d << (2 * (k % 4))
        dirs.append(b)
        dir_off.append(dir_off[-1] + len(b))
    i1, i2, off = decode_directions(np.concatenate(dirs), np.array(dir_off), np.array(L), tok)
    for p in range(P):
        np.testing.assert_array_equal(i1[off[p]:off[p + 1]], paths[p][0])
        np.testing.assert_array_equal(i2[off[p]:off[p + 1]], paths[p][1])


def test_the_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no module of the package (nor the C sources) may import,
    include or execute anything under it -- only tests/, __graft_entry__.smoke() and bench.py's CPU
    legs may."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "abnet3_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                continue
            text = open(os.path.join(dirpath, f), errors="ignore").read()
            if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or \
                    re.search(r"#include\s+[\"<][^\">]*oracle", text) or "import_module(\"oracle" in text:
                offenders.append(os.path.relpath(os.path.join(dirpath, f), root))
    assert offenders == []
    # bench.py: only inside the CPU-baseline / reference-arm functions
    bench = open(os.path.join(root, "bench.py")).read()
    for m in re.finditer(r"^(\s*)import oracle\b|^(\s*)from oracle\b", bench, flags=re.M):
        assert len((m.group(1) or m.group(2))) > 0, "bench.py imports the oracle at module level"
