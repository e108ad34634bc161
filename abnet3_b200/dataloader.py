"""Pair alignment and batch generation of /root/reference/abnet3/dataloader.py,
device resident.

Same class names, constructor arguments, ``load_data`` / ``batch_iterator`` /
``whoami`` / ``statistics_training`` / pickling behaviour as the reference:

* ``OriginalDataLoader``   (:43-352)   batches of ``batch_size`` token pairs
* ``PairsDataLoader``      (:355-546)  frame-indexed knn pair file
* ``FramesDataLoader``     (:580-739)  all pairs aligned once, frame batches
* ``MultiTaskDataLoader``  (:742-792)  adds speaker labels
* ``MultiTaskFramesDataLoader``        (new) frame batches with speaker labels

What changed underneath: the corpus lives on the GPU as one [n_rows, dim] table
(utils.FeatureTable); every 'same' pair of a pair list is aligned in ONE batched
launch of the fused cosine-distance + DTW + traceback kernel (the reference
re-runs DTW pair by pair, every epoch); batches are index lists into that
table and the rows are gathered on the device, so the tensors a batch iterator
yields are CUDA tensors (the trainer's ``.cuda()`` is a no-op).  ``X1`` and
``X2`` are the two halves of one buffer, which lets the network run both
siamese branches as a single batch without a copy.

Reference quirks are reproduced, not fixed: the same numpy
permutation seed for every batch, speaker labels by object identity
(``spk1 is spk2``, i.e. "same file"), and the row/label length mismatch of
``align_different_words=True``.  Labels are yielded as float32 (+1 / -1); the
reference yields float64 / int64 with the same values.

Temporal-coherence batches (``tcl > 0``, dataloader.py:314-352) are built in index
space with the reference's ``random`` draws.  Out of scope:
``MultimodalDataLoader``.
"""
import os
import random
from collections import defaultdict

import numpy as np
import torch

from . import ops
from .utils import (read_feats, read_dataset, group_pairs, read_spkid_file,
                    BatchAligner)


class DataLoader:

    def batch_iterator(self, train_mode=True):
        raise NotImplementedError("You must implement batch iterator in DataLoader class.")

    def whoami(self):
        raise NotImplementedError("You must implement whoami in DataLoader class")


class _Aligned(object):
    """Alignment of the 'same' pairs of one pair list, kept on the host as a
    compact table: pair p's aligned global rows are
    idx1[off[p]:off[p+1]] / idx2[...]; valid[p] tells whether the reference
    would have kept the pair."""

    def __init__(self, tok, feat, max_frames, stack=0):
        self.tok = tok
        P = tok.shape[0]
        if P == 0:
            self.idx1 = self.idx2 = np.zeros(0, np.int32)
            self.off = np.zeros(1, np.int64)
            self.valid = np.zeros(0, bool)
            return
        tok_d = torch.from_numpy(tok).to(feat.device)
        res = ops.align_pairs(feat, tok_d, max_frames=max_frames, stack=stack)
        d1, d2, doff = ops.compact_paths(res)
        self.idx1, self.idx2 = d1.cpu().numpy(), d2.cpu().numpy()
        self.off = doff.cpu().numpy()
        self.valid = res.valid.cpu().numpy().astype(bool)
        self.dev = (d1, d2, doff, res.valid)


def _diff_rows(tok, stretch):
    """Row selection of one 'diff' pair on the host (dataloader.py:208-231):
    -> (rows1, rows2, n_labels)."""
    s1, n1, s2, n2 = (int(v) for v in tok)
    if stretch:
        smax, smin = (s1 if n1 >= n2 else s2), (s1 if n1 <= n2 else s2)
        lmax, lmin = max(n1, n2), min(n1, n2)
        mapping = np.rint(np.linspace(0, lmin - 1, num=lmax)).astype(np.int64)
        return smax + np.arange(lmax), smin + mapping, min(n1, n2)
    m = min(n1, n2)
    return s1 + np.arange(m), s2 + np.arange(m), m


class OriginalDataLoader(DataLoader):
    """abnet3/dataloader.py:43-352."""

    TCL_DISTANCE_SAME = [1]  # abnet3/dataloader.py:51-52 (Synnaeve & Dupoux temporal coherence loss)
    TCL_DISTANCES_DIFF = [15, 20, 25, 30]

    def __init__(self, pairs_path, features_path, num_max_minibatches=1000,
                 seed=None, batch_size=8, shuffle_between_epochs=False,
                 align_different_words=False,
                 tcl=0.0):
        assert 0 <= tcl < 1
        self.pairs_path = pairs_path
        self.features_path = features_path
        self.statistics_training = defaultdict(int)
        self.seed = seed
        self.num_max_minibatches = num_max_minibatches
        self.batch_size = batch_size
        self.features = None
        self.shuffle_between_epochs = shuffle_between_epochs
        self.align_different_words = align_different_words
        self.tcl = tcl
        self.train_files = None
        self.pairs = {'train': None, 'dev': None}
        self._cache = {}
        self._shard = None

    def shard(self, rank, world_size):
        """Data parallelism (new): this process keeps every ``world_size``-th pair of the
        train and dev lists, starting at ``rank`` (pairs are independent: DTW alignment needs
        no communication).  Call before ``load_data``; the trainer does so under torchrun."""
        if self.pairs['train'] is not None and self._shard != (rank, world_size):
            raise RuntimeError("shard() must be called before the pairs are loaded")
        self._shard = (int(rank), int(world_size))

    def _sharded(self, pairs):
        if self._shard is None or self._shard[1] <= 1:
            return pairs
        return pairs[self._shard[0]::self._shard[1]]

    # pickling without the features (dataloader.py:86-117)
    def __getstate__(self):
        return (self.pairs_path,
                self.features_path,
                self.statistics_training,
                self.seed,
                self.num_max_minibatches,
                self.batch_size)

    def __setstate__(self, state):
        (self.pairs_path, self.features_path, self.statistics_training, self.seed,
         self.num_max_minibatches, self.batch_size) = state
        self.features = None
        self.shuffle_between_epochs = False
        self.align_different_words = False
        self.tcl = 0.0
        self.train_files = None
        self.pairs = {'train': None, 'dev': None}
        self._cache = {}
        self._shard = None
        self.load_data()

    def whoami(self):
        return {
            'params': self.__getstate__(),
            'class_name': self.__class__.__name__
        }

    def load_data(self):
        """Load the features and the pairs once (dataloader.py:119-145)."""
        if self.features is None:
            print("Loading features")
            features, _, _ = read_feats(self.features_path)
            self.features = features
        if self.pairs['train'] is None:
            print("Loading word pairs")
            self.pairs['train'] = self._sharded(read_dataset(
                os.path.join(self.pairs_path, 'train_pairs/dataset')))
        if self.pairs['dev'] is None:
            self.pairs['dev'] = self._sharded(read_dataset(
                os.path.join(self.pairs_path, 'dev_pairs/dataset')))
        self.train_files = list({pair[0] for pair in self.pairs['train']} |
                                {pair[3] for pair in self.pairs['train']})

    # ---------------------------------------------------------------- tokens
    @property
    def table(self):
        return self.features.table

    def _tokens(self, plist, frames=False):
        """[(f1,s1,e1,f2,s2,e2), ...] -> int32 [P, 4] (row1, n1, row2, n2); a
        pair the reference skips (s > e, dataloader.py:184) gets n = 0."""
        P = len(plist)
        tok = np.zeros((P, 4), dtype=np.int32)
        if P == 0:
            return tok
        if frames:
            for k, (f1, s1, e1, f2, s2, e2) in enumerate(plist):
                tok[k, 0:2] = self.table.token_by_frames(f1, s1, e1)
                tok[k, 2:4] = self.table.token_by_frames(f2, s2, e2)
        else:
            f1 = [p[0] for p in plist]
            f2 = [p[3] for p in plist]
            tok[:, 0:2] = self.table.tokens_by_time(f1, [p[1] for p in plist], [p[2] for p in plist])
            tok[:, 2:4] = self.table.tokens_by_time(f2, [p[4] for p in plist], [p[5] for p in plist])
        skip = np.array([(p[1] > p[2]) or (p[4] > p[5]) for p in plist])
        tok[skip, 1] = 0
        tok[skip, 3] = 0
        return tok

    def get_token_feats(self, pairs, frames=False):
        """dataloader.py:147-164 (host rows, for inspection / the oracle)."""
        get = self.features.get_between_frames if frames else self.features.get
        token_feats = {}
        for kind in ('same', 'diff'):
            for f1, s1, e1, f2, s2, e2 in pairs[kind]:
                if (f1, s1, e1) not in token_feats:
                    token_feats[f1, s1, e1] = get(f1, s1, e1)
                if (f2, s2, e2) not in token_feats:
                    token_feats[f2, s2, e2] = get(f2, s2, e2)
        return token_feats

    def _align(self, tok):
        longest = int(tok[:, [1, 3]].max()) if len(tok) else 1
        return _Aligned(tok, self.table.feat, max(longest, 1), stack=self.table.stack)

    # ------------------------------------------------------------- batching
    def _assemble(self, same_plist, same_tok, same_al, same_ids, diff_plist, diff_tok,
                  seed=0, fid2spk=None):
        """Index-space version of dataloader.py:183-259 for ONE batch: returns
        device tensors (X1, X2, y_phn) or (X1, X2, y_spk, y_phn)."""
        rows1, rows2, y_phn, y_spk = [], [], [], []
        for k, pid in enumerate(same_ids):
            tk = same_tok[pid]
            if tk[1] <= 0 or tk[3] <= 0 or not same_al.valid[pid]:
                continue                       # s > e, or the DTW "exception" (:184, :188-191)
            a, b = same_al.off[pid], same_al.off[pid + 1]
            self.statistics_training['SameType'] += 1
            if fid2spk:
                f1, f2 = same_plist[k][0], same_plist[k][3]
                if fid2spk[f1] is fid2spk[f2]:             # quirk q2: object identity
                    y_spk.append(np.ones(b - a))
                    self.statistics_training['SameTypeSameSpk'] += 1
                else:
                    y_spk.append(-1 * np.ones(b - a))
                    self.statistics_training['SameTypeDiffSpk'] += 1
            rows1.append(same_al.idx1[a:b])
            rows2.append(same_al.idx2[a:b])
            y_phn.append(np.ones(b - a))
        for k, tk in enumerate(diff_tok):
            if tk[1] <= 0 or tk[3] <= 0:
                if diff_plist[k][1] > diff_plist[k][2] or diff_plist[k][4] > diff_plist[k][5]:
                    continue
            r1, r2, nlab = _diff_rows(tk, self.align_different_words)
            rows1.append(r1)
            rows2.append(r2)
            y_phn.append(-1 * np.ones(nlab))
            self.statistics_training['DiffType'] += 1
            if fid2spk:
                f1, f2 = diff_plist[k][0], diff_plist[k][3]
                if fid2spk[f1] is fid2spk[f2]:
                    y_spk.append(np.ones(nlab))
                    self.statistics_training['DiffTypeSameSpk'] += 1
                else:
                    y_spk.append(-1 * np.ones(nlab))
                    self.statistics_training['DiffTypeDiffSpk'] += 1
        if fid2spk:
            assert len(y_phn) == len(y_spk), 'not same number of labels...'
        rows1, rows2 = np.concatenate(rows1), np.concatenate(rows2)     # np.vstack, :247
        y_phn = np.concatenate(y_phn)
        np.random.seed(seed)                                           # :248, quirk q1
        n_pairs = len(y_phn)
        ind = np.random.permutation(n_pairs)
        rows1, rows2 = rows1[ind], rows2[ind]       # rows beyond n_pairs are dropped (quirk q3)
        y_phn = y_phn[ind]
        dev = self.table.feat.device
        i1 = torch.from_numpy(rows1.astype(np.int32)).to(dev)
        i2 = torch.from_numpy(rows2.astype(np.int32)).to(dev)
        buf = torch.empty((2 * n_pairs, self.table.dim), dtype=torch.float32, device=dev)
        yo = torch.empty(n_pairs, dtype=torch.float32, device=dev)
        X1, X2 = buf[:n_pairs], buf[n_pairs:]
        ops.gather_batch(self.table.feat, i1, i2, None, None, n_pairs, out=(X1, X2, yo))
        yp = torch.from_numpy(y_phn.astype(np.float32)).to(dev)
        if fid2spk:
            ys = torch.from_numpy(np.concatenate(y_spk)[ind].astype(np.float32)).to(dev)
            return X1, X2, ys, yp
        return X1, X2, yp

    def load_frames_from_pairs(self, pairs, seed=0, fid2spk=None, frames=False):
        """dataloader.py:166-261 for one grouped batch {'same': [...], 'diff': [...]}."""
        same_tok = self._tokens(pairs['same'], frames)
        diff_tok = self._tokens(pairs['diff'], frames)
        al = self._align(same_tok)
        return self._assemble(pairs['same'], same_tok, al, range(len(pairs['same'])),
                              pairs['diff'], diff_tok, seed, fid2spk)

    def _prepared(self, mode, pairs, frames=False):
        """Tokens + alignment of the whole pair list of `mode`, computed once:
        DTW results do not depend on how pairs are batched."""
        key = (mode, id(pairs), len(pairs))
        if self._cache.get('key_' + mode) != key:
            same_pos = [i for i, p in enumerate(pairs) if p[6] == 'same']
            same_plist = [tuple(pairs[i][:6]) for i in same_pos]
            tok = self._tokens(same_plist, frames)
            self._cache[mode] = (dict((i, k) for k, i in enumerate(same_pos)), tok,
                                 self._align(tok))
            self._cache['key_' + mode] = key
        return self._cache[mode]

    def _batch_from_slice(self, mode, pairs, batch, lo, fid2spk=None, frames=False):
        pos2same, same_tok, al = self._prepared(mode, pairs, frames)
        grouped = group_pairs(batch)
        same_ids = [pos2same[lo + i] for i, p in enumerate(batch) if p[6] == 'same']
        diff_tok = self._tokens(grouped['diff'], frames)
        return self._assemble(grouped['same'], same_tok, al, same_ids, grouped['diff'], diff_tok,
                              0, fid2spk)

    def add_tcl_to_batch(self, batch):
        """dataloader.py:314-322: append tcl / (1 - tcl) x the batch's frame pairs of
        temporal-coherence pairs (after the batch's own, un-shuffled, like the reference)."""
        X1, X2, Y = batch
        num_pairs = Y.shape[0]
        num_pairs_to_add = int((self.tcl * num_pairs) / (1 - self.tcl))
        X1_tcl, X2_tcl, Y_tcl = self.temporal_coherence_loss(num_pairs_to_add)
        if Y_tcl.shape[0] == 0:
            return batch
        n = num_pairs + Y_tcl.shape[0]
        buf = torch.empty((2 * n, X1.shape[1]), dtype=X1.dtype, device=X1.device)
        buf[:num_pairs].copy_(X1)
        buf[num_pairs:n].copy_(X1_tcl)
        buf[n:n + num_pairs].copy_(X2)
        buf[n + num_pairs:].copy_(X2_tcl)
        return buf[:n], buf[n:], torch.cat((Y, Y_tcl))

    def temporal_coherence_loss(self, num_pairs):
        """dataloader.py:324-352 in index space: per iteration one random (file, t), the frame
        pair (t, t + 1) as 'same' and (t, t + 15 / 20 / 25 / 30) as 'different'; same draws from
        ``random`` as the reference; the rows are gathered on the device."""
        rows1, rows2, Y = [], [], []
        pairs_per_iteration = len(self.TCL_DISTANCES_DIFF) + len(self.TCL_DISTANCE_SAME)
        tab = self.table
        for _ in range(round(num_pairs / pairs_per_iteration)):
            files = list(tab.files)
            if self.train_files is not None:
                files = self.train_files
            f = random.choice(files)
            key = tab._key(f)
            t = random.choice(range(tab.nrows[key] - max(self.TCL_DISTANCES_DIFF)))
            for delta, lab in [(d, 1) for d in self.TCL_DISTANCE_SAME] + \
                              [(d, -1) for d in self.TCL_DISTANCES_DIFF]:
                rows1.append(tab.row0[key] + t)
                rows2.append(tab.row0[key] + t + delta)
                Y.append(lab)
        dev = tab.feat.device
        n = len(Y)
        if n == 0:
            e = torch.empty((0, tab.dim), dtype=torch.float32, device=dev)
            return e, e.clone(), torch.empty(0, dtype=torch.float32, device=dev)
        i1 = torch.tensor(rows1, dtype=torch.int32, device=dev)
        i2 = torch.tensor(rows2, dtype=torch.int32, device=dev)
        x1, x2, _ = ops.gather_batch(tab.feat, i1, i2, None, None, n)
        return x1, x2, torch.tensor(Y, dtype=torch.float32, device=dev)

    def batch_iterator(self, train_mode=True):
        """dataloader.py:263-312: batches of `batch_size` token pairs, at most
        `num_max_minibatches` random batches per epoch.  Yields (X1, X2, y)."""
        self.load_data()
        mode = 'train' if train_mode else 'dev'
        pairs = self.pairs[mode]
        num_pairs = len(pairs)
        if self.shuffle_between_epochs:
            random.shuffle(pairs)
            self._cache.pop('key_' + mode, None)
        starts = list(range(0, num_pairs, self.batch_size))
        num_batches = len(starts)
        if self.num_max_minibatches < num_batches:
            selected = np.random.choice(range(num_batches), self.num_max_minibatches,
                                        replace=False)
        else:
            print("Number of batches not sufficient, iterating over all the batches")
            selected = np.random.permutation(range(num_batches))
        for batch_id in selected:
            lo = starts[batch_id]
            batch = self._batch_from_slice(mode, pairs, pairs[lo:lo + self.batch_size], lo)
            if self.tcl > 0:                    # dataloader.py:303-305
                batch = self.add_tcl_to_batch(batch)
            yield batch


class PairsDataLoader(OriginalDataLoader):
    """abnet3/dataloader.py:355-546: pairs given in FRAMES in one text file
    (`file1 file2 begin1 end1 begin2 end2 distance`), own train/test split,
    negatives made of two random tokens."""
    SPLIT_FILES = "files"
    SPLIT_EACH_FILE = "split_each_file"
    SPLIT_METHODS = [SPLIT_FILES, SPLIT_EACH_FILE]

    def __init__(self, pairs_path, features_path, id_to_file,
                 ratio_split_train_test=0.7,
                 batch_size=8, train_iterations=10000, test_iterations=500,
                 proportion_positive_pairs=0.5,
                 align_different_words=True,
                 split_method=SPLIT_EACH_FILE):
        self.pairs_path = pairs_path
        self.features_path = features_path
        self.features = None
        self.id_to_file = id_to_file
        self.pairs = {'train': None, 'test': None}
        self.ratio_split_train_test = ratio_split_train_test
        self.batch_size = batch_size
        self.align_different_words = align_different_words
        self.iterations = {'train': train_iterations, 'test': test_iterations}
        self.proportion_positive_pairs = proportion_positive_pairs
        self.split_method = split_method
        assert split_method in self.SPLIT_METHODS
        self.tokens = {'train': [], 'test': []}
        self.statistics_training = defaultdict(int)
        self.files = set()
        self.seed = 0
        self._cache = {}
        self._shard = None

    def __getstate__(self):
        return (self.pairs_path,
                self.features_path,
                self.id_to_file,
                self.ratio_split_train_test,
                self.align_different_words,
                self.proportion_positive_pairs)

    def __setstate__(self, state):
        (self.pairs_path, self.features_path, self.id_to_file, self.ratio_split_train_test,
         self.align_different_words, self.proportion_positive_pairs) = state
        self.features = None
        self.pairs = {'train': None, 'test': None}
        self.load_data()

    def load_data(self):
        if self.pairs['train'] is None:
            self.load_pairs()
        if self.features is None:
            print("Loading features")
            features, _, _ = read_feats(self.features_path)
            self.features = features

    def load_pairs(self):
        """dataloader.py:429-462"""
        pairs = []
        file_mapping = {}
        if self.id_to_file is not None:
            with open(self.id_to_file, 'r') as f:
                for (fid, name) in (line.strip().split() for line in f):
                    file_mapping[int(fid)] = name
        with open(self.pairs_path, 'r') as f:
            for line in f:
                file1, file2, begin1, end1, begin2, end2, _dist = line.split(' ')
                file1, file2, begin1, end1, begin2, end2 = (
                    int(file1), int(file2), int(begin1), int(end1), int(begin2), int(end2))
                file1 = file_mapping.get(file1, file1)
                file2 = file_mapping.get(file2, file2)
                self.files.add(file1)
                self.files.add(file2)
                pairs.append([file1, begin1, end1, file2, begin2, end2])
        if self.split_method == self.SPLIT_FILES:
            self.pairs['train'], self.pairs['test'] = self.split_train_test(pairs)
        elif self.split_method == self.SPLIT_EACH_FILE:
            self.pairs['train'], self.pairs['test'] = self.split_train_test_each_file(pairs)
        for mode in ('train', 'test'):
            self.pairs[mode] = self._sharded(self.pairs[mode])
            toks = set()
            for file1, begin1, end1, file2, begin2, end2 in self.pairs[mode]:
                toks.add((file1, begin1, end1))
                toks.add((file2, begin2, end2))
            self.tokens[mode] = list(toks)

    def split_train_test(self, pairs):
        """dataloader.py:464-482: hold out whole files."""
        num_files_test = int(len(self.files) * (1 - self.ratio_split_train_test))
        dev_files = set(random.sample(sorted(self.files, key=str), num_files_test))
        train_pairs, dev_pairs = [], []
        print("File selected for validation set : %s" % dev_files)
        for pair in pairs:
            file1, file2 = pair[0], pair[3]
            if file1 in dev_files and file2 in dev_files:
                dev_pairs.append(pair)
            elif file1 not in dev_files and file2 not in dev_files:
                train_pairs.append(pair)
        return train_pairs, dev_pairs

    def split_train_test_each_file(self, pairs):
        """dataloader.py:484-508: the first `ratio` of every file trains."""
        len_files = defaultdict(int)
        for file1, s1, e1, file2, s2, e2 in pairs:
            len_files[file1] = max(len_files[file1], e1)
            len_files[file2] = max(len_files[file2], e2)
        threshold = {f: n * self.ratio_split_train_test for f, n in len_files.items()}
        train_pairs, dev_pairs = [], []
        for p in pairs:
            file1, s1, e1, file2, s2, e2 = p
            if s1 > threshold[file1] and s2 > threshold[file2]:
                dev_pairs.append(p)
            elif s1 < threshold[file1] and s2 <= threshold[file2]:
                train_pairs.append(p)
        return train_pairs, dev_pairs

    def batch_iterator(self, train_mode=True):
        """dataloader.py:510-546"""
        print("constructing batches")
        mode = 'train' if train_mode else 'test'
        iterations = self.iterations[mode]
        self.load_data()
        all_positive_pairs = self.pairs[mode]
        tokens = self.tokens[mode]
        num_pairs = iterations * self.batch_size
        num_positive_pairs = int(num_pairs * self.proportion_positive_pairs)
        if num_positive_pairs > len(all_positive_pairs):
            print("Not enough positive pairs to sample this number of "
                  "iterations. There is only {}, but {} requested"
                  .format(len(all_positive_pairs), num_positive_pairs))
            num_positive_pairs = len(all_positive_pairs)
        num_negative_pairs = num_pairs - num_positive_pairs
        positive_pairs = random.sample(all_positive_pairs, num_positive_pairs)
        positive_pairs = [pair + ['same'] for pair in positive_pairs]
        tokens = random.choices(tokens, k=2 * num_negative_pairs)
        negative_pairs = [list(tokens[i]) + list(tokens[i + 1]) + ["diff"]
                          for i in range(0, len(tokens), 2)]
        pairs = positive_pairs + negative_pairs
        random.shuffle(pairs)
        print("done constructing batches for epoch")
        self._cache.pop('key_' + mode, None)
        for i in range(iterations):
            lo = i * self.batch_size
            pairs_batch = pairs[lo:lo + self.batch_size]
            if len(pairs_batch) == 0:
                break
            yield self._batch_from_slice(mode, pairs, pairs_batch, lo, frames=True)


class FramesDataLoader(OriginalDataLoader):
    """abnet3/dataloader.py:580-739: align every pair once, keep the frame
    pairs, shuffle across the whole set, cut fixed-size FRAME batches.  The
    frame-pair table (idx1, idx2 int32 global rows, y int8) stays on the GPU.

    ``exact_numpy_shuffle``: shuffle with the host numpy RNG exactly like
    ``np.random.shuffle(frames)`` (:670, :717); when False (default for tables
    above ``EXACT_SHUFFLE_LIMIT`` rows) a device permutation is used instead.

    New: ``epoch_table(train_mode)`` hands the trainer the shuffled device table and the
    batch range of the epoch instead of materialised batches, so the fused training step
    gathers its rows itself (abnet3_b200.engine.SiameseTrainStep.sweep_table);
    ``from_tokens`` builds a loader from token row ranges already on the device.
    """
    EXACT_SHUFFLE_LIMIT = 4_000_000

    def __init__(self, pairs_path, features_path,
                 batch_size=100, randomize_dataset=True, max_batches_per_epoch=None,
                 exact_numpy_shuffle=None):
        super().__init__(pairs_path, features_path)
        self.randomize_dataset = randomize_dataset
        self.batch_size = batch_size
        self.token_features = {'train': None, 'dev': None}
        self.frame_pairs = {'train': None, 'dev': None}
        self.max_batches_per_epoch = max_batches_per_epoch
        self.exact_numpy_shuffle = exact_numpy_shuffle
        self._tokens_given = None
        if self.max_batches_per_epoch is not None:
            self.batch_position = 0

    @classmethod
    def from_tokens(cls, table, tokens, batch_size=100, randomize_dataset=True,
                    max_batches_per_epoch=None, **kwargs):
        """A loader over a device-resident ``utils.FeatureTable`` and token row ranges
        instead of files: ``tokens = {'train': (same_tok, diff_tok), 'dev': (...)}`` with
        int32 [P, 4] CUDA tensors (row1, n1, row2, n2) -- what ``load_data`` derives from
        the pair files and the frame times."""
        self = cls(None, None, batch_size=batch_size, randomize_dataset=randomize_dataset,
                   max_batches_per_epoch=max_batches_per_epoch, **kwargs)
        self.features = _TableOnly(table)
        self.pairs = {'train': [], 'dev': []}
        self._shard = (0, 1)            # the caller hands every rank its own token arrays
        self._tokens_given = tokens
        return self

    def load_data(self):
        if self._tokens_given is None:
            super(FramesDataLoader, self).load_data()
        for mode in ('train', 'dev'):
            if self.frame_pairs[mode] is None:
                if mode == 'train':
                    print("Loading all frames..", end='', flush=True)
                self.token_features[mode], self.frame_pairs[mode] = \
                    self.load_all_frames(self.pairs[mode], mode)
                if mode == 'train':
                    print("Done. %s frame pairs in total." % self.frame_pairs[mode][0].numel())

    def realign(self, mode='train'):
        """Drop and recompute the frame-pair table of ``mode`` (alignment of every same pair +
        diff pairs + shuffle): what the reference does once per run (:642-671)."""
        self.frame_pairs[mode] = None
        self.token_features[mode], self.frame_pairs[mode] = \
            self.load_all_frames(self.pairs[mode], mode)
        return self.frame_pairs[mode]

    def _shuffle(self, table):
        """np.random.shuffle of the frame-pair list (:670, :717), in place (the table's device
        addresses stay valid for CUDA graphs built on them)."""
        n = table[0].numel()
        if n == 0:
            return table
        exact = self.exact_numpy_shuffle
        if exact is None:
            exact = n <= self.EXACT_SHUFFLE_LIMIT
        if exact:
            perm = np.arange(n)
            np.random.shuffle(perm)             # same draws as shuffling the list of n tuples
            perm = torch.from_numpy(perm).to(table[0].device)
        else:
            perm = torch.randperm(n, device=table[0].device)
        for t in table:
            t.copy_(t[perm])
        return table

    def _pair_labels(self, plist_same, plist_diff, mode):
        """Extra per-PAIR label columns (none here; see MultiTaskFramesDataLoader)."""
        return []

    def _frames_from_tokens(self, same_tok, diff_tok, pair_labels=()):
        """Device side of load_all_frames: same_tok / diff_tok int32 [P, 4] (host arrays or
        CUDA tensors); pair_labels: [(same int8 [Ps], diff int8 [Pd]), ...] extra label
        columns, expanded to frames.  -> (idx1, idx2, y[, extra...]) before the shuffle."""
        dev = self.table.feat.device

        def on_dev(tok):
            return tok if isinstance(tok, torch.Tensor) else torch.from_numpy(tok).to(dev)

        parts1, parts2, labels = [], [], []
        extra = [[] for _ in pair_labels]
        if len(same_tok):
            tok_d = on_dev(same_tok)
            longest = int(tok_d[:, [1, 3]].max().item())
            res = ops.align_pairs(self.table.feat, tok_d, max_frames=max(longest, 1),
                                  stack=self.table.stack)
            d1, d2, doff = ops.compact_paths(res)
            parts1.append(d1)
            parts2.append(d2)
            labels.append(torch.ones(d1.numel(), dtype=torch.int8, device=dev))
            for k, (ls, _) in enumerate(pair_labels):
                extra[k].append(torch.repeat_interleave(on_dev(ls), res.path_len.long()))
            self.statistics_training['SameType'] += int(res.valid.sum().item())
        if len(diff_tok):
            tok_d = on_dev(diff_tok)
            i1, i2, off = ops.diff_pairs(tok_d, stretch=False)
            parts1.append(i1)
            parts2.append(i2)
            labels.append(-torch.ones(i1.numel(), dtype=torch.int8, device=dev))
            for k, (_, ld) in enumerate(pair_labels):
                extra[k].append(torch.repeat_interleave(on_dev(ld), off[1:] - off[:-1]))
            self.statistics_training['DiffType'] += int(((tok_d[:, 1] > 0) &
                                                         (tok_d[:, 3] > 0)).sum().item())
        if not parts1:
            empty = torch.zeros(0, dtype=torch.int32, device=dev)
            e8 = torch.zeros(0, dtype=torch.int8, device=dev)
            return (empty, empty.clone()) + tuple(e8.clone() for _ in range(1 + len(extra)))
        y = torch.cat(labels).contiguous()
        cols = [torch.cat(e).contiguous() for e in extra]
        # label column order of the table: extra columns (y_spk) first, y_phn last
        return (torch.cat(parts1).contiguous(), torch.cat(parts2).contiguous()) + \
            tuple(cols) + (y,)

    def load_all_frames(self, pairs, mode='train'):
        """dataloader.py:617-671: -> (token table, (idx1, idx2, y)) on the device;
        same pairs in list order (each a DTW path), then diff pairs truncated to
        min(n1, n2) leading frames, then one global shuffle."""
        if self._tokens_given is not None:
            same_tok, diff_tok = self._tokens_given[mode][:2]
            plabels = self._tokens_given[mode][2] if len(self._tokens_given[mode]) > 2 else ()
        else:
            grouped = group_pairs(pairs)
            same_tok = self._tokens(grouped['same'])
            diff_tok = self._tokens(grouped['diff'])
            plabels = self._pair_labels(grouped['same'], grouped['diff'], mode)
        table = self._frames_from_tokens(same_tok, diff_tok, plabels)
        return (same_tok, diff_tok), self._shuffle(table)

    def load_batch(self, frames, token_feats=None):
        """dataloader.py:673-684.  ``frames`` = (lo, hi) slice of the frame-pair
        table (or an int64 index tensor); returns device (X1, X2, y)."""
        table = self._active_table
        idx1, idx2, y = table[0], table[1], table[-1]
        dev = idx1.device
        if isinstance(frames, tuple):
            lo, hi = frames
            n = hi - lo
            sel = None
            idx1, idx2, cols = idx1[lo:hi], idx2[lo:hi], [c[lo:hi] for c in table[2:]]
        else:
            n = frames.numel()
            sel, cols = frames, list(table[2:])
        buf = torch.empty((2 * n, self.table.dim), dtype=torch.float32, device=dev)
        outs = []
        for k, col in enumerate(cols):
            yo = torch.empty(n, dtype=torch.float32, device=dev)
            if k == 0:
                ops.gather_batch(self.table.feat, idx1, idx2, col, sel, n, out=(buf[:n], buf[n:], yo))
            else:
                yo.copy_((col if sel is None else col[sel]).float())
            outs.append(yo)
        return (buf[:n], buf[n:]) + tuple(outs)

    def _epoch_plan(self, mode):
        """The batch schedule of one epoch (dataloader.py:686-739): reshuffles as the
        reference does and returns (table, first_batch, n_batches, num_pairs)."""
        self.load_data()
        num_pairs = self.frame_pairs[mode][0].numel()
        num_batches = num_pairs // self.batch_size
        if num_batches == 0:
            num_batches = 1
        if mode == 'dev' or self.max_batches_per_epoch is None:
            first, count = 0, num_batches
            if self.randomize_dataset:
                self.frame_pairs[mode] = self._shuffle(self.frame_pairs[mode])
        else:
            if self.batch_position >= num_batches:
                print("Arrived at the end of the dataset. Starting over.")
                if self.randomize_dataset:
                    self.frame_pairs[mode] = self._shuffle(self.frame_pairs[mode])
                self.batch_position = 0
            first = self.batch_position
            count = min(self.batch_position + self.max_batches_per_epoch, num_batches) - first
            self.batch_position += self.max_batches_per_epoch
        self._active_table = self.frame_pairs[mode]
        return self._active_table, first, count, num_pairs

    def epoch_table(self, train_mode=True):
        """-> (feat, table, batch_size, first_row, n_batches): the epoch's batches are the
        consecutive ``batch_size``-row slices of ``table`` from ``first_row`` on (the single
        batch of a table smaller than ``batch_size`` is the whole table)."""
        table, first, count, num_pairs = self._epoch_plan('train' if train_mode else 'dev')
        bs = min(self.batch_size, num_pairs)
        if num_pairs == 0:
            count = 0
        return self.table.feat, table, bs, first * self.batch_size, count

    def batch_iterator(self, train_mode=True):
        """dataloader.py:686-739"""
        table, first, count, num_pairs = self._epoch_plan('train' if train_mode else 'dev')
        for i in range(first, first + count):
            lo = i * self.batch_size
            hi = min(lo + self.batch_size, num_pairs)
            yield self.load_batch((lo, hi))


class _TableOnly(object):
    """Stand-in for a Features_Accessor when only the device table exists."""

    def __init__(self, table):
        self.table = table


class MultiTaskFramesDataLoader(FramesDataLoader):
    """Frame batches with speaker labels (new): FramesDataLoader's one-time alignment and
    fixed-size frame batches (dataloader.py:580-739) with MultiTaskDataLoader's labels
    (:742-792; load_frames_from_pairs :195-202, :235-242).  Yields (X1, X2, y_spk, y_phn);
    the table is (idx1, idx2, y_spk, y_phn).  ``y_spk`` follows the reference's rule
    ``spk1 is spk2`` on the values of ``fid2spk`` (quirk q2: object identity)."""

    def __init__(self, pairs_path, features_path, fid2spk_file=None, **kwargs):
        super().__init__(pairs_path, features_path, **kwargs)
        self.fid2spk_file = fid2spk_file
        self._fid2spk = None

    def _pair_labels(self, plist_same, plist_diff, mode):
        if self._fid2spk is None:
            self._fid2spk = read_spkid_file(self.fid2spk_file)
        spk = self._fid2spk

        def col(plist, same_key, diff_key):
            out = np.empty(len(plist), dtype=np.int8)
            for k, p in enumerate(plist):
                same = spk[p[0]] is spk[p[3]]
                out[k] = 1 if same else -1
                self.statistics_training[same_key if same else diff_key] += 1
            return out

        return [(col(plist_same, 'SameTypeSameSpk', 'SameTypeDiffSpk'),
                 col(plist_diff, 'DiffTypeSameSpk', 'DiffTypeDiffSpk'))]


class MultiTaskDataLoader(OriginalDataLoader):
    """abnet3/dataloader.py:742-792: yields (X1, X2, y_spk, y_phn)."""

    def __init__(self, pairs_path, features_path, fid2spk_file=None,
                 **kwargs):
        super().__init__(pairs_path, features_path, **kwargs)
        self.fid2spk_file = fid2spk_file

    def batch_iterator(self, train_mode=True):
        self.load_data()
        mode = 'train' if train_mode else 'dev'
        pairs = self.pairs[mode]
        num_pairs = len(pairs)
        starts = list(range(0, num_pairs, self.batch_size))
        num_batches = len(starts)
        fid2spk = read_spkid_file(self.fid2spk_file)
        if self.num_max_minibatches < num_batches:
            selected = np.random.choice(range(num_batches), self.num_max_minibatches,
                                        replace=False)
        else:
            print("Number of batches not sufficient, iterating over all the batches")
            selected = np.random.permutation(range(num_batches))
        for idx in selected:
            lo = starts[idx]
            yield self._batch_from_slice(mode, pairs, pairs[lo:lo + self.batch_size], lo,
                                         fid2spk=fid2spk)
