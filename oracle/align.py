"""Oracle: pair alignment and batch generation.  TEST INFRASTRUCTURE ONLY.

Restates, on plain numpy arrays and Python loops (small cases only):

* ``Features_Accessor``          /root/reference/abnet3/utils.py:118-145
* ``get_dtw_alignment``          abnet3/utils.py:147-153
* ``load_frames_from_pairs``     abnet3/dataloader.py:166-261
* ``load_all_frames``            abnet3/dataloader.py:617-671

DTW itself is oracle/dtw_oracle.c (PARITY UNPINNED, see its header).
"""
from collections import defaultdict

import numpy as np

from .cosine import cosine_distance
from .dtw import dtw


class FeaturesAccessor(object):
    """abnet3/utils.py:118-145 (features already float32)."""

    def __init__(self, times, features):
        self.times = times
        self.features = features

    @staticmethod
    def get_features_between(feature, time, start, end):
        # abnet3/utils.py:128-131: inclusive on both ends
        t = np.where(np.logical_and(time >= start, time <= end))[0]
        return feature[t, :]

    def get(self, f, on, off):
        return self.get_features_between(self.features[f], self.times[f],
                                         on, off)

    def get_between_frames(self, f, frame_on, frame_off):
        # abnet3/utils.py:141-145
        return self.features[f][frame_on:frame_off]


def get_dtw_alignment(feat1, feat2):
    """abnet3/utils.py:147-153.  Raises (AssertionError / ValueError) exactly
    where the reference raises, so callers can drop the pair."""
    distance_array = cosine_distance(feat1, feat2)
    _, path1, path2 = dtw(distance_array)
    assert len(path1) == len(path2)
    return path1, path2


def align_pairs(feat, pair_tok):
    """Batched form used by the parity tests.

    ``feat`` [T, D] float32 table, ``pair_tok`` [P, 4] int (start1, n1, start2,
    n2).  Returns a list of dicts with local paths, cost, validity and the
    distance matrix -- what ``get_dtw_alignment`` yields pair by pair
    (abnet3/dataloader.py:183-191: any exception drops the pair).
    """
    out = []
    for s1, n1, s2, n2 in np.asarray(pair_tok).tolist():
        f1, f2 = feat[s1:s1 + n1], feat[s2:s2 + n2]
        rec = {"valid": False, "path1": None, "path2": None, "cost": np.nan,
               "dist": None, "ties": 0}
        if n1 > 0 and n2 > 0:
            try:
                d = cosine_distance(f1, f2)
                cost, p1, p2, ties = dtw(d, return_ties=True)
                rec.update(valid=True, path1=p1, path2=p2, cost=cost, dist=d,
                           ties=ties)
            except Exception:          # abnet3/dataloader.py:190-191
                pass
        out.append(rec)
    return out


def diff_pair_indices(n1, n2, align_different_words=False):
    """Row selection for a 'diff' pair, abnet3/dataloader.py:208-231.

    Returns ``(swap, i1, i2, n_labels)``: X1 rows are ``tokA[i1]`` and X2 rows
    ``tokB[i2]`` where (tokA, tokB) = (feat1, feat2) unless ``swap`` says
    which operand each side reads (see below); ``n_labels = min(n1, n2)`` is
    how many -1 labels the reference appends (:231) -- with
    ``align_different_words`` that is FEWER than the rows it appends (quirk q3).

    ``swap`` is a pair ``(srcA, srcB)`` of 1/2 telling which token feeds
    X1 / X2: the reference puts ``max_word`` in X1 and the stretched
    ``min_word`` in X2 (:216-225), and Python's ``min``/``max`` both return
    the FIRST operand on a length tie, so for n1 == n2 both sides read feat1.
    """
    if align_different_words:
        # min((feat1, feat2), key=len) / max(...): first wins on ties
        src_min = 1 if n1 <= n2 else 2
        src_max = 1 if n1 >= n2 else 2
        len_min, len_max = min(n1, n2), max(n1, n2)
        mapping = np.linspace(0, len_min - 1, num=len_max)
        mapping = np.rint(mapping).astype(int)
        return (src_max, src_min), np.arange(len_max), mapping, min(n1, n2)
    m = min(n1, n2)
    return (1, 2), np.arange(m), np.arange(m), m


def load_frames_from_pairs(features, pairs, seed=0, fid2spk=None,
                           frames=False, align_different_words=False,
                           statistics=None):
    """abnet3/dataloader.py:166-261, verbatim control flow.

    ``features`` is a FeaturesAccessor, ``pairs`` = {'same': [...], 'diff':
    [...]} of (f1, s1, e1, f2, s2, e2).  Returns (X1, X2, y_phn) or
    (X1, X2, y_spk, y_phn).
    """
    stats = statistics if statistics is not None else defaultdict(int)
    get_features = features.get_between_frames if frames else features.get
    token_feats = {}
    for kind in ('same', 'diff'):                       # :147-164
        for f1, s1, e1, f2, s2, e2 in pairs[kind]:
            if (f1, s1, e1) not in token_feats:
                token_feats[f1, s1, e1] = get_features(f1, s1, e1)
            if (f2, s2, e2) not in token_feats:
                token_feats[f2, s2, e2] = get_features(f2, s2, e2)

    X1, X2, y_phn, y_spk = [], [], [], []
    for f1, s1, e1, f2, s2, e2 in pairs['same']:         # :183-206
        if (s1 > e1) or (s2 > e2):
            continue
        feat1 = token_feats[f1, s1, e1]
        feat2 = token_feats[f2, s2, e2]
        try:
            path1, path2 = get_dtw_alignment(feat1, feat2)
        except Exception:
            continue
        stats['SameType'] += 1
        if fid2spk:
            spk1, spk2 = fid2spk[f1], fid2spk[f2]
            if spk1 is spk2:                             # quirk q2: identity
                y_spk.append(np.ones(len(path1)))
                stats['SameTypeSameSpk'] += 1
            else:
                y_spk.append(-1 * np.ones(len(path1)))
                stats['SameTypeDiffSpk'] += 1
        X1.append(feat1[path1, :])
        X2.append(feat2[path2, :])
        y_phn.append(np.ones(len(path1)))

    for f1, s1, e1, f2, s2, e2 in pairs['diff']:         # :208-242
        if (s1 > e1) or (s2 > e2):
            continue
        feat1 = token_feats[f1, s1, e1]
        feat2 = token_feats[f2, s2, e2]
        n1, n2 = feat1.shape[0], feat2.shape[0]
        if align_different_words:
            min_word = min((feat1, feat2), key=len)
            max_word = max((feat1, feat2), key=len)
            mapping = np.linspace(0, len(min_word) - 1, num=len(max_word))
            mapping = np.rint(mapping).astype(int)
            word1 = max_word
            word2 = min_word[mapping, :]
        else:
            word1 = feat1[:min(n1, n2), :]
            word2 = feat2[:min(n1, n2), :]
        X1.append(word1)
        X2.append(word2)
        y_phn.append(-1 * np.ones(min(n1, n2)))
        stats['DiffType'] += 1
        if fid2spk:
            spk1, spk2 = fid2spk[f1], fid2spk[f2]
            if spk1 is spk2:
                y_spk.append(np.ones(min(n1, n2)))
                stats['DiffTypeSameSpk'] += 1
            else:
                y_spk.append(-1 * np.ones(min(n1, n2)))
                stats['DiffTypeDiffSpk'] += 1

    X1, X2, y_phn = np.vstack(X1), np.vstack(X2), np.concatenate(y_phn)  # :247
    np.random.seed(seed)                                 # :248 (quirk q1)
    n_pairs = len(y_phn)
    ind = np.random.permutation(n_pairs)
    X1 = X1[ind, :]
    X2 = X2[ind, :]
    y_phn = y_phn[ind]
    if fid2spk:
        y_spk = np.concatenate(y_spk)[ind]
        return X1, X2, y_spk, y_phn
    return X1, X2, y_phn


def load_all_frames(features, pairs, shuffle=False):
    """abnet3/dataloader.py:617-671 up to (and optionally including) the final
    ``np.random.shuffle``.  ``pairs`` is the grouped dict.  Returns
    ``(token_feats, frames)`` with frames = [(f1,s1,e1,i1,f2,s2,e2,i2,±1)]."""
    token_feats = {}
    for kind in ('same', 'diff'):
        for f1, s1, e1, f2, s2, e2 in pairs[kind]:
            for key in ((f1, s1, e1), (f2, s2, e2)):
                if key not in token_feats:
                    token_feats[key] = features.get(*key)
    frames = []
    for f1, s1, e1, f2, s2, e2 in pairs['same']:
        if (s1 > e1) or (s2 > e2):
            continue
        feat1, feat2 = token_feats[f1, s1, e1], token_feats[f2, s2, e2]
        try:
            path1, path2 = get_dtw_alignment(feat1, feat2)
        except Exception:
            continue
        for i1, i2 in zip(path1, path2):
            frames.append((f1, s1, e1, int(i1), f2, s2, e2, int(i2), 1))
    for f1, s1, e1, f2, s2, e2 in pairs['diff']:
        if (s1 > e1) or (s2 > e2):
            continue
        n1 = token_feats[f1, s1, e1].shape[0]
        n2 = token_feats[f2, s2, e2].shape[0]
        for i in range(min(n1, n2)):
            frames.append((f1, s1, e1, i, f2, s2, e2, i, -1))
    if shuffle:
        np.random.shuffle(frames)
    return token_feats, frames


TCL_DISTANCE_SAME = [1]                                  # abnet3/dataloader.py:51-52
TCL_DISTANCES_DIFF = [15, 20, 25, 30]


def temporal_coherence_loss(features, num_pairs, train_files=None):
    """abnet3/dataloader.py:324-352, verbatim control flow (the ``random`` module's state
    decides the draws): -> (X1, X2, Y)."""
    import random
    X1, X2, Y = [], [], []
    pairs_per_iteration = len(TCL_DISTANCES_DIFF) + len(TCL_DISTANCE_SAME)
    for _ in range(round(num_pairs / pairs_per_iteration)):
        files = list(features.features.keys())
        if train_files is not None:
            files = train_files
        f = random.choice(files)
        file_features = features.features[f]
        t = random.choice(range(len(file_features) - max(TCL_DISTANCES_DIFF)))
        for delta in TCL_DISTANCE_SAME:
            X1.append(file_features[t])
            X2.append(file_features[t + delta])
            Y.append(1)
        for delta in TCL_DISTANCES_DIFF:
            X1.append(file_features[t])
            X2.append(file_features[t + delta])
            Y.append(-1)
    return np.vstack(X1), np.vstack(X2), np.array(Y)


def add_tcl_to_batch(features, batch, tcl, train_files=None):
    """abnet3/dataloader.py:314-322."""
    X1, X2, Y = batch
    num_pairs = len(Y)
    num_pairs_to_add = int((tcl * num_pairs) / (1 - tcl))
    X1_tcl, X2_tcl, Y_tcl = temporal_coherence_loss(features, num_pairs_to_add, train_files)
    return np.vstack((X1, X1_tcl)), np.vstack((X2, X2_tcl)), np.concatenate((Y, Y_tcl))
