"""Second opinion for oracle/dtw_oracle.c -- TEST INFRASTRUCTURE ONLY.

An independently written DTW in pure numpy, batched over many matrices of one shape
(every statement is vectorised over the batch; the two loops run over cells, not over
matrices).  It follows the textbook formulation instead of the C file's row sweep with
special-cased borders:

* accumulated costs on an INF-BORDERED (n1 + 1) x (n2 + 1) matrix with ``C[0, 0] = 0``
  (no first-row / first-column special case);
* traceback by ``numpy.argmin`` over the stacked predecessors ``(C[i-1, j-1], C[i-1, j],
  C[i, j-1])`` -- first minimum wins, i.e. diagonal, then up (i-1), then left (j-1): the
  convention of the ABXpy / abnet ``dtw.pyx`` lineage the reference's un-vendored
  DTW_Cython comes from (abnet3/utils.py:14, :149-151).

PARITY UNPINNED like dtw_oracle.c: both restate the published algorithm; neither can be
checked against DTW_Cython itself (absent from /root/reference, not fetchable).  What this
file adds is that the tie rule and the recurrence have been written down twice, differently,
and agree bit for bit on 10 k random and tie-heavy matrices (tests/test_oracle_golden.py).
"""
import numpy as np


def dtw_batch(D):
    """D: float64 [B, n1, n2] -> (cost [B], paths: list of (path1, path2) int arrays)."""
    D = np.asarray(D, dtype=np.float64)
    B, n1, n2 = D.shape
    C = np.full((B, n1 + 1, n2 + 1), np.inf)
    C[:, 0, 0] = 0.0
    for i in range(1, n1 + 1):
        for j in range(1, n2 + 1):
            best = np.minimum(np.minimum(C[:, i - 1, j - 1], C[:, i - 1, j]), C[:, i, j - 1])
            C[:, i, j] = D[:, i - 1, j - 1] + best
    cost = C[:, n1, n2].copy()
    # traceback, all matrices in lock step; finished ones idle at (1, 1)
    i = np.full(B, n1)
    j = np.full(B, n2)
    rows = np.arange(B)
    steps_i, steps_j = [i - 1], [j - 1]
    alive = (i > 1) | (j > 1)
    while alive.any():
        cand = np.stack([C[rows, i - 1, j - 1], C[rows, i - 1, j], C[rows, i, j - 1]])
        move = np.argmin(cand, axis=0)                 # first minimum: diag, up, left
        di = np.where(move == 2, 0, 1)
        dj = np.where(move == 1, 0, 1)
        i = np.where(alive, i - di, i)
        j = np.where(alive, j - dj, j)
        steps_i.append(np.where(alive, i - 1, -1))
        steps_j.append(np.where(alive, j - 1, -1))
        alive = (i > 1) | (j > 1)
    si, sj = np.stack(steps_i, 1), np.stack(steps_j, 1)
    paths = []
    for b in range(B):
        keep = si[b] >= 0
        paths.append((si[b][keep][::-1].astype(np.int32), sj[b][keep][::-1].astype(np.int32)))
    return cost, paths
