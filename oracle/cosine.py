"""Oracle: cosine (angular) frame distance.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/abnet3/utils.py:40-60 line for line; the only change
is ``np.arccos`` for ``scipy.arccos`` (an alias of ``numpy.arccos`` that
modern scipy removed).  With float32 inputs every operation up to and
including ``/ np.pi`` runs in float32 (numpy keeps the array dtype when the
other operand is a Python float) and only the final cast widens to float64,
which is what the reference does on its float32 features
(abnet3/utils.py:122-125 forces float32).
"""
import numpy as np


def cosine_distance(x, y):
    # abnet3/utils.py:41-42
    assert (x.dtype == np.float64 and y.dtype == np.float64) or (
        x.dtype == np.float32 and y.dtype == np.float32)
    # abnet3/utils.py:43-46
    x2 = np.sqrt(np.sum(x ** 2, axis=1))
    y2 = np.sqrt(np.sum(y ** 2, axis=1))
    ix = x2 == 0.
    iy = y2 == 0.
    # abnet3/utils.py:47
    with np.errstate(divide='ignore', invalid='ignore'):
        d = np.dot(x, y.T) / (np.outer(x2, y2))
        # abnet3/utils.py:49-53
        if d.shape == (1, 1):
            # the reference collapses the (1,1) array to a scalar and re-wraps
            d = np.array([[np.float64((np.arccos(d) / np.pi)[0, 0])]])
        else:
            d = np.float64(np.arccos(d) / np.pi)
    # abnet3/utils.py:55-58
    d[ix, :] = 1.
    d[:, iy] = 1.
    for i in np.where(ix)[0]:
        d[i, iy] = 0.
    # abnet3/utils.py:59 -- NaN (|cos| > 1 by rounding) fails this assert and
    # the callers drop the pair (abnet3/dataloader.py:188-191)
    assert np.all(d >= 0)
    return d
