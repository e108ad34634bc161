"""Host-side mirror of the alignment primitives of /root/reference/abnet3/utils.py.

Same names, argument meaning and error behaviour as the reference for the hot
path -- ``cosine_distance`` (:40-60), ``DTW`` (the external call at :149-151),
``get_dtw_alignment`` (:147-153), ``Features_Accessor`` (:118-145),
``read_dataset`` / ``group_pairs`` / ``read_pairs`` (:156-208),
``read_spkid_file`` (:23-31) -- with the arithmetic done by the sm_100a kernels
behind the C ABI (abnet3_b200/_lib.py).  There is no CPU fallback: without the
library or an sm_100 GPU these functions raise.

New, batched surface (what the dataloaders use): ``FeatureTable`` (the whole
corpus as one device-resident [n_rows, dim] table plus per-file row ranges),
``BatchAligner`` (pre-allocated, device-resident alignment of a pair list) and
``align_pairs_host`` (host buffers in, host paths out).
"""
import numpy as np
import torch

from . import ops

__all__ = ["cosine_distance", "DTW", "get_dtw_alignment", "Features_Accessor",
           "FeatureTable", "BatchAligner", "align_pairs_host", "decode_directions", "read_dataset",
           "group_pairs", "read_pairs", "read_spkid_file", "read_spk_list", "read_feats"]


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("abnet3_b200 needs an sm_100 GPU: there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


# ----------------------------------------------------------- per-pair API ---
def cosine_distance(x, y):
    """abnet3/utils.py:40-60.  float32 (or float64, computed in float32 after
    the reference's own dtype assert) [n1, D], [n2, D] -> float64 [n1, n2];
    raises AssertionError where the reference's ``assert np.all(d >= 0)`` fails."""
    assert (x.dtype == np.float64 and y.dtype == np.float64) or (
        x.dtype == np.float32 and y.dtype == np.float32)
    n1, n2 = x.shape[0], y.shape[0]
    dev = _device()
    feat = torch.from_numpy(np.ascontiguousarray(
        np.concatenate([x, y]).astype(np.float32))).to(dev)
    tok = torch.tensor([[0, n1, n1, n2]], dtype=torch.int32, device=dev)
    dist, _, valid = ops.cosine_distance(feat, tok, max_frames=max(n1, n2))
    d = dist[:n1 * n2].cpu().numpy().astype(np.float64).reshape(n1, n2)
    assert bool(valid.item()) and np.all(d >= 0)
    return d


def DTW(x, y, return_alignment=False, dist_function=None, dist_array=None):
    """The external ``dtw.DTW`` as the reference calls it (abnet3/utils.py:149-151):
    returns ``(cost, None, (path_len, path1, path2))`` when ``return_alignment``
    else ``cost``.  ``dist_array`` float64 [n1, n2] (computed with
    ``cosine_distance`` when omitted)."""
    if dist_array is None:
        dist_array = cosine_distance(x, y) if dist_function is None else dist_function(x, y)
    d = np.ascontiguousarray(dist_array, dtype=np.float64)
    n1, n2 = d.shape
    dev = _device()
    p1, p2, _, plen, cost, valid = ops.dtw_from_dist(
        torch.from_numpy(d.ravel()).to(dev),
        torch.tensor([0, n1 * n2], dtype=torch.int64, device=dev),
        torch.tensor([[n1, n2]], dtype=torch.int32, device=dev), max_frames=max(n1, n2))
    if not bool(valid.item()):
        raise ValueError("invalid distance matrix (NaN or negative entry)")
    L = int(plen.item())
    c = float(cost.item())
    if not return_alignment:
        return c
    return c, None, (L, p1[:L].cpu().numpy().astype(np.intp), p2[:L].cpu().numpy().astype(np.intp))


def get_dtw_alignment(feat1, feat2):
    """abnet3/utils.py:147-153 for ONE pair (fused kernel).  Raises
    AssertionError when the reference would (NaN distance)."""
    n1, n2 = feat1.shape[0], feat2.shape[0]
    dev = _device()
    feat = torch.from_numpy(np.ascontiguousarray(
        np.concatenate([feat1, feat2]).astype(np.float32))).to(dev)
    tok = torch.tensor([[0, n1, n1, n2]], dtype=torch.int32, device=dev)
    res = ops.align_pairs(feat, tok, max_frames=max(n1, n2))
    assert bool(res.valid.item()), "cosine distance holds a NaN (utils.py:59)"
    L = int(res.path_len.item())
    path1 = res.idx1[:L].cpu().numpy().astype(np.intp)
    path2 = (res.idx2[:L].cpu().numpy() - n1).astype(np.intp)
    assert len(path1) == len(path2)
    return path1, path2


# ------------------------------------------------------------ batched API ---
class BatchAligner(object):
    """Device-resident batched ``get_dtw_alignment`` over a pair list.

    Buffers are allocated once and reused, so ``align`` enqueues exactly one
    kernel (plus a tiny prefix sum when the pair list changes) and never
    synchronises.  The returned AlignResult aliases the internal buffers and
    is valid until the next ``align``."""

    def __init__(self, feat, max_pairs=0, max_frames=ops.MAX_TOKEN_FRAMES, stack=0):
        self.feat = feat
        self.max_frames = int(max_frames)
        self.stack = int(stack)      # 7: the table is a verified 7x40 stack (fast path)
        self._cap_pairs = 0
        self._cap_rows = 0
        self._off_key = None
        if max_pairs:
            self._reserve(max_pairs, max_pairs * (2 * self.max_frames - 1))

    def _reserve(self, n_pairs, n_rows):
        dev = self.feat.device
        if n_rows > self._cap_rows:
            self.idx1 = torch.empty(n_rows, dtype=torch.int32, device=dev)
            self.idx2 = torch.empty(n_rows, dtype=torch.int32, device=dev)
            self._cap_rows = n_rows
        if n_pairs > self._cap_pairs:
            self.path_len = torch.empty(n_pairs, dtype=torch.int32, device=dev)
            self.cost = torch.empty(n_pairs, dtype=torch.float64, device=dev)
            self.valid = torch.empty(n_pairs, dtype=torch.uint8, device=dev)
            self._ws, self._ws_bytes = ops._workspace(n_pairs, dev, self.max_frames)
            self._cap_pairs = n_pairs

    def _offsets(self, pair_tok):
        key = (pair_tok.data_ptr(), pair_tok._version, pair_tok.shape[0])
        if key != self._off_key:
            cap = (pair_tok[:, 1] + pair_tok[:, 3] - 1).clamp_min(0)
            self._off = ops._excl_cumsum(cap)
            self._off_total = int(self._off[-1].item())
            self._off_key = key
        return self._off, self._off_total

    def align(self, pair_tok):
        from . import _lib
        P = pair_tok.shape[0]
        off, total = self._offsets(pair_tok)
        self._reserve(P, max(total, 1))
        _lib.check(_lib.lib().abn_align_pairs(
            _lib.ptr(self.feat), self.feat.shape[0], self.feat.shape[1], _lib.ptr(pair_tok), P,
            self.max_frames, self.stack, _lib.ptr(off), _lib.ptr(self.idx1), _lib.ptr(self.idx2),
            _lib.ptr(self.path_len), _lib.ptr(self.cost), _lib.ptr(self.valid),
            _lib.ptr(self._ws), self._ws_bytes, _lib.stream_ptr()))
        return ops.AlignResult(self.idx1, self.idx2, off, self.path_len[:P], self.cost[:P],
                               self.valid[:P])


_pinned = {}


def _pinned_buf(name, numel, dtype):
    buf = _pinned.get(name)
    if buf is None or buf.numel() < numel or buf.dtype != dtype:
        buf = torch.empty(max(numel, 1), dtype=dtype, pin_memory=True)
        _pinned[name] = buf
    return buf[:numel]


def align_pairs_host(feat_host, pair_tok_host, max_frames=None, stack=0, last_row_of_file=None,
                     chunks=None, frames_host=None, paths="indices"):
    """Host buffers in, host results out -- the call a user of the reference's
    dataloader would make for a whole pair list.

    Input, one of:

    * ``feat_host`` [n_rows, dim] float32 -- the stacked table the reference keeps in RAM
      (abnet3/utils.py:211-217).  The whole table is uploaded; with ``last_row_of_file`` the
      device then checks whether it is a 7 x 40 stack and, if so, aligns on the stacked fast path
      (same bits).  (``stack=7`` + ``last_row_of_file``: the caller vouches for the structure and
      only the middle blocks are uploaded; kept for compatibility -- prefer ``frames_host``,
      which needs no promise.)
    * ``frames_host`` [n_rows, f] float32 -- the UN-STACKED frames (what abnet3/features.py:135-159
      stacks) with ``last_row_of_file`` ([n_rows] uint8 CPU tensor marking the last row of every
      file): uploaded as they are (7 x fewer bytes) and stacked on the device; the alignment then
      takes the stacked fast path.  Token rows in ``pair_tok_host`` index frames = table rows.

    ``pair_tok_host`` [P, 4] int32 (row1, n1, row2, n2); CPU tensors, ideally pinned.

    Output, ``paths``:

    * ``"indices"``: ``(idx1, idx2, pair_off, path_len, cost, valid)``: pair p's aligned global
      rows are ``idx1[pair_off[p]:pair_off[p+1]]`` / ``idx2[...]`` (two int32 per path step);
    * ``"directions"``: ``(dirs, dir_off, path_len, cost, valid)``: the paths as 2-bit step
      directions (abn_pack_directions; 32 x fewer bytes over PCIe); ``decode_directions``
      restores the index pairs on the host when they are wanted there."""
    dev = _device()
    if paths not in ("indices", "directions"):
        raise ValueError("paths must be 'indices' or 'directions'")
    if frames_host is not None:
        last = last_row_of_file.to(dev, non_blocking=True) if last_row_of_file is not None else None
        stack = stack or 7
        feat = ops.stack_from_frames(frames_host, stack, last)
    elif stack and last_row_of_file is not None:
        last = last_row_of_file.to(dev, non_blocking=True)
        feat = ops.stack_upload(feat_host, stack, last)
    else:
        feat = feat_host.to(dev, non_blocking=True)
        if last_row_of_file is not None and feat.shape[1] == 280 and feat.shape[0] > 1:
            # the whole table was uploaded; whether it is a 7 x 40 stack is CHECKED on the device
            # (1 ms), not promised: if so the alignment takes the stacked fast path (same bits)
            last = last_row_of_file.to(dev, non_blocking=True)
            stack = 7 if ops.stack_violations(feat, 7, last) == 0 else 0
    tok = pair_tok_host.to(dev, non_blocking=True)
    P = tok.shape[0]
    if max_frames is None:
        max_frames = int(pair_tok_host[:, [1, 3]].max().item()) if P else 1
    # The pair list is worked through in chunks so that the device -> host copy of a chunk's
    # paths (side stream) overlaps the alignment of the next chunk.  Nothing in the loop may
    # issue a small cudaMemcpy (a .item()): it would queue behind the bulk copy on the copy
    # engine and hold back the next chunk's launches -- capacities come from the host copy of
    # the list, the data-dependent total through a kernel store into pinned memory.
    n_chunks = int(chunks) if chunks else (4 if P >= 200_000 else 1)
    n_chunks = max(1, min(n_chunks, max(P, 1)))
    bounds = [P * c // n_chunks for c in range(n_chunks + 1)]
    cap_total = int((pair_tok_host[:, 1].long() + pair_tok_host[:, 3].long() - 1).clamp_min(0).sum())
    as_dirs = paths == "directions"
    if as_dirs:
        h_dirs = _pinned_buf("dirs", (cap_total + 2 * P) // 4 + n_chunks + 1, torch.uint8)
    else:
        h_idx1 = _pinned_buf("idx1", max(cap_total, 1), torch.int32)
        h_idx2 = _pinned_buf("idx2", max(cap_total, 1), torch.int32)
    h_off = _pinned_buf("off", P + 1, torch.int64)
    h_len = _pinned_buf("len", P, torch.int32)
    h_cost = _pinned_buf("cost", P, torch.float64)
    h_valid = _pinned_buf("valid", P, torch.uint8)
    h_total = _pinned_buf("total", 1, torch.int64)
    main = torch.cuda.current_stream()
    side = _side_stream(dev)
    base = 0
    keep = []
    for c in range(n_chunks):
        lo, hi = bounds[c], bounds[c + 1]
        if hi == lo:
            continue
        cap_c = int((pair_tok_host[lo:hi, 1].long() + pair_tok_host[lo:hi, 3].long() - 1).clamp_min(0).sum())
        res = ops.align_pairs(feat, tok[lo:hi], max_frames=max_frames, stack=stack, total_cap=cap_c)
        d1, d2, doff = ops.compact_paths(res, host_scalar=h_total)    # (host learns the chunk's total here)
        n = d1.numel()
        if as_dirs:
            dirs, dir_off = ops.pack_directions(d1, d2, doff, res.path_len, n)
            nb = dirs.numel()
            goff = dir_off + base if base else dir_off
        else:
            goff = doff + base if base else doff
        done = torch.cuda.Event()
        done.record(main)
        side.wait_event(done)
        with torch.cuda.stream(side):
            if as_dirs:
                h_dirs[base:base + nb].copy_(dirs, non_blocking=True)
            else:
                h_idx1[base:base + n].copy_(d1, non_blocking=True)
                h_idx2[base:base + n].copy_(d2, non_blocking=True)
            h_off[lo:hi + 1].copy_(goff, non_blocking=True)
            h_len[lo:hi].copy_(res.path_len, non_blocking=True)
            h_cost[lo:hi].copy_(res.cost, non_blocking=True)
            h_valid[lo:hi].copy_(res.valid, non_blocking=True)
        keep.append((res, d1, d2, goff, dirs if as_dirs else None))   # alive until the side stream has read them
        base += nb if as_dirs else n
    if P == 0:
        h_off.zero_()
    side.synchronize()
    main.synchronize()
    if as_dirs:
        return h_dirs[:base], h_off, h_len, h_cost, h_valid
    return h_idx1[:base], h_idx2[:base], h_off, h_len, h_cost, h_valid


def decode_directions(dirs, dir_off, path_len, pair_tok):
    """Host decoder of the ``paths="directions"`` result: -> (idx1, idx2, pair_off) as int32 /
    int64 numpy arrays, identical to what ``paths="indices"`` returns.  Every path starts at the
    pair's first rows; direction 0 advances both tokens, 1 the first, 2 the second."""
    dirs = np.asarray(dirs, dtype=np.uint8)
    dir_off = np.asarray(dir_off, dtype=np.int64)
    L = np.asarray(path_len, dtype=np.int64)
    tok = np.asarray(pair_tok, dtype=np.int64)
    P = L.shape[0]
    pair_off = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(L, out=pair_off[1:])
    total = int(pair_off[-1])
    idx1 = np.zeros(total, dtype=np.int64)
    idx2 = np.zeros(total, dtype=np.int64)
    if total == 0:
        return idx1.astype(np.int32), idx2.astype(np.int32), pair_off
    # step k >= 1 of pair p sits at bit pair (k - 1) of the pair's byte run
    pid = np.repeat(np.arange(P), L)
    k = np.arange(total) - pair_off[pid]
    km1 = np.maximum(k - 1, 0)
    byte = dirs[np.minimum(dir_off[pid] + (km1 >> 2), max(dirs.shape[0] - 1, 0))]
    d = (byte >> (2 * (km1 & 3)).astype(np.uint8)) & 3
    first = k == 0
    di = np.where(first, 0, (d != 2).astype(np.int64))
    dj = np.where(first, 0, (d != 1).astype(np.int64))
    c1, c2 = np.cumsum(di), np.cumsum(dj)
    start = pair_off[:-1][L > 0]
    base1 = np.repeat(c1[start], L[L > 0])
    base2 = np.repeat(c2[start], L[L > 0])
    idx1 = tok[pid, 0] + c1 - base1
    idx2 = tok[pid, 2] + c2 - base2
    return idx1.astype(np.int32), idx2.astype(np.int32), pair_off


_side = {}


def _side_stream(dev):
    key = str(dev)
    if key not in _side:
        _side[key] = torch.cuda.Stream(device=dev)
    return _side[key]


# ------------------------------------------------- features / pair parsing ---
class FeatureTable(object):
    """The whole corpus as ONE [n_rows, dim] float32 table (what
    ``read_feats`` loads into RAM, abnet3/utils.py:211-217), device resident,
    with per-file row ranges and frame times for token slicing."""

    def __init__(self, features, times=None, device=None):
        """``features``: {file: ndarray [n, dim]}; ``times``: {file: ndarray [n]}
        (frame centres in seconds; default 0.0025 + 0.01 k, features.py:195)."""
        self.files = list(features.keys())
        self.row0, self.nrows, self.times = {}, {}, {}
        rows = 0
        for f in self.files:
            a = features[f]
            self.row0[f], self.nrows[f] = rows, a.shape[0]
            self.times[f] = (np.asarray(times[f], dtype=np.float64) if times is not None
                             else 0.0025 + 0.01 * np.arange(a.shape[0]))
            rows += a.shape[0]
        host = np.ascontiguousarray(
            np.concatenate([np.asarray(features[f]) for f in self.files]).astype(np.float32))
        self.dim = host.shape[1]
        self.host = host
        self.device = device if device is not None else _device()
        self.feat = torch.from_numpy(host).to(self.device)
        self.stack = 0
        if self.feat.is_cuda:
            self.stack = self._detect_stack()

    @classmethod
    def from_device(cls, feat, file_off, files=None):
        """Wrap a table that is already on the GPU: ``feat`` [n_rows, dim] float32 CUDA
        tensor, ``file_off`` [n_files + 1] row offsets (any integer sequence).  No host copy
        is kept (``host`` is None); frame times default to 0.0025 + 0.01 k."""
        self = cls.__new__(cls)
        off = [int(v) for v in (file_off.tolist() if hasattr(file_off, "tolist") else file_off)]
        self.files = list(files) if files is not None else ["file%d" % i for i in range(len(off) - 1)]
        self.row0 = {f: off[i] for i, f in enumerate(self.files)}
        self.nrows = {f: off[i + 1] - off[i] for i, f in enumerate(self.files)}
        self.times = {f: 0.0025 + 0.01 * np.arange(self.nrows[f]) for f in self.files}
        self.dim = int(feat.shape[1])
        self.host = None
        self.device = feat.device
        self.feat = feat
        self.stack = self._detect_stack() if feat.is_cuda else 0
        return self

    @classmethod
    def from_frames(cls, frames, file_off, files=None, stack=7, device=None):
        """First-class UN-STACKED input: ``frames`` [n_rows, f] float32 (CPU, ideally pinned, or
        CUDA) are the f-wide frames of every file back to back; the 7-frame stack of
        abnet3/features.py:135-159 is built on the device (only the frames cross PCIe: 7 x fewer
        bytes than the stacked table) and the alignment takes the stacked fast path."""
        dev = device if device is not None else (frames.device if frames.is_cuda else _device())
        off = [int(v) for v in (file_off.tolist() if hasattr(file_off, "tolist") else file_off)]
        last = torch.zeros(frames.shape[0], dtype=torch.uint8)
        ends = [o - 1 for o in off[1:] if o > 0]
        last[torch.tensor(ends, dtype=torch.int64)] = 1
        feat = ops.stack_from_frames(frames, stack, last.to(dev, non_blocking=True))
        self = cls.from_device(feat, off, files)
        return self

    @classmethod
    def from_host(cls, feat_host, file_off, files=None, device=None):
        """Upload a host table ([n_rows, dim] float32 CPU tensor, ideally pinned) in ONE
        asynchronous copy and wrap it (see from_device)."""
        dev = device if device is not None else _device()
        return cls.from_device(feat_host.to(dev, non_blocking=True), file_off, files)

    def _detect_stack(self, stack=7):
        """7 when every file is a 7-frame stack of dim/7-wide frames
        (abnet3/features.py:135-159), checked row by row on the device; the
        alignment kernels then take their stacked fast path (same results)."""
        if self.dim != 280 or self.feat.shape[0] < 2:
            return 0
        last = torch.zeros(self.feat.shape[0], dtype=torch.uint8, device=self.device)
        ends = [self.row0[f] + self.nrows[f] - 1 for f in self.files if self.nrows[f] > 0]
        last[torch.tensor(ends, dtype=torch.int64, device=self.device)] = 1
        return stack if ops.stack_violations(self.feat, stack, last) == 0 else 0

    def _key(self, f):
        if f in self.row0:
            return f
        alt = f.encode("UTF-8") if isinstance(f, str) else f.decode("UTF-8")
        return alt                     # utils.py:135-137: bytes or str file keys

    def token_by_time(self, f, on, off):
        """(row_start, n_frames) of the frames with on <= t <= off
        (abnet3/utils.py:128-131, inclusive on both ends)."""
        f = self._key(f)
        t = self.times[f]
        lo = int(np.searchsorted(t, on, side="left"))
        hi = int(np.searchsorted(t, off, side="right"))
        return self.row0[f] + lo, max(hi - lo, 0)

    def tokens_by_time(self, files, on, off):
        """Vectorised ``token_by_time`` over arrays: -> int32 [n, 2] (row_start, n)."""
        files = np.asarray([self._key(f) for f in files], dtype=object)
        on, off = np.asarray(on, dtype=np.float64), np.asarray(off, dtype=np.float64)
        out = np.zeros((len(files), 2), dtype=np.int32)
        for f in set(files.tolist()):
            m = np.nonzero(files == f)[0]
            t = self.times[f]
            lo = np.searchsorted(t, on[m], side="left")
            hi = np.searchsorted(t, off[m], side="right")
            out[m, 0] = self.row0[f] + lo
            out[m, 1] = np.maximum(hi - lo, 0)
        return out

    def token_by_frames(self, f, frame_on, frame_off):
        """abnet3/utils.py:141-145: ``features[f][frame_on:frame_off]`` with
        Python slice clamping."""
        f = self._key(f)
        n = self.nrows[f]
        lo, hi, _ = slice(frame_on, frame_off).indices(n)
        return self.row0[f] + lo, max(hi - lo, 0)


class Features_Accessor(object):
    """abnet3/utils.py:118-145 on top of a FeatureTable: ``get`` /
    ``get_between_frames`` return host float32 rows like the reference."""

    def __init__(self, times, features):
        first = features[list(features.keys())[0]]
        if first.dtype != np.float32:            # utils.py:122-125, :228-235
            features = {k: v.astype(np.float32) for k, v in features.items()}
            print('Casted features to correct type np.float32')
        self.times = times
        self.features = features
        self.table = FeatureTable(features, times)

    def get(self, f, on, off):
        s, n = self.table.token_by_time(f, on, off)
        return self.table.host[s:s + n]

    def get_between_frames(self, f, frame_on, frame_off):
        s, n = self.table.token_by_frames(f, frame_on, frame_off)
        return self.table.host[s:s + n]


def read_feats(features_file, align_features_file=None):
    """abnet3/utils.py:211-226: load the WHOLE feature file and return
    ``(features_accessor, align_features, feat_dim)``.

    The reference reads the h5features container with the `h5features` package,
    which is not installed in this image; on-disk formats are a
    "next" row of SURVEY.md 8f.  Accepted here: a ``.npz`` archive with one
    ``<file>`` array [n, dim] per file and optional ``times/<file>`` arrays, an
    in-memory ``{file: array}`` dict, or -- when `h5features` is importable --
    the reference's own format."""
    if isinstance(features_file, Features_Accessor):
        return features_file, None, features_file.table.dim
    if isinstance(features_file, dict):
        feats, times = features_file, None
    elif str(features_file).endswith(".npz"):
        z = np.load(features_file)
        feats = {k: z[k] for k in z.files if not k.startswith("times/")}
        times = {k: z["times/" + k] for k in feats} if any(
            k.startswith("times/") for k in z.files) else None
    else:
        try:
            import h5features
        except ImportError:
            raise ImportError("reading %r needs the `h5features` package (not installed); "
                              "pass a .npz archive or a {file: array} dict" % (features_file,))
        with h5features.Reader(features_file, 'features') as fh:
            data = fh.read()
        times, feats = data.dict_labels(), data.dict_features()
    if times is None:
        times = {k: 0.0025 + 0.01 * np.arange(v.shape[0]) for k, v in feats.items()}
    acc = Features_Accessor(times, feats)
    return acc, None, acc.table.dim


def read_spkid_file(spkid_file):
    """abnet3/utils.py:23-31"""
    spk = {}
    with open(spkid_file, 'r') as fh:
        for line in fh.readlines():
            fid, spkid = line.strip().split(" ")
            assert not (fid in spk)
            spk[fid] = spkid
    return spk


def read_spk_list(spk_file):
    """abnet3/utils.py:34-37"""
    with open(spk_file, 'r') as fh:
        return [line.strip() for line in fh.readlines()]


def read_dataset(dataset_file):
    """abnet3/utils.py:156-173: [(f1, s1, e1, f2, s2, e2, pair_type), ...]"""
    pairs = []
    with open(dataset_file, 'r') as fh:
        for line in fh.readlines():
            tokens = line.strip().split(" ")
            assert len(tokens) == 7
            f1, s1, e1, f2, s2, e2, pair_type = tokens
            s1, e1, s2, e2 = float(s1), float(e1), float(s2), float(e2)
            assert pair_type in ['same', 'diff'], \
                'Unsupported pair type {0}'.format(pair_type)
            pairs.append((f1, s1, e1, f2, s2, e2, pair_type))
    return pairs


def group_pairs(pairs):
    """abnet3/utils.py:176-193"""
    grouped_pairs = {'same': [], 'diff': []}
    for f1, s1, e1, f2, s2, e2, pair_type in pairs:
        assert pair_type in grouped_pairs, \
            'Unsupported pair type {0}'.format(pair_type)
        grouped_pairs[pair_type].append((f1, s1, e1, f2, s2, e2))
    return grouped_pairs


def read_pairs(pair_file):
    """abnet3/utils.py:196-208"""
    return group_pairs(read_dataset(pair_file))
