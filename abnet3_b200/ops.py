"""Typed wrappers over the C ABI: torch tensors in, torch tensors out.

torch is used here for device memory, streams and prefix sums only; all the
arithmetic of the hot path happens inside libabnet3_b200.so.  Every function
requires CUDA tensors and raises if the library or an sm_100 device is missing.
"""
import ctypes as _c
import os
from collections import namedtuple

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

MAX_TOKEN_FRAMES = 512           # ABN_MAX_TOKEN_FRAMES (abn_align_pairs / abn_cosine_distance)
MAX_DTW_FROM_DIST = 96           # abn_dtw_from_dist test hook (single tile)
ACT = {None: 0, "none": 0, "sigmoid": 1, "tanh": 2, "relu": 3}
LOSS_KIND = {"coscos2": 0, "cosmargin": 1}
OPT_KIND = {"sgd": 0, "adadelta": 1, "adam": 2}

AlignResult = namedtuple(
    "AlignResult", "idx1 idx2 path_off path_len cost valid")


def _req(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError("%s must be a CUDA tensor (no CPU path exists)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def _excl_cumsum(counts):
    """[n] -> int64 [n+1] exclusive prefix sums (device)."""
    off = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=counts.device)
    torch.cumsum(counts.to(torch.int64), 0, out=off[1:])
    return off


ALIGN_WS_BUDGET = int(float(os.environ.get("ABN_ALIGN_WS_GB", "16")) * (1 << 30))


def _workspace(n_pairs, device, max_frames=1, budget=None):
    """Workspace of abn_align_pairs: the class-sorted pair order plus the hand-over
    slots of the distance matrices.  One slot per pair when that fits `budget` bytes
    (default ABN_ALIGN_WS_GB = 16), else the fewest rounds that do."""
    fn = _lib.lib().abn_align_workspace_bytes
    budget = ALIGN_WS_BUDGET if budget is None else int(budget)
    rounds = 1
    nbytes = fn(n_pairs, max_frames, rounds)
    while nbytes > budget and rounds < max(n_pairs, 1):
        rounds = min(max(n_pairs, 1), max(rounds + 1, -(-nbytes * rounds // budget)))
        nbytes = fn(n_pairs, max_frames, rounds)
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


def read_scalar(src, host_scalar):
    """int(src[0]) for a 1-element int64 device tensor WITHOUT a cudaMemcpy: a kernel stores it into
    ``host_scalar`` (pinned 1-element int64 tensor) and the host waits on an event.  A small
    device -> host memcpy would queue behind a bulk copy that is using the copy engine."""
    check(_lib.lib().abn_store_scalar64(ptr(src), host_scalar.data_ptr(), stream_ptr()))
    ev = torch.cuda.Event()
    ev.record()
    ev.synchronize()
    return int(host_scalar[0])


def align_pairs(feat, pair_tok, max_frames=None, stack=0, total_cap=None):
    """Fused cosine distance + DTW + traceback for every pair
    (abnet3/utils.py:147-153 per pair).  Returns an AlignResult whose idx1/idx2
    hold GLOBAL feature rows in per-pair slots of capacity n1+n2-1 starting at
    path_off[p]; valid[p] == 0 means the reference would have dropped the pair.
    """
    _req(feat, torch.float32, "feat")
    _req(pair_tok, torch.int32, "pair_tok")
    P = pair_tok.shape[0]
    dev = feat.device
    if max_frames is None:
        max_frames = int(pair_tok[:, [1, 3]].max().item()) if P else 1
    cap = (pair_tok[:, 1] + pair_tok[:, 3] - 1).clamp_min(0)
    path_off = _excl_cumsum(cap)
    # total_cap: sum(n1 + n2 - 1) when the caller already knows it from a host copy of the list
    total = int(total_cap) if total_cap is not None else (int(path_off[-1].item()) if P else 0)
    idx1 = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    idx2 = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    path_len = torch.zeros(P, dtype=torch.int32, device=dev)
    cost = torch.full((P,), float("nan"), dtype=torch.float64, device=dev)
    valid = torch.zeros(P, dtype=torch.uint8, device=dev)
    ws, ws_bytes = _workspace(P, dev, max(max_frames, 1))
    check(_lib.lib().abn_align_pairs(
        ptr(feat), feat.shape[0], feat.shape[1], ptr(pair_tok), P, max(max_frames, 1), int(stack),
        ptr(path_off), ptr(idx1), ptr(idx2), ptr(path_len), ptr(cost), ptr(valid),
        ptr(ws), ws_bytes, stream_ptr()))
    return AlignResult(idx1, idx2, path_off, path_len, cost, valid)


def cosine_distance(feat, pair_tok, max_frames=None, stack=0):
    """Batched abnet3/utils.py:40-60.  Returns (dist float32 flat, dist_off, valid)."""
    _req(feat, torch.float32, "feat")
    _req(pair_tok, torch.int32, "pair_tok")
    P = pair_tok.shape[0]
    dev = feat.device
    if max_frames is None:
        max_frames = int(pair_tok[:, [1, 3]].max().item()) if P else 1
    cells = (pair_tok[:, 1].to(torch.int64) * pair_tok[:, 3].to(torch.int64)).clamp_min(0)
    dist_off = _excl_cumsum(cells)
    total = int(dist_off[-1].item()) if P else 0
    dist = torch.full((max(total, 1),), float("nan"), dtype=torch.float32, device=dev)
    valid = torch.zeros(P, dtype=torch.uint8, device=dev)
    ws, ws_bytes = _workspace(P, dev)
    check(_lib.lib().abn_cosine_distance(
        ptr(feat), feat.shape[0], feat.shape[1], ptr(pair_tok), P, max(max_frames, 1), int(stack),
        ptr(dist_off), ptr(dist), ptr(valid), ptr(ws), ws_bytes, stream_ptr()))
    return dist, dist_off, valid


def stack_violations(feat, stack=7, last_row_of_file=None):
    """Number of rows whose overlap with the next row breaks the S-frame stack
    structure (0 = the table qualifies for the stacked fast path)."""
    _req(feat, torch.float32, "feat")
    if last_row_of_file is not None:
        _req(last_row_of_file, torch.uint8, "last_row_of_file")
    count = torch.zeros(1, dtype=torch.int64, device=feat.device)
    check(_lib.lib().abn_stack_violations(ptr(feat), feat.shape[0], feat.shape[1], int(stack),
                                          ptr(last_row_of_file), ptr(count), stream_ptr()))
    return int(count.item())


def stack_upload(feat_host, stack=7, last_row_of_file=None, out=None):
    """Device copy of a HOST table that is a verified S-frame stack: uploads the middle
    blocks only (S x fewer PCIe bytes) and rebuilds the rest on the device."""
    if feat_host.is_cuda or feat_host.dtype != torch.float32 or not feat_host.is_contiguous():
        raise TypeError("feat_host must be a contiguous float32 CPU tensor")
    dev = torch.device("cuda", torch.cuda.current_device())
    if last_row_of_file is not None:
        _req(last_row_of_file, torch.uint8, "last_row_of_file")
    feat = out if out is not None else torch.empty(feat_host.shape, dtype=torch.float32, device=dev)
    check(_lib.lib().abn_stack_upload(ptr(feat), feat_host.data_ptr(), feat_host.shape[0],
                                      feat_host.shape[1], int(stack), ptr(last_row_of_file),
                                      stream_ptr()))
    return feat


def stack_from_frames(frames, stack=7, last_row_of_file=None, out=None):
    """Stacked device table [n_rows, stack * f] from the UN-STACKED frames [n_rows, f] (a CPU
    tensor, ideally pinned, or a CUDA tensor): abnet3/features.py:135-159 on the device.  The
    frames are all that crosses PCIe."""
    if frames.dtype != torch.float32 or not frames.is_contiguous() or frames.dim() != 2:
        raise TypeError("frames must be a contiguous float32 [n_rows, f] tensor")
    dev = frames.device if frames.is_cuda else torch.device("cuda", torch.cuda.current_device())
    if last_row_of_file is not None:
        _req(last_row_of_file, torch.uint8, "last_row_of_file")
    n, f = frames.shape
    frames_d = frames if frames.is_cuda else frames.to(dev, non_blocking=True)   # ONE contiguous upload
    feat = out if out is not None else torch.empty((n, stack * f), dtype=torch.float32, device=dev)
    check(_lib.lib().abn_stack_from_frames(ptr(feat), ptr(frames_d), n, f, int(stack),
                                           ptr(last_row_of_file), stream_ptr()))
    return feat


def pack_directions(d1, d2, dst_off, path_len, n_total):
    """The dense paths (compact_paths: ``n_total`` index pairs) as 2-bit step directions, four
    per byte (abn_pack_directions).  Returns (dirs uint8 [(n_total + 2 P) // 4 + 1], dir_off int64
    [P + 1]); pair p's path_len[p] - 1 directions start at byte dir_off[p]."""
    _req(d1, torch.int32, "idx1")
    _req(d2, torch.int32, "idx2")
    _req(path_len, torch.int32, "path_len")
    P = path_len.numel()
    nbytes = ((path_len.to(torch.int64) - 1).clamp_min(0) + 3) // 4
    dir_off = _excl_cumsum(nbytes)
    dirs = torch.empty((int(n_total) + 2 * P) // 4 + 1, dtype=torch.uint8, device=d1.device)
    check(_lib.lib().abn_pack_directions(ptr(d1), ptr(d2), ptr(dst_off), ptr(path_len), P,
                                         ptr(dir_off), ptr(dirs), stream_ptr()))
    return dirs, dir_off


def dtw_from_dist(dist, dist_off, shape, max_frames=None):
    """Batched DTW on given float64 matrices (the call at abnet3/utils.py:149-151).
    Returns (path1, path2, path_off, path_len, cost, valid) with LOCAL indices."""
    _req(dist, torch.float64, "dist")
    _req(dist_off, torch.int64, "dist_off")
    _req(shape, torch.int32, "shape")
    P = shape.shape[0]
    dev = dist.device
    if max_frames is None:
        max_frames = int(shape.max().item()) if P else 1
    cap = (shape[:, 0] + shape[:, 1] - 1).clamp_min(0)
    path_off = _excl_cumsum(cap)
    total = int(path_off[-1].item()) if P else 0
    p1 = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    p2 = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    path_len = torch.zeros(P, dtype=torch.int32, device=dev)
    cost = torch.full((P,), float("nan"), dtype=torch.float64, device=dev)
    valid = torch.zeros(P, dtype=torch.uint8, device=dev)
    check(_lib.lib().abn_dtw_from_dist(
        ptr(dist), ptr(dist_off), ptr(shape), P, max(max_frames, 1), ptr(path_off),
        ptr(p1), ptr(p2), ptr(path_len), ptr(cost), ptr(valid), stream_ptr()))
    return p1, p2, path_off, path_len, cost, valid


def diff_pairs(pair_tok, stretch=False):
    """Frame-index pairs of 'diff' pairs (abnet3/dataloader.py:208-231).
    Returns (idx1, idx2, out_off): rows of pair p at out_off[p]..out_off[p+1]."""
    _req(pair_tok, torch.int32, "pair_tok")
    P = pair_tok.shape[0]
    n1, n2 = pair_tok[:, 1], pair_tok[:, 3]
    rows = torch.maximum(n1, n2) if stretch else torch.minimum(n1, n2)
    rows = torch.where((n1 > 0) & (n2 > 0), rows, torch.zeros_like(rows))
    out_off = _excl_cumsum(rows)
    total = int(out_off[-1].item()) if P else 0
    idx1 = torch.empty(max(total, 1), dtype=torch.int32, device=pair_tok.device)
    idx2 = torch.empty(max(total, 1), dtype=torch.int32, device=pair_tok.device)
    check(_lib.lib().abn_diff_pairs(ptr(pair_tok), P, int(bool(stretch)), ptr(out_off),
                                    ptr(idx1), ptr(idx2), stream_ptr()))
    return idx1[:total], idx2[:total], out_off


def compact_paths(res, host_scalar=None):
    """Dense (idx1, idx2, dst_off) table from an AlignResult
    (what FramesDataLoader.load_all_frames accumulates, dataloader.py:642-653).
    host_scalar: pinned int64 [1] tensor through which the total is read (see read_scalar)."""
    P = res.path_len.numel()
    dst_off = _excl_cumsum(res.path_len)
    if P == 0:
        total = 0
    elif host_scalar is not None:
        total = read_scalar(dst_off[-1:], host_scalar)
    else:
        total = int(dst_off[-1].item())
    d1 = torch.empty(max(total, 1), dtype=torch.int32, device=res.idx1.device)
    d2 = torch.empty(max(total, 1), dtype=torch.int32, device=res.idx1.device)
    check(_lib.lib().abn_compact_paths(ptr(res.idx1), ptr(res.idx2), ptr(res.path_off),
                                       ptr(dst_off), ptr(res.path_len), P, ptr(d1), ptr(d2),
                                       stream_ptr()))
    return d1[:total], d2[:total], dst_off


def gather_batch(feat, idx1, idx2, y=None, sel=None, n=None, out=None):
    """X1 = feat[idx1[sel]], X2 = feat[idx2[sel]], y_out = float(y[sel])."""
    _req(feat, torch.float32, "feat")
    _req(idx1, torch.int32, "idx1")
    _req(idx2, torch.int32, "idx2")
    if sel is not None:
        _req(sel, torch.int64, "sel")
        n = sel.numel() if n is None else n
    elif n is None:
        n = idx1.numel()
    if y is not None:
        _req(y, torch.int8, "y")
    dim = feat.shape[1]
    if out is None:
        x1 = torch.empty((n, dim), dtype=torch.float32, device=feat.device)
        x2 = torch.empty((n, dim), dtype=torch.float32, device=feat.device)
        yo = torch.empty(n, dtype=torch.float32, device=feat.device)
    else:
        x1, x2, yo = out
    check(_lib.lib().abn_gather_batch(ptr(feat), dim, ptr(idx1), ptr(idx2), ptr(y), ptr(sel),
                                      n, ptr(x1), ptr(x2), ptr(yo), stream_ptr()))
    return x1, x2, yo


def pair_loss(e1, e2, y, kind="coscos2", margin=0.5, scale=1.0, loss_out=None,
              need_grad=True, grads=None):
    """Fused loss value + gradients (abnet3/loss.py:46-67, :85-105).
    Returns (loss[1], de1, de2); ``loss_out`` is accumulated into when given."""
    for t, nm in ((e1, "e1"), (e2, "e2")):
        if not (t.is_cuda and t.dtype == torch.float32 and t.stride(-1) == 1):
            raise TypeError("%s must be a CUDA float32 tensor with a contiguous last dim" % nm)
    _req(y, torch.float32, "y")
    n, dim = e1.shape
    ld = e1.stride(0)
    if e2.stride(0) != ld:
        raise ValueError("e1 and e2 must share their row stride")
    loss = loss_out if loss_out is not None else torch.zeros(1, dtype=torch.float32,
                                                            device=e1.device)
    if grads is not None:
        de1, de2 = grads
    else:
        de1 = torch.empty_like(e1) if need_grad else None
        de2 = torch.empty_like(e2) if need_grad else None
    if de1 is not None and (de1.stride(0) != ld or de2.stride(0) != ld):
        raise ValueError("gradient buffers must share the embeddings' row stride")
    check(_lib.lib().abn_pair_loss(ptr(e1), ptr(e2), ptr(y), n, dim, ld, LOSS_KIND[kind],
                                   float(margin), float(scale), ptr(loss), ptr(de1), ptr(de2),
                                   stream_ptr()))
    return loss, de1, de2


# ------------------------------------------------------------------ dropout ---
def dropout_state(device, seed=None):
    """Device {seed, step} of the dropout masks (abn_dropout): an int64 [2] tensor.  ``seed``
    None: drawn from torch's CPU generator (so ``torch.manual_seed`` controls it).  Advance the
    step with ``state[1] += 1`` (stream ordered; CUDA-graph capturable)."""
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return torch.tensor([seed, 0], dtype=torch.int64, device=device)


def dropout_spec(state, p, layer):
    """ctypes abn_dropout for one layer (state None or p == 0: no dropout)."""
    d = _lib.Dropout()
    if state is not None and p > 0:
        _req(state, torch.int64, "dropout state")
        d.state, d.p, d.layer = ptr(state), float(p), int(layer)
        d._keep = state
    else:
        d.state, d.p, d.layer = None, 0.0, 0
    return d


def _drop_ref(drop):
    import ctypes
    return ctypes.byref(drop) if drop is not None and drop.state else None


def dropout_mask(state, p, layer, rows, cols):
    """The keep mask (uint8 [rows, cols], 1 = kept) of ``layer`` at the state's CURRENT step:
    what the kernels evaluate in their epilogues (tests replay it in the oracle)."""
    import ctypes
    mask = torch.empty((rows, cols), dtype=torch.uint8, device=state.device)
    d = dropout_spec(state, p, layer)
    check(_lib.lib().abn_dropout_mask(ctypes.byref(d), rows, cols, ptr(mask), stream_ptr()))
    return mask


def linear_forward(x, W, b, act, precision=0, out=None, drop=None, row_offset=0):
    _req(x, torch.float32, "x")
    _req(W, torch.float32, "W")
    m, n_in = x.shape
    n_out = W.shape[0]
    y = out if out is not None else torch.empty((m, n_out), dtype=torch.float32, device=x.device)
    check(_lib.lib().abn_linear_forward_drop(ptr(x), ptr(W), ptr(b), m, n_in, n_out, ACT[act],
                                             precision, ptr(y), _drop_ref(drop), int(row_offset),
                                             stream_ptr()))
    return y


def linear_backward(x, W, y, dy, act, precision=0, need_dx=True, dW=None, db=None,
                    accumulate=False, dx=None, accumulate_dx=False, drop=None, row_offset=0):
    """dy is overwritten with dz = dy * act'(y).  Returns (dx, dW, db)."""
    _req(dy, torch.float32, "dy")
    m, n_in = x.shape
    n_out = W.shape[0]
    if dx is None:
        dx = torch.empty_like(x) if need_dx else None
        accumulate_dx = False
    if dW is None:
        dW = torch.empty_like(W)
        db = torch.empty(n_out, dtype=torch.float32, device=x.device)
        accumulate = False
    check(_lib.lib().abn_linear_backward_drop(ptr(x), ptr(W), ptr(y), ptr(dy), m, n_in, n_out,
                                              ACT[act], precision,
                                              int(bool(accumulate)) | (2 if accumulate_dx else 0),
                                              ptr(dx), ptr(dW), ptr(db), _drop_ref(drop),
                                              int(row_offset), stream_ptr()))
    return dx, dW, db


def optimizer_step(param, grad, state0, state1, kind, lr, momentum=0.0, grad_scale=1.0,
                   step=1):
    _req(param, torch.float32, "param")
    _req(grad, torch.float32, "grad")
    check(_lib.lib().abn_optimizer_step(ptr(param), ptr(grad), ptr(state0), ptr(state1),
                                        param.numel(), OPT_KIND[kind], float(lr),
                                        float(momentum), float(grad_scale), int(step),
                                        stream_ptr()))


# ------------------------------------------------------- tensor-core path ---
def pad8(n):
    return (n + 7) // 8 * 8


def pad_row(n):
    """Leading dimension (bf16 elements) of a GEMM operand / output array: rows start on
    128-byte lines.  The kernel only needs a multiple of 8; with 1008-byte rows (504 elements)
    every 128-byte TMA row segment straddles two L2 lines and a half sector -- measured
    40.7 -> 32.3 us for the forward chain (tools/time_chain.py, PADTO=8 vs 64)."""
    return (n + 63) // 64 * 64


def cast_bf16(src, dst=None, dstT=None):
    """fp32 [rows, cols] -> bf16 copy and/or transposed bf16 copy (padded leading dims)."""
    _req(src, torch.float32, "src")
    rows, cols = src.shape
    check(_lib.lib().abn_cast_bf16(ptr(src), rows, cols, src.stride(0), ptr(dst),
                                   dst.stride(0) if dst is not None else 0, ptr(dstT),
                                   dstT.stride(0) if dstT is not None else 0, stream_ptr()))


# ------------------------------------- persistent grouped tcgen05 GEMM ---
GE_BIAS_ACT, GE_DACT, GE_ATOMIC = 0, 1, 2
GEMM_MAX_GROUP = 8


def gemm_problem(A, B, M, N, K, epilogue, out, a_mn=False, b_mn=False, act=None, bias=None,
                 yprev=None, split_k=1, ones_col=False, ones_out=None, signal=None, wait=None,
                 wait_count=0):
    """One abn_gemm_problem.  A / B: bf16 2-D CUDA tensors with a contiguous last
    dimension ([M, K] / [N, K], or [K, M] / [K, N] when a_mn / b_mn); ``out``: bf16 or
    float32 2-D tensor.  Keeps references to the tensors alive on the returned object."""
    for t, nm in ((A, "A"), (B, "B")):
        if not (t.is_cuda and t.dtype == torch.bfloat16 and t.stride(-1) == 1):
            raise TypeError("%s must be a CUDA bf16 tensor with a contiguous last dim" % nm)
    if not (out.is_cuda and out.stride(-1) == 1 and out.dtype in (torch.float32, torch.bfloat16)):
        raise TypeError("out must be a CUDA float32 / bf16 tensor with a contiguous last dim")
    q = _lib.GemmProblem()
    q.A, q.lda, q.a_mn = ptr(A), A.stride(0), int(bool(a_mn))
    q.B, q.ldb, q.b_mn = ptr(B), B.stride(0), int(bool(b_mn))
    q.M, q.N, q.K = int(M), int(N), int(K)
    q.epilogue, q.act, q.split_k = int(epilogue), ACT[act], int(split_k)
    q.bias = ptr(bias)
    q.out, q.ldo, q.out_f32 = ptr(out), out.stride(0), int(out.dtype == torch.float32)
    q.yprev, q.ld_yprev = ptr(yprev), (yprev.stride(0) if yprev is not None else 0)
    q.ones_col = int(bool(ones_col))
    q.ones_out = ptr(ones_out)
    q.signal, q.wait, q.wait_count = ptr(signal), ptr(wait), int(wait_count)
    q._keep = (A, B, out, bias, yprev, ones_out, signal, wait)
    return q


def gemm_group(problems):
    """Run up to GEMM_MAX_GROUP problems in ONE persistent tcgen05 launch."""
    n = len(problems)
    arr = (_lib.GemmProblem * n)(*problems)
    check(_lib.lib().abn_gemm_bf16_group(arr, n, stream_ptr()))


MLP_MAX_LAYERS = 8
MLP_MAX_WIDTH = 512


def mlp_layers(specs):
    """specs: [(W bf16 [n_out, ld], n_in, bias fp32 or None, act, out tensor, ones_col
    [, (dropout state, p, layer) or None])] -> ctypes array of abn_mlp_layer (keeps the tensors
    alive)."""
    arr = (_lib.MlpLayer * len(specs))()
    keep = []
    for a, spec in zip(arr, specs):
        W, n_in, bias, act, out, ones_col = spec[:6]
        a.W, a.ldw, a.bias = ptr(W), W.stride(0), ptr(bias)
        a.n_in, a.n_out, a.act = int(n_in), int(W.shape[0]), ACT[act]
        a.out, a.ldo, a.out_f32 = ptr(out), out.stride(0), int(out.dtype == torch.float32)
        a.ones_col = int(bool(ones_col))
        if len(spec) > 6 and spec[6] is not None:          # (state, p, layer): dropout
            a.drop = dropout_spec(*spec[6])
            keep.append(spec[6][0])
        keep += [W, bias, out]
    arr._keep = keep
    return arr


def mlp_forward_fused(x, rows, layers):
    """Every layer of the forward pass in ONE launch, activations resident in shared memory
    (abn_mlp_forward_fused); x: bf16 [rows, ld]."""
    check(_lib.lib().abn_mlp_forward_fused(ptr(x), x.stride(0), int(rows), layers, len(layers),
                                           stream_ptr()))


def mlp_forward_loss_fused(x, rows, layers, y, dz, kind="coscos2", margin=0.5, scale=1.0,
                           loss_out=None, write_embeddings=False):
    """mlp_forward_fused on INTERLEAVED pair rows (x[2k], x[2k + 1] = the two frames of pair k) with
    the pair loss (abnet3/loss.py:46-67, :85-105) computed in the last layer's epilogue: the loss
    is accumulated into ``loss_out`` and dz of the output layer written to ``dz`` (bf16
    [rows, ld]); the fp32 embeddings go to the last layer's ``out`` only on request."""
    _req(y, torch.float32, "y")
    if y.numel() * 2 < rows:
        raise ValueError("one label per pair of rows expected")
    if not (dz.is_cuda and dz.dtype == torch.bfloat16 and dz.stride(-1) == 1 and dz.shape[0] >= rows):
        raise TypeError("dz must be a CUDA bf16 [>= rows, ld] tensor")
    loss = loss_out if loss_out is not None else torch.zeros(1, dtype=torch.float32, device=x.device)
    spec = _lib.MlpLoss(ptr(y), ptr(loss), ptr(dz), dz.stride(0), LOSS_KIND[kind], float(margin),
                        float(scale), int(bool(write_embeddings)))
    check(_lib.lib().abn_mlp_forward_loss_fused(ptr(x), x.stride(0), int(rows), layers, len(layers),
                                                _c.byref(spec), stream_ptr()))
    return loss


def mlp_dlayers(specs):
    """specs, top layer first: [(W bf16 [n_out, ld], n_in, act_below, y_below, dz_below)] ->
    ctypes array of abn_mlp_dlayer."""
    arr = (_lib.MlpDLayer * len(specs))()
    keep = []
    for a, spec in zip(arr, specs):
        W, n_in, act_below, y_below, dz_below = spec[:5]
        a.W, a.ldw, a.n_in, a.n_out = ptr(W), W.stride(0), int(n_in), int(W.shape[0])
        a.act_below = ACT[act_below]
        a.y_below, a.ld_y = ptr(y_below), y_below.stride(0)
        a.dz_below, a.ld_dz = ptr(dz_below), dz_below.stride(0)
        if len(spec) > 5 and spec[5] is not None:          # dropout of the layer below
            a.drop_below = dropout_spec(*spec[5])
            keep.append(spec[5][0])
        keep += [W, y_below, dz_below]
    arr._keep = keep
    return arr


def mlp_dgrad_fused(dz_top, rows, layers):
    """dz of every layer below the top one in ONE launch (abn_mlp_dgrad_fused)."""
    check(_lib.lib().abn_mlp_dgrad_fused(ptr(dz_top), dz_top.stride(0), int(rows), layers,
                                         len(layers), stream_ptr()))


# ------------------------------- fused companions of the tensor-core step ---
def gather_batch_bf16(feat, idx1, idx2, y, sel, n, xb, y_out=None, zero=None, y2=None, y2_out=None,
                      cursor=None, loss_acc=None, table_rows=0, interleave=False):
    """xb[:n] = bf16(feat[idx1[pos]]), xb[n:2n] = bf16(feat[idx2[pos]]) -- or, ``interleave``,
    xb[2k] and xb[2k + 1], the layout of mlp_forward_loss_fused --, y_out = float(y[pos]) (and
    y2_out = float(y2[pos])) with pos = sel[k], or cursor[0] + k when ``sel`` is None and a
    ``cursor`` (device int64 [2]) is given -- the kernel then advances cursor[0] by n.  ``zero``
    (contiguous 4-byte-element tensor) is cleared by the same kernel, after its word 0 (the
    previous step's loss) has been added to ``loss_acc`` (device float64 [1]) when given.
    ``table_rows`` > 0: in cursor mode a batch past the end of the table is not gathered."""
    _req(feat, torch.float32, "feat")
    _req(idx1, torch.int32, "idx1")
    _req(idx2, torch.int32, "idx2")
    if sel is not None:
        _req(sel, torch.int64, "sel")
    if cursor is not None:
        _req(cursor, torch.int64, "cursor")
    if loss_acc is not None:
        _req(loss_acc, torch.float64, "loss_acc")
    for lab, nm in ((y, "y"), (y2, "y2")):
        if lab is not None:
            _req(lab, torch.int8, nm)
    if not (xb.is_cuda and xb.dtype == torch.bfloat16 and xb.stride(-1) == 1 and xb.shape[0] >= 2 * n):
        raise TypeError("xb must be a CUDA bf16 [>= 2n, ld] tensor")
    check(_lib.lib().abn_gather_step_bf16(ptr(feat), feat.shape[1], ptr(idx1), ptr(idx2), ptr(y), ptr(y2),
                                          ptr(sel), ptr(cursor), int(table_rows), n, ptr(xb),
                                          xb.stride(0), ptr(y_out),
                                          ptr(y2_out), ptr(zero),
                                          zero.numel() if zero is not None else 0, ptr(loss_acc),
                                          int(bool(interleave)), stream_ptr()))


def pair_loss_dz(e1, e2, y, dz1, dz2, kind="coscos2", margin=0.5, scale=1.0, act=None,
                 loss_out=None, drop=None, row2_offset=0, col_offset=0):
    """Loss value (accumulated into loss_out) and dz = dL/de * act'(e) as bf16 rows."""
    _req(y, torch.float32, "y")
    n, dim = e1.shape
    ld = e1.stride(0)
    if e2.stride(0) != ld or dz1.stride(0) != dz2.stride(0):
        raise ValueError("e1/e2 and dz1/dz2 must share their row strides")
    loss = loss_out if loss_out is not None else torch.zeros(1, dtype=torch.float32, device=e1.device)
    check(_lib.lib().abn_pair_loss_dz_drop(ptr(e1), ptr(e2), ptr(y), n, dim, ld, LOSS_KIND[kind],
                                           float(margin), float(scale), ACT[act], ptr(loss), ptr(dz1),
                                           ptr(dz2), dz1.stride(0), _drop_ref(drop), int(row2_offset),
                                           int(col_offset), stream_ptr()))
    return loss


def param_segments(entries):
    """entries: [(offset, count, bf16 tensor [rows, ld] or None, n_in)] -> ctypes array."""
    arr = (_lib.ParamSegment * len(entries))()
    for a, (off, cnt, wb, n_in) in zip(arr, entries):
        a.offset, a.count = int(off), int(cnt)
        a.ld = wb.stride(0) if wb is not None else 0
        a.bf16 = ptr(wb)
        a.n_in = int(n_in)
    return arr


def optimizer_step_fused(param, grad, state0, state1, kind, lr, momentum, grad_scale, step,
                         segments, zero_grad=True):
    _req(param, torch.float32, "param")
    _req(grad, torch.float32, "grad")
    check(_lib.lib().abn_optimizer_step_fused(ptr(param), ptr(grad), ptr(state0), ptr(state1),
                                              OPT_KIND[kind], float(lr), float(momentum),
                                              float(grad_scale), int(step), segments,
                                              len(segments), int(bool(zero_grad)), stream_ptr()))


# ------------------------- data parallelism over NVLink peer memory (no NCCL) ---
def dp_setup(grad, group=None):
    """Map every rank's gradient bucket and flag block into this process (CUDA IPC).
    Collective over ``group``; returns an opaque peers object for dp_optimizer_step /
    dp_grad_reset, or raises if the buffers cannot be shared (then use NCCL)."""
    import ctypes
    import torch.distributed as dist
    lib = _lib.lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world > 8:
        raise RuntimeError("peer-memory data parallelism serves one box (world <= 8)")
    flags = torch.zeros(32, dtype=torch.int64, device=grad.device)
    torch.cuda.synchronize()

    def export(t):
        h = (ctypes.c_ubyte * 64)()
        off = ctypes.c_int64(0)
        check(lib.abn_ipc_export(ptr(t), h, ctypes.byref(off)))
        return bytes(h), int(off.value)

    mine = {"grad": export(grad), "flags": export(flags), "numel": grad.numel()}
    everyone = [None] * world
    dist.all_gather_object(everyone, mine, group=group)
    peers = _lib.DpPeers()
    peers.rank, peers.world = rank, world
    for r, info in enumerate(everyone):
        if info["numel"] != grad.numel():
            raise RuntimeError("rank %d has a different gradient bucket" % r)
        if r == rank:
            peers.grad[r], peers.flags[r] = ptr(grad), ptr(flags)
            continue
        for key, arr in (("grad", peers.grad), ("flags", peers.flags)):
            hb, off = info[key]
            h = (ctypes.c_ubyte * 64).from_buffer_copy(hb)
            out = ctypes.c_void_p()
            check(lib.abn_ipc_import(h, off, ctypes.byref(out)))
            arr[r] = out.value
    dist.barrier(group=group)
    peers._keep = (grad, flags)
    return peers


def dp_optimizer_step(param, state0, state1, kind, lr, momentum, grad_scale, step, segments, peers):
    import ctypes
    check(_lib.lib().abn_dp_optimizer_step(ptr(param), ptr(state0), ptr(state1), OPT_KIND[kind],
                                           float(lr), float(momentum), float(grad_scale), int(step),
                                           segments, len(segments), ctypes.byref(peers), stream_ptr()))


def dp_grad_reset(grad, peers):
    import ctypes
    check(_lib.lib().abn_dp_grad_reset(ptr(grad), grad.numel(), ctypes.byref(peers), stream_ptr()))


def dp_push_setup(param, n_trained, group=None, one_shot=True, ll=False):
    """Write-only exchange over NVLink peer memory: share the parameter bucket, a receive buffer
    and a flag block with every rank of ``group`` (CUDA IPC).  Collective.  one_shot: every rank
    pushes its whole gradient bucket to every peer (one flag exchange); otherwise the two-shot
    reduce-scatter / all-gather form."""
    import ctypes
    import torch.distributed as dist
    lib = _lib.lib()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world > 8:
        raise RuntimeError("peer-memory data parallelism serves one box (world <= 8)")
    if n_trained % 4:
        raise RuntimeError("the trained parameter count must be a multiple of 4")
    if ll:          # flag-in-data two-shot: uint64 gradient inbox [world, cap] + parameter inbox [n]
        cap = ((n_trained + world - 1) // world + 3) // 4 * 4
        recv = torch.zeros(2 * (world * cap + n_trained), dtype=torch.float32, device=param.device)
        one_shot = 2
    elif one_shot:
        cap = (n_trained + 3) // 4 * 4
        recv = torch.zeros(2 * world * cap, dtype=torch.float32, device=param.device)
    else:
        cap = ((n_trained + world - 1) // world + 3) // 4 * 4
        recv = torch.zeros(world * cap, dtype=torch.float32, device=param.device)
    flags = torch.zeros(32, dtype=torch.int64, device=param.device)
    torch.cuda.synchronize()

    def export(t):
        h = (ctypes.c_ubyte * 64)()
        off = ctypes.c_int64(0)
        check(lib.abn_ipc_export(ptr(t), h, ctypes.byref(off)))
        return bytes(h), int(off.value)

    mine = {"param": export(param), "recv": export(recv), "flags": export(flags), "n": n_trained,
            "one_shot": int(one_shot)}
    everyone = [None] * world
    dist.all_gather_object(everyone, mine, group=group)
    pp = _lib.DpPush()
    pp.rank, pp.world, pp.n, pp.slice_cap, pp.one_shot = rank, world, n_trained, cap, int(one_shot)
    local = {"param": param, "recv": recv, "flags": flags}
    for r, info in enumerate(everyone):
        if info["n"] != n_trained or info["one_shot"] != int(one_shot):
            raise RuntimeError("rank %d set the exchange up differently" % r)
        for key, arr in (("param", pp.param), ("recv", pp.recv), ("flags", pp.flags)):
            if r == rank:
                arr[r] = ptr(local[key])
                continue
            hb, off = info[key]
            h = (ctypes.c_ubyte * 64).from_buffer_copy(hb)
            out = ctypes.c_void_p()
            check(lib.abn_ipc_import(h, off, ctypes.byref(out)))
            arr[r] = out.value
    dist.barrier(group=group)
    pp._keep = (param, recv, flags)
    return pp


def dp_push_step(grad, state0, state1, kind, lr, momentum, grad_scale, step, segments, pp):
    import ctypes
    check(_lib.lib().abn_dp_push_step(ptr(grad), ptr(state0), ptr(state1), OPT_KIND[kind], float(lr),
                                      float(momentum), float(grad_scale), int(step), segments,
                                      len(segments), ctypes.byref(pp), stream_ptr()))
