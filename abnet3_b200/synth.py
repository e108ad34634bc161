"""Synthetic Buckeye-shaped corpus.

There is no network and no speech data here, so benchmarks and tests run on a
seeded synthetic corpus with the SHAPE of the reference's input:

* 40-dim "filterbank" frames, AR(1)-smoothed along time so that neighbouring
  frames are correlated and DTW paths are not trivial;
* word *types* (clusters): every token of a type is a randomly time-warped
  (local rate +-30 %), noisy (SNR ~ 10 dB) rendition of the type's prototype,
  its length drawn uniformly from 20-80 frames, so 'same' pairs align
  meaningfully and no two frames are bit-identical;
* tokens laid out back to back in files, per-file mean/variance normalisation,
  then the 7-frame stack with zero padding at file edges exactly as
  /root/reference/abnet3/features.py:135-159 (`stack_fbanks`), giving
  feat[n_frames, 280] float32.

Data generation is not part of the hot path; it uses plain torch ops on
whatever device is asked for.
"""
from collections import namedtuple

import numpy as np
import torch

Corpus = namedtuple(
    "Corpus",
    "feat tok_start tok_len tok_cluster tok_file file_off cluster_members n_fbank stack")


def stack_frames(frames, file_id, nframes=7):
    """abnet3/features.py:135-159 applied per file on a concatenated corpus:
    row t = [x[t-3], ..., x[t+3]] with zeros outside the file."""
    T, dim = frames.shape
    half = nframes // 2
    out = torch.zeros((T, dim * nframes), dtype=frames.dtype, device=frames.device)
    t = torch.arange(T, device=frames.device)
    for k, s in enumerate(range(-half, half + 1)):
        src = t + s
        ok = (src >= 0) & (src < T)
        srcc = src.clamp(0, T - 1)
        ok &= file_id[srcc] == file_id
        out[:, k * dim:(k + 1) * dim] = torch.where(ok.unsqueeze(1), frames[srcc],
                                                    torch.zeros((), dtype=frames.dtype,
                                                                device=frames.device))
    return out


def make_corpus(n_tokens, cluster_size=16, len_range=(20, 80), n_fbank=40, stack=7,
                rho=0.9, snr_db=10.0, warp=0.3, tokens_per_file=2000, seed=0,
                device="cpu"):
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lo, hi = len_range
    n_clusters = (n_tokens + cluster_size - 1) // cluster_size
    PL = hi                                             # prototype buffer length
    f32 = torch.float32

    def ar1(shape_lead, length):
        x = torch.randn(*shape_lead, length, n_fbank, generator=g, device=dev, dtype=f32)
        c = float(np.sqrt(1.0 - rho * rho))
        for t in range(1, length):
            x[..., t, :] = rho * x[..., t - 1, :] + c * x[..., t, :]
        return x

    proto = ar1((n_clusters,), PL)                      # [C, PL, 40]
    proto_len = torch.randint(lo, hi + 1, (n_clusters,), generator=g, device=dev)

    # tokens in shuffled order so that renditions of one type are far apart
    tok_cluster = torch.arange(n_clusters * cluster_size, device=dev) // cluster_size
    tok_cluster = tok_cluster[torch.randperm(n_clusters * cluster_size, generator=g,
                                             device=dev)][:n_tokens]
    base = proto_len[tok_cluster].to(f32)
    # token lengths ~ U{lo..hi} (BASELINE.json: "20-80 frames"), independent of the type
    tok_len = torch.randint(lo, hi + 1, (n_tokens,), generator=g, device=dev)

    # random monotone warp: positive increments, normalised to [0, proto_len-1]
    pos_idx = torch.arange(hi, device=dev).unsqueeze(0)                     # [1, hi]
    live = pos_idx < tok_len.unsqueeze(1)                                   # [N, hi]
    inc = 1.0 + warp * (2.0 * torch.rand(n_tokens, hi, generator=g, device=dev) - 1.0)
    cum = torch.cumsum(inc * live, 1)
    first = cum[:, :1]
    last = torch.gather(cum, 1, (tok_len - 1).unsqueeze(1))
    denom = (last - first).clamp_min(1e-6)
    pos = (cum - first) / denom * (base - 1.0).unsqueeze(1)
    pos = pos.clamp_min(0.0)
    pos = torch.minimum(pos, (base - 1.0).unsqueeze(1))
    i0 = pos.floor().to(torch.int64)
    i1 = torch.minimum(i0 + 1, (proto_len[tok_cluster] - 1).unsqueeze(1))
    frac = (pos - i0.to(f32)).unsqueeze(2)

    chunk = 16384
    tok_frames = torch.empty(int(tok_len.sum().item()), n_fbank, dtype=f32, device=dev)
    tok_start = torch.zeros(n_tokens, dtype=torch.int64, device=dev)
    tok_start[1:] = torch.cumsum(tok_len, 0)[:-1]
    amp = float(10.0 ** (-snr_db / 20.0))
    for a in range(0, n_tokens, chunk):
        b = min(n_tokens, a + chunk)
        pc = proto[tok_cluster[a:b]]                                        # [n, PL, 40]
        x0 = torch.gather(pc, 1, i0[a:b].unsqueeze(2).expand(-1, -1, n_fbank))
        x1 = torch.gather(pc, 1, i1[a:b].unsqueeze(2).expand(-1, -1, n_fbank))
        x = x0 + (x1 - x0) * frac[a:b] + amp * ar1((b - a,), hi)
        tok_frames[int(tok_start[a].item()):int((tok_start[b - 1] + tok_len[b - 1]).item())] = \
            x[live[a:b]]

    # files, per-file mean/variance normalisation, stacking
    tok_file = torch.arange(n_tokens, device=dev) // tokens_per_file
    n_files = int(tok_file[-1].item()) + 1
    frame_file = torch.repeat_interleave(tok_file, tok_len)
    cnt = torch.zeros(n_files, device=dev, dtype=f32).index_add_(
        0, frame_file, torch.ones_like(frame_file, dtype=f32))
    mean = torch.zeros(n_files, n_fbank, device=dev, dtype=f32).index_add_(
        0, frame_file, tok_frames) / cnt.unsqueeze(1)
    tok_frames -= mean[frame_file]
    var = torch.zeros(n_files, n_fbank, device=dev, dtype=f32).index_add_(
        0, frame_file, tok_frames * tok_frames) / cnt.unsqueeze(1)
    tok_frames /= var.sqrt().clamp_min(1e-6)[frame_file]
    feat = stack_frames(tok_frames, frame_file, stack) if stack > 1 else tok_frames
    file_off = torch.zeros(n_files + 1, dtype=torch.int64, device=dev)
    file_off[1:] = torch.cumsum(cnt.to(torch.int64), 0)

    # cluster -> member token ids (padded with -1)
    order = torch.argsort(tok_cluster, stable=True)
    counts = torch.bincount(tok_cluster, minlength=n_clusters)
    members = torch.full((n_clusters, cluster_size), -1, dtype=torch.int64, device=dev)
    rank = torch.arange(n_tokens, device=dev) - torch.repeat_interleave(
        torch.cumsum(counts, 0) - counts, counts)
    members[tok_cluster[order], rank] = order
    return Corpus(feat.contiguous(), tok_start.to(torch.int32), tok_len.to(torch.int32),
                  tok_cluster.to(torch.int32), tok_file.to(torch.int32), file_off, members,
                  n_fbank, stack)


def make_same_pairs(corpus, n_pairs, seed=1):
    """[n_pairs, 4] int32 (start1, n1, start2, n2): two different renditions of
    one word type."""
    dev = corpus.feat.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    members = corpus.cluster_members
    size = (members >= 0).sum(1)
    good = torch.nonzero(size >= 2).squeeze(1)
    c = good[torch.randint(0, good.numel(), (n_pairs,), generator=g, device=dev)]
    sz = size[c]
    a = (torch.rand(n_pairs, generator=g, device=dev) * sz).long().clamp_max(sz - 1)
    b = (torch.rand(n_pairs, generator=g, device=dev) * (sz - 1)).long().clamp_max(sz - 2)
    b = b + (b >= a).long()
    ta, tb = members[c, a], members[c, b]
    return torch.stack([corpus.tok_start[ta], corpus.tok_len[ta],
                        corpus.tok_start[tb], corpus.tok_len[tb]], 1).to(torch.int32).contiguous()


def make_diff_pairs(corpus, n_pairs, seed=2):
    """Two tokens of different word types."""
    dev = corpus.feat.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n_tokens = corpus.tok_start.numel()
    ta = torch.randint(0, n_tokens, (n_pairs,), generator=g, device=dev)
    tb = torch.randint(0, n_tokens, (n_pairs,), generator=g, device=dev)
    clash = corpus.tok_cluster[ta] == corpus.tok_cluster[tb]
    tb = torch.where(clash, (tb + 1) % n_tokens, tb)     # neighbours differ in type w.h.p.
    return torch.stack([corpus.tok_start[ta], corpus.tok_len[ta],
                        corpus.tok_start[tb], corpus.tok_len[tb]], 1).to(torch.int32).contiguous()
