"""Phases of abnet3_b200.utils.align_pairs_host at the bench's C2 workload (1 M pairs), each timed
with a synchronize on both sides: where does the end-to-end time go?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops, synth, utils
dev = torch.device("cuda", 0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
corpus = synth.make_corpus(40_000, seed=0, device=dev)
pairs = synth.make_same_pairs(corpus, P, seed=1)
feat = corpus.feat
last = torch.zeros(feat.shape[0], dtype=torch.uint8, device=dev)
last[(corpus.file_off[1:] - 1).long()] = 1
host_feat = torch.empty(feat.shape, dtype=feat.dtype, pin_memory=True); host_feat.copy_(feat)
host_pairs = torch.empty(pairs.shape, dtype=pairs.dtype, pin_memory=True); host_pairs.copy_(pairs)
host_last = torch.empty(last.shape, dtype=last.dtype, pin_memory=True); host_last.copy_(last)
max_frames = int(pairs[:, [1, 3]].max().item())
torch.cuda.synchronize()

def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, r

ms_up, f2 = t(lambda: ops.stack_upload(host_feat, 7, host_last.to(dev, non_blocking=True)))
print("upload (middle blocks, %d MB) + rebuild  %7.2f ms" % (host_feat.numel() * 4 // 7 >> 20, ms_up))
ms_full, _ = t(lambda: host_feat.to(dev, non_blocking=True))
print("upload (full table, %d MB)               %7.2f ms" % (host_feat.numel() * 4 >> 20, ms_full))
tok = host_pairs.to(dev)
ms_al, res = t(lambda: ops.align_pairs(f2, tok, max_frames=max_frames, stack=7))
print("align_pairs                               %7.2f ms" % ms_al)
ms_cp, (d1, d2, doff) = t(lambda: ops.compact_paths(res))
print("compact_paths                             %7.2f ms" % ms_cp)
h1 = torch.empty(d1.shape, dtype=d1.dtype, pin_memory=True); h2 = torch.empty(d2.shape, dtype=d2.dtype, pin_memory=True)
def back():
    h1.copy_(d1, non_blocking=True); h2.copy_(d2, non_blocking=True)
ms_b, _ = t(back)
print("copy back (%d MB)                        %7.2f ms" % ((d1.numel() + d2.numel()) * 4 >> 20, ms_b))
for ch in (1, 2, 4, 8):
    ms, _ = t(lambda: utils.align_pairs_host(host_feat, host_pairs, max_frames=max_frames, stack=7,
                                             last_row_of_file=host_last, chunks=ch))
    print("align_pairs_host chunks=%d                 %7.2f ms  = %.1f M pairs/s" % (ch, ms, P / ms / 1e3))

# does the copy-back overlap the next chunk's alignment?  (device table, no upload)
side = torch.cuda.Stream()
hh1 = torch.empty(d1.numel() + 16, dtype=torch.int32, pin_memory=True)
hh2 = torch.empty(d1.numel() + 16, dtype=torch.int32, pin_memory=True)
def chunked(n_chunks, copy):
    main = torch.cuda.current_stream()
    bounds = [P * c // n_chunks for c in range(n_chunks + 1)]
    base, keep = 0, []
    for c in range(n_chunks):
        r = ops.align_pairs(f2, tok[bounds[c]:bounds[c + 1]], max_frames=max_frames, stack=7)
        a, b, off = ops.compact_paths(r)
        n = a.numel()
        if copy:
            ev = torch.cuda.Event(); ev.record(main); side.wait_event(ev)
            with torch.cuda.stream(side):
                hh1[base:base + n].copy_(a, non_blocking=True); hh2[base:base + n].copy_(b, non_blocking=True)
        keep.append((r, a, b)); base += n
    side.synchronize(); main.synchronize()
for ch in (1, 2, 4):
    for copy in (False, True):
        ms, _ = t(lambda: chunked(ch, copy))
        print("device table, chunks=%d, copy back %-5s     %7.2f ms" % (ch, copy, ms))
# the same with ONE reusable aligner (no per-call allocations)
al = utils.BatchAligner(f2, max_pairs=P, max_frames=max_frames, stack=7)
ms, _ = t(lambda: al.align(tok))
print("BatchAligner.align (resident buffers)     %7.2f ms" % ms)
