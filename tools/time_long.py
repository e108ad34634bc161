"""pairs/s of the long-token path at one token length (ABN_LONG_R forces the tile class)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops, synth, utils
n = int(sys.argv[1]); P = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
dev = "cuda"
c = synth.make_corpus(max(2000, 400000 // n), seed=n, device=dev, len_range=(n, n), tokens_per_file=max(50, 100000 // n))
pairs = synth.make_same_pairs(c, P, seed=1)
al = utils.BatchAligner(c.feat, max_pairs=P, max_frames=n, stack=7)
for _ in range(2): r = al.align(pairs)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3): r = al.align(pairs)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print("n=%d R=%s: %.1f ms, %.0f pairs/s, cost checksum %.6f" % (n, os.environ.get("ABN_LONG_R", "auto"), ms, P / ms * 1e3, float(r.cost.sum())), flush=True)
