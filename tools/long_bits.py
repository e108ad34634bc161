"""Bit checksum of the long path's distances and costs (ABN_LONG_R forces the tile class)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops, synth
n = int(sys.argv[1])
c = synth.make_corpus(300, seed=n, device="cuda", len_range=(n, n), tokens_per_file=50)
pairs = synth.make_same_pairs(c, 200, seed=1)
for stack in (7, 0):
    d, off, v = ops.cosine_distance(c.feat, pairs, stack=stack)
    r = ops.align_pairs(c.feat, pairs, stack=stack)
    torch.cuda.synchronize()
    print("n=%d R=%s stack=%d dist bits %d cost bits %d plen %d" % (
        n, os.environ.get("ABN_LONG_R", "auto"), stack, int(d.view(torch.int32).long().sum()),
        int(r.cost.view(torch.int64).sum() % (1 << 40)), int(r.path_len.sum())))
