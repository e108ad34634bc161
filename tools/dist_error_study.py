"""Diagnostic (GPU box): element-wise error of the GPU cosine distance and of
the numpy float32 oracle against float64-exact arithmetic, and the resulting
DTW cost differences.  Not part of the product or the test-suite."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, oracle
from abnet3_b200 import ops, synth

c = synth.make_corpus(600, cluster_size=8, tokens_per_file=150, seed=3)
pairs = synth.make_same_pairs(c, 400, seed=4)
feat = c.feat.numpy()
dist, doff, valid = ops.cosine_distance(c.feat.cuda(), pairs.cuda())
res = ops.align_pairs(c.feat.cuda(), pairs.cuda())
torch.cuda.synchronize()
dist, doff, gcost = dist.cpu().numpy(), doff.cpu().numpy(), res.cost.cpu().numpy()

def exact(x, y):
    x = x.astype(np.float64); y = y.astype(np.float64)
    cs = x @ y.T / np.outer(np.sqrt((x ** 2).sum(1)), np.sqrt((y ** 2).sum(1)))
    return cs, np.arccos(np.clip(cs, -1, 1)) / np.pi

eg, en, rg, rn, rgn, cosmax = [], [], [], [], [], []
for p, (s1, n1, s2, n2) in enumerate(pairs.numpy().tolist()):
    x, y = feat[s1:s1 + n1], feat[s2:s2 + n2]
    cs, de = exact(x, y)
    dn = oracle.cosine_distance(x, y)
    dg = dist[doff[p]:doff[p + 1]].reshape(n1, n2).astype(np.float64)
    eg.append((dg - de).ravel()); en.append((dn - de).ravel())
    ce, cn = oracle.dtw(de)[0], oracle.dtw(dn)[0]
    rg.append((gcost[p] - ce) / ce); rn.append((cn - ce) / ce); rgn.append((gcost[p] - cn) / cn)
    cosmax.append(cs.max())
eg, en = np.concatenate(eg), np.concatenate(en)
for nm, e in (("gpu  - exact", eg), ("numpy - exact", en)):
    print("%s  elementwise: mean %+.3e  rms %.3e  max|.| %.3e" % (nm, e.mean(), np.sqrt((e ** 2).mean()), np.abs(e).max()))
for nm, r in (("gpu cost vs exact", rg), ("numpy cost vs exact", rn), ("gpu cost vs numpy", rgn)):
    r = np.array(r)
    print("%-20s rel: mean %+.3e median|.| %.3e p95 %.3e max %.3e" % (nm, r.mean(), np.median(np.abs(r)), np.percentile(np.abs(r), 95), np.abs(r).max()))
print("max cosine over pairs: median %.6f max %.6f" % (np.median(cosmax), np.max(cosmax)))
