"""Print how the GPU paths follow the live reference's loss trajectory (tests/golden/trajectory.npz)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_trajectory import _run
from oracle import trajectory as tj

gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "trajectory.npz"))
ref = gold["losses"]
feat = torch.from_numpy(tj.features()).cuda()
for prec, table in (("fp32", False), ("bf16", True), ("bf16", False)):
    losses, _ = _run(prec, feat, tj.batches(), tj.state_dict(), table)
    rel = np.abs(losses - ref) / ref
    sgn = (losses - ref) / ref
    print("%s table=%d: max %.3e at %d | mean %.3e | signed mean %.3e | last-20 gap %.3e | steps 0,50,100,200,299: %s"
          % (prec, table, rel.max(), rel.argmax(), rel.mean(), sgn.mean(),
             abs(losses[-20:].mean() - ref[-20:].mean()) / ref[-20:].mean(),
             " ".join("%.1f/%.1f" % (losses[k], ref[k]) for k in (0, 50, 100, 200, 299))))
