"""Pipelined sweep (prefetch gather on a side stream) vs the plain one-stream sweep at the C3 batch
size: same weights after K steps?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.model import SiameseNetwork
dev = "cuda"
B = 8192
g = torch.Generator(device=dev).manual_seed(1)
feat = torch.randn(1_000_000, 280, device=dev, generator=g)
n_fp = 4_000_000
idx1 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32, generator=g)
idx2 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32, generator=g)
y = (torch.randint(0, 2, (n_fp,), device=dev, generator=g) * 2 - 1).to(torch.int8)
table = (idx1, idx2, y)
def run(pipe, K):
    os.environ["ABN_PIPELINE"] = pipe
    torch.manual_seed(0)
    net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
                         activation_layer="sigmoid").to(dev)
    eng = SiameseTrainStep(net, ("coscos2", 0.0, False), "sgd", lr=1e-3, momentum=0.9)
    tot = eng.sweep_table(feat, table, B, K)
    torch.cuda.synchronize()
    return eng.bucket.trained_param.clone(), float(tot)
for K in (5, 40, 200):
    pa, la = run("0", K)
    pb, lb = run("1", K)
    pc, lc = run("0", K)
    rel = lambda a, b: float((a - b).norm() / (b - p0).norm()) if False else float((a - b).norm() / b.norm())
    print("K=%d  loss plain %.3f pipelined %.3f plain again %.3f | param rel diff pipelined-plain %.2e, plain-plain %.2e"
          % (K, la, lb, lc, rel(pb, pa), rel(pc, pa)), flush=True)
