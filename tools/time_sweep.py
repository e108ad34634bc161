"""us per training step of engine.sweep_table over a large random frame-pair table (C3 shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.model import SiameseNetwork
dev = "cuda"
torch.manual_seed(0)
B = 8192
feat = torch.randn(2_000_000, 280, device=dev)
n_fp = 40_000_000
idx1 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32)
idx2 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32)
y = (torch.randint(0, 2, (n_fp,), device=dev) * 2 - 1).to(torch.int8)
net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
                     activation_layer="sigmoid").to(dev)
eng = SiameseTrainStep(net, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
table = (idx1, idx2, y)
eng.sweep_table(feat, table, B, 20)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tot = eng.sweep_table(feat, table, B, 1500, start=rep * 1500 * B)
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 1500 * 1e3)
print("%-40s %.1f us/step  (loss %.1f)" % (" ".join(sys.argv[1:]) or "default", best, float(tot) / 1500), flush=True)
