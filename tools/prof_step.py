"""A few eager bf16 training steps (step_gather) on synthetic frame pairs, for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.model import SiameseNetwork
dev = "cuda"
torch.manual_seed(0)
B = 8192
feat = torch.randn(400000, 280, device=dev)
n_fp = 2_000_000
idx1 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32)
idx2 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32)
y = (torch.randint(0, 2, (n_fp,), device=dev) * 2 - 1).to(torch.int8)
net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
                     activation_layer="sigmoid", precision="bf16").to(dev)
step = SiameseTrainStep(net, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
sel = step.gather_buffers(B)
graph = len(sys.argv) > 1 and sys.argv[1] == "graph"
for i in range(6):
    sel.copy_(torch.randperm(n_fp, device=dev)[:B])
    loss = step.step_gather(feat, idx1, idx2, y, B, graph=graph)
torch.cuda.synchronize()
print("loss", float(loss))
