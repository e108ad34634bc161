"""Run a few eager C3 training steps (for `ncu --metrics gpu__time_duration.sum`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.model import SiameseNetwork
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.manual_seed(0)
net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100,
                     p_dropout=0.0, activation_layer="sigmoid", precision=prec).cuda()
st = SiameseTrainStep(net, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
n = 8192
x = torch.randn(2 * n, 280, device="cuda")
y = torch.where(torch.rand(n, device="cuda") < 0.5, 1.0, -1.0)
for _ in range(steps):
    st.step(x, n, y)
torch.cuda.synchronize()
print("done", float(st.loss_buf.item()))
