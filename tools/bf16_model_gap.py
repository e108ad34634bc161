"""Print GPU-vs-rounding-model and model-vs-exact gaps of the bf16 step (tests/test_gpu_tc.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import nets as onets
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.model import SiameseNetwork
DEV = "cuda"
torch.manual_seed(5)
net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100,
                     p_dropout=0.0, activation_layer="sigmoid", precision="bf16").to(DEV)
eng = SiameseTrainStep(net, ("coscos2", 0.0, False), "sgd", lr=0.01, momentum=0.0)
n = 8192
x = torch.randn(2 * n, 280, device=DEV)
x[n:] = 0.6 * x[:n] + 0.8 * x[n:]
y = torch.where(torch.rand(n, device=DEV) < 0.5, 1.0, -1.0)
out = eng.forward(x); eng._loss_and_seed(out, n, [y]); eng.backward(x)
torch.cuda.synchronize()
sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
xc, yc = x.cpu(), y.cpu()
sd64 = {k: v.double().requires_grad_() for k, v in sd.items()}
e_ex = onets.siamese_forward_once(sd64, xc.double())
l_ex = onets.coscos2(e_ex[:n], e_ex[n:], yc.double(), avg=False); l_ex.backward()
e_md, l_md, g_md = onets.siamese_step_bf16_model(sd, xc, yc)
e_gpu, l_gpu = out.detach().cpu().double(), float(eng.loss_buf.item())
rel = lambda a, b: float((a - b).norm() / b.norm())
print("emb  gpu-model max abs %.2e | rel gpu-model %.2e model-exact %.2e gpu-exact %.2e" % (
    float((e_gpu - e_md).abs().max()), rel(e_gpu, e_md), rel(e_md, e_ex.detach()), rel(e_gpu, e_ex.detach())))
print("loss gpu %.4f model %.4f exact %.4f" % (l_gpu, float(l_md), float(l_ex.detach())))
for k, p in net.named_parameters():
    g = p.grad.detach().cpu().double()
    print("%-24s gpu-model %.2e | model-exact %.2e | gpu-exact %.2e" % (k, rel(g, g_md[k]), rel(g_md[k], sd64[k].grad), rel(g, sd64[k].grad)))
