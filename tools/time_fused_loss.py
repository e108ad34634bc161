"""us per launch: forward chain alone, forward + loss kernel, forward with the loss in its epilogue
(and each followed by the dgrad chain), 30 launches per CUDA graph."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.model import SiameseNetwork
dev = "cuda"
torch.manual_seed(0)
B = 8192
feat = torch.randn(500_000, 280, device=dev)
n_fp = 1_000_000
idx1 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32)
idx2 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32)
y = (torch.randint(0, 2, (n_fp,), device=dev) * 2 - 1).to(torch.int8)
net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
                     activation_layer="sigmoid").to(dev)
eng = SiameseTrainStep(net, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
sel = eng.gather_buffers(B)
sel.copy_(torch.arange(B, device=dev))
eng._table_step(feat, (idx1, idx2, y), B, sel, True, graph=False)
torch.cuda.synchronize()


def t(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / iters * 1e3)
    return best


o = eng.out_last
fwd = lambda: ops.mlp_forward_fused(eng.xb, eng._fwd_rows, eng._fwd_fused)
loss = lambda: eng._loss_and_seed_bf16(o, B, eng._gy)
fl = lambda emb=False: ops.mlp_forward_loss_fused(eng.xb, eng._fwd_rows, eng._fwd_fused, eng._gy[0], eng.dzb[-1],
                                                  "coscos2", 0.0, 1.0, loss_out=eng.loss_buf, write_embeddings=emb)
dg = lambda: ops.mlp_dgrad_fused(eng.dzb[-1], eng._fwd_rows, eng._dgrad_fused)
print("forward                  %.1f" % t(fwd))
print("forward, loss kernel     %.1f" % t(lambda: (fwd(), loss())))
print("forward+loss fused       %.1f" % t(fl))
print("forward+loss fused, emb  %.1f" % t(lambda: fl(True)))
print("dgrad                    %.1f" % t(dg))
print("forward, loss, dgrad     %.1f" % t(lambda: (fwd(), loss(), dg())))
print("fused, dgrad             %.1f" % t(lambda: (fl(), dg())))
print("forward, dgrad (no loss) %.1f" % t(lambda: (fwd(), dg())))
print("fused, loss kernel, dgrad %.1f" % t(lambda: (fl(), loss(), dg())))
z = torch.zeros(64, device=dev)
print("fused, memset, dgrad     %.1f" % t(lambda: (fl(), z.zero_(), dg())))
