"""Device timeline (globaltimer stamps of abn_tc3.cu) of forward [+ loss kernel] -> dgrad against
forward-with-loss -> dgrad, inside one CUDA graph of 10 rounds (the last round's stamps survive)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops, _lib
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
hook = ctypes.c_void_p.in_dll(_lib.lib(), "abn_gemm_trace_buffer")
o = eng.out_last


def traced(buf, fn):
    hook.value = buf.data_ptr()
    fn()
    hook.value = None


def timeline(first, label):
    ta = torch.zeros(148 * 8 * 16, dtype=torch.int64, device=dev)
    tb = torch.zeros(148 * 8 * 16, dtype=torch.int64, device=dev)
    dg = lambda: ops.mlp_dgrad_fused(eng.dzb[-1], eng._fwd_rows, eng._dgrad_fused)

    def seq():
        traced(ta, first)
        if label == "forward, loss kernel":
            eng._loss_and_seed_bf16(o, B, eng._gy)
        traced(tb, dg)
    for _ in range(3):
        seq()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            seq()
    g.replay(); g.replay()
    torch.cuda.synchronize()
    a, b = ta.cpu().view(148, 8, 16), tb.cpu().view(148, 8, 16)
    t0 = int(a[:, 7, 14][a[:, 7, 14] > 0].min())
    us = lambda v: (v[v > 0].double() - t0) / 1e3
    print("== %s -> dgrad" % label)
    for nm, t in (("first", a), ("dgrad", b)):
        ent, pro, go, end = us(t[:, 7, 12]), us(t[:, 7, 13]), us(t[:, 7, 14]), us(t[:, 7, 15])
        print("  %s: entry %.1f..%.1f (median %.1f) | prologue done %.1f..%.1f (median %.1f) | past griddepcontrol.wait %.1f..%.1f | exit %.1f..%.1f"
              % (nm, ent.min(), ent.max(), ent.median(), pro.min(), pro.max(), pro.median(), go.min(), go.max(), end.min(), end.max()))
        for l in range(4):
            if int(t[:, l].max()) == 0:
                continue
            e0, e1 = us(t[:, l, 8]), us(t[:, l, 10])
            m0 = us(t[:, l, 1])
            if min(e0.numel(), e1.numel(), m0.numel()) == 0:
                continue
            print("    layer %d: first MMA %.1f..%.1f | epilogue top %.1f..%.1f | acc0 read %.1f..%.1f"
                  % (l, m0.min(), m0.max(), e0.min(), e0.max(), e1.min(), e1.max()))


timeline(lambda: ops.mlp_forward_fused(eng.xb, eng._fwd_rows, eng._fwd_fused), "forward, loss kernel")
timeline(lambda: ops.mlp_forward_fused(eng.xb, eng._fwd_rows, eng._fwd_fused), "forward")
timeline(lambda: ops.mlp_forward_loss_fused(eng.xb, eng._fwd_rows, eng._fwd_fused, eng._gy[0], eng.dzb[-1],
                                            "coscos2", 0.0, 1.0, loss_out=eng.loss_buf), "forward+loss fused")
