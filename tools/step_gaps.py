"""Where the graph-replayed training step spends its time: cumulative prefixes of the kernel
sequence (gather | +forward | +loss | +dgrad | +wgrad | +optimizer), each captured as a graph of
10 repetitions -- the marginal cost of a stage IN SEQUENCE (with its launch gap), to compare with
the stand-alone kernel times of bench.py."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.model import SiameseNetwork
dev = "cuda"
torch.manual_seed(0)
B = 8192
feat = torch.randn(400000, 280, device=dev)
n_fp = 4_000_000
idx1 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32)
idx2 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32)
y = (torch.randint(0, 2, (n_fp,), device=dev) * 2 - 1).to(torch.int8)
net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
                     activation_layer="sigmoid").to(dev)
eng = SiameseTrainStep(net, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
table = (idx1, idx2, y)
for _ in range(4):
    eng.sweep_table(feat, table, B, 3, start=0)
torch.cuda.synchronize()
sel = eng.gather_buffers(B)
sel.copy_(torch.arange(B, device=dev))

def stages():
    fused = eng._fuse_loss()
    def gather():
        ops.gather_batch_bf16(feat, idx1, idx2, y, sel, B, eng.xb, y_out=eng._gy[0], zero=eng._zbuf,
                              interleave=fused)
    def fwd():
        if fused:
            eng._fwd_loss(B)            # forward chain with the loss in its last epilogue
        else:
            ops.mlp_forward_fused(eng.xb, eng._fwd_rows, eng._fwd_fused)
    def loss():
        if not fused:
            eng._loss_cleared = True
            eng._loss_and_seed(eng.out_last, B, eng._gy)
    def dgrad():
        ops.mlp_dgrad_fused(eng.dzb[-1], eng._fwd_rows, eng._dgrad_fused)
    def wgrad():
        for grp in eng._backward_groups:
            ops.gemm_group(grp)
    def opt():
        eng._optimizer(1.0, 1)
    return [("gather", gather), ("forward", fwd), ("loss", loss), ("dgrad", dgrad), ("wgrad", wgrad),
            ("optimizer", opt)]

def timeit(fns, reps=10):
    for _ in range(2):
        for f in fns: f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for f in fns: f()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (5 * reps) * 1e3

st = stages()
prev = 0.0
line = []
for k in range(1, len(st) + 1):
    t = timeit([f for _, f in st[:k]])
    line.append("%s +%.1f" % (st[k - 1][0], t - prev))
    prev = t
alone = " ".join("%s %.1f" % (nm, timeit([f])) for nm, f in st)
print("%-28s total %.1f us | in sequence: %s | alone: %s" % (" ".join(sys.argv[1:]) or "default", prev, " | ".join(line), alone), flush=True)
