"""Time the graph-replayed bf16 training step (C3 shape) -- one line; env knobs are read by the engine."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
                     activation_layer="sigmoid", precision="bf16").to(dev)
step = SiameseTrainStep(net, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
sel = step.gather_buffers(B)
for i in range(6):
    sel.copy_(torch.arange(i * B, (i + 1) * B, device=dev))
    loss = step.step_gather(feat, idx1, idx2, y, B, graph=True)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(100):
        sel.copy_(torch.arange((i % 400) * B, (i % 400 + 1) * B, device=dev))
        loss = step.step_gather(feat, idx1, idx2, y, B, graph=True)
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 100 * 1e3)
print("%-50s step %6.1f us  loss %.2f" % (" ".join(sys.argv[1:]) or "-", best, float(loss)), flush=True)

if os.environ.get("PARTS") == "1":
    def timeit(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n): fn()
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): g.replay()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / (5 * n) * 1e3
    from abnet3_b200 import ops
    sy = step._sy if hasattr(step, "_sy") else None
    print("parts (20 back-to-back launches each, graph-replayed):")
    def ga():
        ops.gather_batch_bf16(feat, idx1, idx2, y, step._gsel, B, step.xb, y_out=step._gy, zero=step._zbuf)
    t_ga = timeit(ga)
    print("  gather             %6.1f us" % t_ga)
    def fw():
        ga(); step._loss_cleared = True; step._forward_bf16(None)
    print("  forward            %6.1f us  (gather + forward, minus gather)" % (timeit(fw) - t_ga))
    def ls():
        step._loss_cleared = True
        step._loss_and_seed(step.out_last, B, [step._gy])
    print("  loss->dz           %6.1f us" % timeit(ls))
    def bw():
        ga(); step._grads_clean = True; step._backward_bf16(None)
    print("  backward           %6.1f us  (gather + backward, minus gather)" % (timeit(bw) - t_ga))
    print("  optimizer          %6.1f us" % timeit(lambda: step._optimizer(1.0, 1)))
