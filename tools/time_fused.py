"""Time the fused forward / dgrad kernels (activations resident in shared memory) against the chained
grouped-GEMM launches on the training-step shapes (graph-replayed, L2-warm)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops
DEV = "cuda"
rows = 16384

def bf(r, c):
    return (torch.randn(r, ops.pad_row(c + 1), device=DEV) * 0.05).bfloat16()

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

dims = [280, 500, 500, 500, 100]
acts = [bf(rows, d) for d in dims]
dzs = [bf(rows, d) for d in dims]
Ws = [bf(dims[i + 1], dims[i]) for i in range(4)]
bias = [torch.zeros(dims[i + 1], device=DEV) for i in range(4)]
out_last = torch.zeros(rows, 100, device=DEV)
dep = torch.zeros((12, (rows + 255) // 256), dtype=torch.int32, device=DEV)
fw = [ops.gemm_problem(acts[l], Ws[l], rows, dims[l + 1], dims[l], ops.GE_BIAS_ACT, acts[l + 1] if l < 3 else out_last,
                       act="sigmoid", bias=bias[l], ones_col=(l < 3), signal=dep[l] if l < 3 else None,
                       wait=dep[l - 1] if l > 0 else None) for l in range(4)]
def run_fw():
    dep.zero_(); ops.gemm_group(fw)
fl = ops.mlp_layers([(Ws[l], dims[l], bias[l], "sigmoid", acts[l + 1] if l < 3 else out_last, l < 3) for l in range(4)])
print("forward: chained grouped GEMM %6.1f us | fused %6.1f us" % (timeit(run_fw), timeit(lambda: ops.mlp_forward_fused(acts[0], rows, fl))), flush=True)
dg, k = [], 0
for l in range(3, 0, -1):
    dg.append(ops.gemm_problem(dzs[l + 1], Ws[l], rows, dims[l], dims[l + 1], ops.GE_DACT, dzs[l], b_mn=True, act="sigmoid",
                               yprev=acts[l], signal=dep[4 + k] if l > 1 else None, wait=dep[4 + k - 1] if k > 0 else None))
    k += 1
def run_dg():
    dep.zero_(); ops.gemm_group(dg)
dl = ops.mlp_dlayers([(Ws[l], dims[l], "sigmoid", acts[l], dzs[l]) for l in range(3, 0, -1)])
print("dgrad:   chained grouped GEMM %6.1f us | fused %6.1f us" % (timeit(run_dg), timeit(lambda: ops.mlp_dgrad_fused(dzs[4], rows, dl))), flush=True)
