"""Time the grouped tcgen05 GEMM on the training-step shapes (CUDA events, L2-warm)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops
DEV = "cuda"
rows = 16384

def bf(r, c):
    return (torch.randn(r, ops.pad8(c + 1), device=DEV) * 0.05).bfloat16()

def timeit(fn, n=20):
    """GPU time per call: n calls captured in one CUDA graph (no host launch cost)."""
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
gW = [torch.zeros(dims[i + 1], dims[i], device=DEV) for i in range(4)]
gb = [torch.zeros(dims[i + 1], device=DEV) for i in range(4)]
out_last = torch.zeros(rows, 100, device=DEV)
tot = 0
for l in range(4):
    n_in, n_out = dims[l], dims[l + 1]
    out = acts[l + 1] if l < 3 else out_last
    p = ops.gemm_problem(acts[l], Ws[l], rows, n_out, n_in, ops.GE_BIAS_ACT, out, act="sigmoid", bias=bias[l], ones_col=(l < 3))
    us = timeit(lambda: ops.gemm_group([p]))
    tot += us
    print("fwd   L%d %4dx%4d  %7.1f us  %6.1f TFLOP/s" % (l, n_in, n_out, us, 2.0 * rows * n_in * n_out / us / 1e6))
for l in range(3, 0, -1):
    n_in, n_out = dims[l], dims[l + 1]
    p = ops.gemm_problem(dzs[l + 1], Ws[l], rows, n_in, n_out, ops.GE_DACT, dzs[l], b_mn=True, act="sigmoid", yprev=acts[l])
    us = timeit(lambda: ops.gemm_group([p]))
    tot += us
    print("dgrad L%d %4dx%4d  %7.1f us  %6.1f TFLOP/s" % (l, n_out, n_in, us, 2.0 * rows * n_in * n_out / us / 1e6))
for split in (2, 3, 4, 6, 8, 12):
    ps = [ops.gemm_problem(dzs[l + 1], acts[l], dims[l + 1], dims[l], rows, ops.GE_ATOMIC, gW[l], a_mn=True, b_mn=True, split_k=split, ones_out=gb[l]) for l in range(4)]
    us = timeit(lambda: ops.gemm_group(ps))
    fl = sum(2.0 * rows * dims[l] * dims[l + 1] for l in range(4))
    print("wgrad group split %2d  %7.1f us  %6.1f TFLOP/s" % (split, us, fl / us / 1e6))
print("fwd+dgrad total us", tot)

# chained groups (layers linked inside one launch by per-row-block counters)
tiles_m = (rows + 255) // 256
dep = torch.zeros((8, tiles_m), dtype=torch.int32, device=DEV)
NODEP = os.environ.get('ABN_NODEP') == '1'
fw = []
for l in range(4):
    n_in, n_out = dims[l], dims[l + 1]
    out = acts[l + 1] if l < 3 else out_last
    fw.append(ops.gemm_problem(acts[l], Ws[l], rows, n_out, n_in, ops.GE_BIAS_ACT, out, act="sigmoid", bias=bias[l],
                               ones_col=(l < 3), signal=dep[l] if (l < 3 and not NODEP) else None, wait=dep[l - 1] if (l > 0 and not NODEP) else None))
def run_fw():
    dep.zero_(); ops.gemm_group(fw)
print("fwd chain (4 layers, one launch)   %7.1f us" % timeit(run_fw))
dg = []
k = 0
for l in range(3, 0, -1):
    n_in, n_out = dims[l], dims[l + 1]
    dg.append(ops.gemm_problem(dzs[l + 1], Ws[l], rows, n_in, n_out, ops.GE_DACT, dzs[l], b_mn=True, act="sigmoid", yprev=acts[l],
                               signal=dep[4 + k] if l > 1 else None, wait=dep[4 + k - 1] if k > 0 else None))
    k += 1
def run_dg():
    dep.zero_(); ops.gemm_group(dg)
print("dgrad chain (3 layers, one launch) %7.1f us" % timeit(run_dg))
