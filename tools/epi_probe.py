"""Epilogue cost in isolation: forward-form GEMM with K = 64 (one k-block per tile)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops
DEV = "cuda"
rows = 16384
def bf(r, c):
    return (torch.randn(r, ops.pad8(c + 1), device=DEV) * 0.05).bfloat16()
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
for K in (64, 512):
    x, W = bf(rows, K), bf(500, K)
    bias = torch.zeros(500, device=DEV)
    o16 = bf(rows, 500); o32 = torch.zeros(rows, 500, device=DEV); y = bf(rows, 500); W2 = bf(K, 500)
    for name, p in (
        ("fwd sigmoid bf16", ops.gemm_problem(x, W, rows, 500, K, ops.GE_BIAS_ACT, o16, act="sigmoid", bias=bias, ones_col=True)),
        ("fwd none    bf16", ops.gemm_problem(x, W, rows, 500, K, ops.GE_BIAS_ACT, o16, act="none", bias=bias)),
        ("fwd relu    bf16", ops.gemm_problem(x, W, rows, 500, K, ops.GE_BIAS_ACT, o16, act="relu", bias=bias)),
        ("fwd none    f32 ", ops.gemm_problem(x, W, rows, 500, K, ops.GE_BIAS_ACT, o32, act="none", bias=bias)),
        ("dgrad sigm  bf16", ops.gemm_problem(x, W2, rows, 500, K, ops.GE_DACT, o16, b_mn=True, act="sigmoid", yprev=y)),
        ("N=8 tiny epi    ", ops.gemm_problem(x, W, rows, 8, K, ops.GE_BIAS_ACT, o16, act="none", bias=bias)),
    ):
        print("K=%3d %s %7.1f us" % (K, name, timeit(lambda: ops.gemm_group([p]))))
