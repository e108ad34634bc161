"""Time the tcgen05 GEMM on the embedder's shapes with different epilogue outputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops
dev = "cuda"
def bf(r, c): return (torch.randn(r, ops.pad8(c), device=dev) * 0.05).bfloat16()
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
M = 16384
for (N, K) in [(500, 500), (500, 280), (100, 500)]:
    A, B = bf(M, K), bf(N, K)
    bias = torch.zeros(N, device=dev)
    o32 = torch.empty((M, N), device=dev); o16 = torch.zeros((M, ops.pad8(N)), dtype=torch.bfloat16, device=dev)
    oT = torch.zeros((N, ops.pad8(M)), dtype=torch.bfloat16, device=dev)
    y = bf(M, N).abs().clamp(0, 1); db = torch.zeros(N, device=dev)
    gf = 2.0 * M * N * K / 1e9
    cases = {
        "store, no outputs": lambda: ops.gemm_bf16_tn(A, B, M, N, K, ops.EPI_STORE),
        "store f32": lambda: ops.gemm_bf16_tn(A, B, M, N, K, ops.EPI_STORE, out_f32=o32),
        "bias+sigmoid, no outputs": lambda: ops.gemm_bf16_tn(A, B, M, N, K, ops.EPI_BIAS_ACT, bias, "sigmoid"),
        "bias+sigmoid bf16": lambda: ops.gemm_bf16_tn(A, B, M, N, K, ops.EPI_BIAS_ACT, bias, "sigmoid", out_bf16=o16),
        "bias+sigmoid bf16+T": lambda: ops.gemm_bf16_tn(A, B, M, N, K, ops.EPI_BIAS_ACT, bias, "sigmoid", out_bf16=o16, outT_bf16=oT),
        "dgrad_act bf16+T+db": lambda: ops.gemm_bf16_tn(A, B, M, N, K, ops.EPI_DGRAD_ACT, act="sigmoid", yprev=y, out_bf16=o16, outT_bf16=oT, db=db),
    }
    for nm, fn in cases.items():
        us = timeit(fn)
        print("M=%d N=%d K=%d  %-28s %7.1f us  %6.1f TFLOP/s" % (M, N, K, nm, us, gf / us * 1e-3 * 1e3 / 1e3 * 1e3 / 1e3))
# wgrad shape
A, B = bf(500, M), bf(500, M)
out = torch.zeros((500, 500), device=dev)
for sk in (1, 4, 9, 18, 36):
    us = timeit(lambda: ops.gemm_bf16_tn(A, B, 500, 500, M, ops.EPI_ATOMIC, out_f32=out, split_k=sk))
    print("wgrad 500x500xK=%d split_k=%d  %7.1f us  %6.1f TFLOP/s" % (M, sk, us, 2.0 * 500 * 500 * M / us / 1e6))
