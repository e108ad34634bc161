"""One eager launch of each grouped-GEMM form on the training-step shapes (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops
DEV = "cuda"
rows = 16384
def bf(r, c):
    return (torch.randn(r, ops.pad8(c + 1), device=DEV) * 0.05).bfloat16()
x, a1, dz1, dz2 = bf(rows, 500), bf(rows, 500), bf(rows, 500), bf(rows, 500)
W = bf(500, 500)
bias = torch.zeros(500, device=DEV)
gW = torch.zeros(500, 500, device=DEV); gb = torch.zeros(500, device=DEV)
fwd = ops.gemm_problem(x, W, rows, 500, 500, ops.GE_BIAS_ACT, a1, act="sigmoid", bias=bias, ones_col=True)
dg = ops.gemm_problem(dz2, W, rows, 500, 500, ops.GE_DACT, dz1, b_mn=True, act="sigmoid", yprev=a1)
wg = ops.gemm_problem(dz2, x, 500, 500, rows, ops.GE_ATOMIC, gW, a_mn=True, b_mn=True, split_k=18, ones_out=gb)
for _ in range(3):
    ops.gemm_group([fwd]); ops.gemm_group([dg]); ops.gemm_group([wg])
torch.cuda.synchronize()
print("done")
