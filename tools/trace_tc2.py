import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops, _lib
DEV = "cuda"
rows = 16384
def bf(r, c):
    return (torch.randn(r, ops.pad8(c + 1), device=DEV) * 0.05).bfloat16()
x, a1 = bf(rows, 500), bf(rows, 500)
W = bf(500, 500)
bias = torch.zeros(500, device=DEV)
fwd = ops.gemm_problem(x, W, rows, 500, 500, ops.GE_BIAS_ACT, a1, act="sigmoid", bias=bias, ones_col=True)
for _ in range(3): ops.gemm_group([fwd])
torch.cuda.synchronize()
tr = torch.zeros(148 * 4 * 8, dtype=torch.int64, device=DEV)
ctypes.c_void_p.in_dll(_lib.lib(), "abn_gemm_trace_buffer").value = tr.data_ptr()
ops.gemm_group([fwd])
torch.cuda.synchronize()
ctypes.c_void_p.in_dll(_lib.lib(), "abn_gemm_trace_buffer").value = None
t = tr.cpu().view(148, 4, 8)
t0 = int(t[t > 0].min())
names = ["prod_start", "prod_done", "mma_start", "mma_acc_free", "mma_commit", "epi_start", "epi_tfull", "epi_done"]
for cta in (0, 1, 73, 147):
    for it in range(2):
        print("cta %3d tile %d: " % (cta, it) + "  ".join("%s %6.2f" % (n, (int(v) - t0) / 1e3) if v > 0 else "%s    -  " % n for n, v in zip(names, t[cta, it])))
print("last event us", (int(t.max()) - t0) / 1e3)
