import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops, _lib
DEV = "cuda"
rows = 16384
PAD = int(os.environ.get("PADTO", "8"))
def bf(r, c):
    return (torch.randn(r, (c + 1 + PAD - 1) // PAD * PAD, device=DEV) * 0.05).bfloat16()
dims = [280, 500, 500, 500, 100]
acts = [bf(rows, d) for d in dims]
dzs = [bf(rows, d) for d in dims]
Ws = [bf(dims[i + 1], dims[i]) for i in range(4)]
bias = torch.zeros(500, device=DEV)
gW = [torch.zeros(dims[i + 1], dims[i], device=DEV) for i in range(4)]
gb = [torch.zeros(dims[i + 1], device=DEV) for i in range(4)]
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
split = int(sys.argv[2]) if len(sys.argv) > 2 else 12
if which == "fwd":
    probs = [ops.gemm_problem(acts[1], Ws[1], rows, 500, 500, ops.GE_BIAS_ACT, acts[2], act="sigmoid", bias=bias, ones_col=True)]
elif which == "dgrad":
    probs = [ops.gemm_problem(dzs[3], Ws[2], rows, 500, 500, ops.GE_DACT, dzs[2], b_mn=True, act="sigmoid", yprev=acts[2])]
elif which in ("chain", "chain_nodep"):
    tiles_m = (rows + 255) // 256
    dep = torch.zeros((8, tiles_m), dtype=torch.int32, device=DEV)
    out_last = torch.zeros(rows, 100, device=DEV)
    biases = [torch.zeros(dims[l + 1], device=DEV) for l in range(4)]
    probs = []
    for l in range(4):
        out = acts[l + 1] if l < 3 else out_last
        probs.append(ops.gemm_problem(acts[l], Ws[l], rows, dims[l + 1], dims[l], ops.GE_BIAS_ACT, out, act="sigmoid", bias=biases[l],
                                      ones_col=(l < 3), signal=dep[l] if (l < 3 and which == 'chain') else None, wait=dep[l - 1] if (l > 0 and which == 'chain') else None))
else:
    probs = [ops.gemm_problem(dzs[l + 1], acts[l], dims[l + 1], dims[l], rows, ops.GE_ATOMIC, gW[l], a_mn=True, b_mn=True, split_k=split, ones_out=gb[l]) for l in range(4)]
for _ in range(3):
    if which == 'chain': dep.zero_()
    ops.gemm_group(probs)
torch.cuda.synchronize()
tr = torch.zeros(148 * 8 * 16, dtype=torch.int64, device=DEV)
ctypes.c_void_p.in_dll(_lib.lib(), "abn_gemm_trace_buffer").value = tr.data_ptr()
if which == 'chain': dep.zero_()
ops.gemm_group(probs)
torch.cuda.synchronize()
ctypes.c_void_p.in_dll(_lib.lib(), "abn_gemm_trace_buffer").value = None
t = tr.cpu().view(148, 8, 16)
t0 = int(t[t > 0].min())
names = ["p0", "p1", "m0", "mfree", "mcommit", "e0", "etfull", "edone", "c_ld0", "c_ld1", "c_math", "c_sts", "c_bar", "c_tma", "-", "-"]
for cta in (0, 2, 64, 100, 126):
    for it in range(8):
        if int(t[cta, it].max()) == 0: continue
        print("cta %3d tile %d: " % (cta, it) + "  ".join("%s %6.2f" % (n, (int(v) - t0) / 1e3) if v > 0 else "%s    -  " % n for n, v in list(zip(names, t[cta, it]))[:8]))
print("last event us", (int(t.max()) - t0) / 1e3)
