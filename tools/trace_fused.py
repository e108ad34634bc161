"""Device-side timeline of the fused forward kernel (abn_tc3.cu debug stamps)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops, _lib
DEV = "cuda"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
def bf(r, c):
    return (torch.randn(r, ops.pad_row(c + 1), device=DEV) * 0.05).bfloat16()
dims = [280, 500, 500, 500, 100]
acts = [bf(rows, d) for d in dims]
dzs = [bf(rows, d) for d in dims]
Ws = [bf(dims[i + 1], dims[i]) for i in range(4)]
bias = [torch.zeros(dims[i + 1], device=DEV) for i in range(4)]
out_last = torch.zeros(rows, 100, device=DEV)
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
if which == "fwd":
    fl = ops.mlp_layers([(Ws[l], dims[l], bias[l], "sigmoid", acts[l + 1] if l < 3 else out_last, l < 3) for l in range(4)])
    run = lambda: ops.mlp_forward_fused(acts[0], rows, fl)
else:
    dl = ops.mlp_dlayers([(Ws[l], dims[l], "sigmoid", acts[l], dzs[l]) for l in range(3, 0, -1)])
    run = lambda: ops.mlp_dgrad_fused(dzs[4], rows, dl)
dbg = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for _ in range(3): run()
torch.cuda.synchronize()
tr = torch.zeros(148 * 8 * 16 + 2 * 148 * 8 * 2 * 8, dtype=torch.int64, device=DEV)
ctypes.c_void_p.in_dll(_lib.lib(), "abn_gemm_trace_buffer").value = tr.data_ptr()
ctypes.c_int.in_dll(_lib.lib(), "abn_chain_debug").value = dbg
run()
torch.cuda.synchronize()
ctypes.c_int.in_dll(_lib.lib(), "abn_chain_debug").value = 0
ctypes.c_void_p.in_dll(_lib.lib(), "abn_gemm_trace_buffer").value = None
t = tr[:148 * 8 * 16].cpu().view(148, 8, 16)
kbt = tr[148 * 8 * 16:].cpu().view(2, 148, 8, 2, 8)
t0 = int(t[t > 0].min())
names = ["m0wait", "m0go", "m0done", "-", "m1wait", "m1go", "m1done", "-", "e_top", "e_afull", "e_acc0", "e_acc1", "b_math", "b_gate", "b_stored"]
for cta in (0, 64):
    for l in range(8):
        if int(t[cta, l].max()) == 0: continue
        print("cta %3d layer %d: " % (cta, l) + "  ".join("%s %6.2f" % (n, (int(v) - t0) / 1e3) if v > 0 else "%s    -  " % n
                                                            for n, v in zip(names, t[cta, l]) if n != "-"))
print("last event us", (int(t.max()) - t0) / 1e3)

for cta in (0, 64):
    for l in range(4):
        for nt in range(2):
            if int(kbt[1, cta, l, nt].max()) == 0: continue
            req = ["%6.2f" % ((int(v) - t0) / 1e3) if v > 0 else "   -  " for v in kbt[0, cta, l, nt]]
            ful = ["%6.2f" % ((int(v) - t0) / 1e3) if v > 0 else "   -  " for v in kbt[1, cta, l, nt]]
            print("cta %3d layer %d tile %d: B requested %s" % (cta, l, nt, " ".join(req)))
            print("                        seen full   %s" % " ".join(ful))
