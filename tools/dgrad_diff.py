"""Where does the dgrad chain differ from the layer-by-layer GEMMs?  (debug)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops
DEV = "cuda"


def _bf(rows, cols, seed, scale=1.0, pad_val=7.0):
    g = torch.Generator().manual_seed(seed)
    t = torch.full((rows, ops.pad_row(cols + 1)), pad_val, dtype=torch.bfloat16)
    t[:, :cols] = (torch.randn(rows, cols, generator=g) * scale).bfloat16()
    return t.to(DEV)


rows = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
dims = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [280, 500, 100]
act = "sigmoid"
n_layers = len(dims) - 1
g = torch.Generator().manual_seed(5)
Ws = [_bf(dims[l + 1], dims[l], 40 + l, dims[l + 1] ** -0.5) for l in range(n_layers)]
ys = [None] + [_bf(rows, dims[l], 50 + l) for l in range(1, n_layers)]
for l in range(1, n_layers):
    ys[l][:, :dims[l]] = torch.sigmoid(torch.randn(rows, dims[l], generator=g)).bfloat16().to(DEV)
dz_top = _bf(rows, dims[-1], 60, 0.5)
buffers = lambda: [None] + [torch.zeros((rows, ops.pad_row(dims[l])), dtype=torch.bfloat16, device=DEV) for l in range(1, n_layers)]
torch.cuda.synchronize()
for rep in range(8):
    sync_between = rep >= 4
    ref = buffers()
    dz = dz_top
    for l in range(n_layers - 1, 0, -1):
        ops.gemm_group([ops.gemm_problem(dz, Ws[l], rows, dims[l], dims[l + 1], ops.GE_DACT, ref[l], b_mn=True, act=act, yprev=ys[l])])
        dz = ref[l]
    if sync_between:
        torch.cuda.synchronize()
    got = buffers()
    layers = ops.mlp_dlayers([(Ws[l], dims[l], act, ys[l], got[l]) for l in range(n_layers - 1, 0, -1)])
    ops.mlp_dgrad_fused(dz_top, rows, layers)
    torch.cuda.synchronize()
    for l in range(n_layers - 1, 0, -1):
        a, b = got[l][:, :dims[l]].float(), ref[l][:, :dims[l]].float()
        bad = (a != b)
        if not bad.any():
            print("rep %d layer %d: identical" % (rep, l))
            continue
        r, c = bad.nonzero(as_tuple=True)
        blocks = sorted(set((int(x) // 128, int(y) // 64) for x, y in zip(r[:200000:97].tolist(), c[:200000:97].tolist())))
        print("rep %d layer %d: %d elements differ; rows %d..%d cols %d..%d; max |diff| %.3g" % (rep, l, int(bad.sum()), int(r.min()), int(r.max()), int(c.min()), int(c.max()), float((a - b).abs().max())))
        print("   (128-row block, 64-col block) samples:", blocks[:16])
        rb = int(r[0]) // 128
        sub = bad[rb * 128:(rb + 1) * 128]
        print("   first bad row block %d: bad per 64-col block %s; bad rows in block: %s" %
              (rb, [int(sub[:, k * 64:(k + 1) * 64].sum()) for k in range(8)], sub.any(1).nonzero().flatten().tolist()[:12]))
        cols = sorted(set((c % 64).tolist()))
        print("   bad columns mod 64:", cols[:40])
