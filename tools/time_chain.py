"""Experiment: forward chain / single layer timing of the grouped GEMM under the debug knobs of
the experiment build (ABN_LIB=.../libabnet3_b200_dbg.so; ABN_GEMM_DBG bit mask, abn_tc2.cu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops
DEV = "cuda"
rows = 16384

PAD = int(os.environ.get("PADTO", "8"))
def bf(r, c):
    return (torch.randn(r, (c + 1 + PAD - 1) // PAD * PAD, device=DEV) * 0.05).bfloat16()

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
Ws = [bf(dims[i + 1], dims[i]) for i in range(4)]
bias = [torch.zeros(dims[i + 1], device=DEV) for i in range(4)]
out_last = torch.zeros(rows, 100, device=DEV)
tiles_m = (rows + 255) // 256
dep = torch.zeros((12, tiles_m), dtype=torch.int32, device=DEV)

def chain(nodep):
    fw = []
    for l in range(4):
        out = acts[l + 1] if l < 3 else out_last
        fw.append(ops.gemm_problem(acts[l], Ws[l], rows, dims[l + 1], dims[l], ops.GE_BIAS_ACT, out, act="sigmoid",
                                   bias=bias[l], ones_col=(l < 3), signal=dep[l] if (l < 3 and not nodep) else None,
                                   wait=dep[l - 1] if (l > 0 and not nodep) else None))
    return fw

fw_dep, fw_nodep = chain(False), chain(True)
# 4 x the same 500x500 layer, no dependencies: 512 equal tiles
same = [ops.gemm_problem(acts[1], Ws[1], rows, 500, 500, ops.GE_BIAS_ACT, acts[2], act="sigmoid", bias=bias[1], ones_col=True)
        for _ in range(4)]
one = same[:1]
def run_dep():
    dep.zero_(); ops.gemm_group(fw_dep)
dims_d = dims
dzs = [bf(rows, d) for d in dims]
def dchain(nodep):
    dg, k = [], 0
    for l in range(3, 0, -1):
        n_in, n_out = dims[l], dims[l + 1]
        dg.append(ops.gemm_problem(dzs[l + 1], Ws[l], rows, n_in, n_out, ops.GE_DACT, dzs[l], b_mn=True, act="sigmoid",
                                   yprev=acts[l], signal=dep[4 + k] if (l > 1 and not nodep) else None,
                                   wait=dep[4 + k - 1] if (k > 0 and not nodep) else None))
        k += 1
    return dg
dg_dep, dg_nodep = dchain(False), dchain(True)
def run_dg():
    dep.zero_(); ops.gemm_group(dg_dep)
gW = [torch.zeros(dims[i + 1], dims[i], device=DEV) for i in range(4)]
gb = [torch.zeros(dims[i + 1], device=DEV) for i in range(4)]
wg = [ops.gemm_problem(dzs[l + 1], acts[l], dims[l + 1], dims[l], rows, ops.GE_ATOMIC, gW[l], a_mn=True, b_mn=True,
                       split_k=int(os.environ.get("WSPLIT", "9")), ones_out=gb[l]) for l in range(4)]
def merged(split):
    """dgrad(l), wgrad(l), dgrad(l-1), ... in ONE launch; wgrad waits on the dz its layer needs."""
    out, k, ready = [], 0, {3: None}
    for l in range(3, 0, -1):
        out.append(ops.gemm_problem(dzs[l + 1], Ws[l], rows, dims[l], dims[l + 1], ops.GE_DACT, dzs[l], b_mn=True,
                                    act="sigmoid", yprev=acts[l], signal=dep[4 + k], wait=ready[l]))
        out.append(ops.gemm_problem(dzs[l + 1], acts[l], dims[l + 1], dims[l], rows, ops.GE_ATOMIC, gW[l], a_mn=True,
                                    b_mn=True, split_k=split, ones_out=gb[l], wait=ready[l]))
        ready[l - 1] = dep[4 + k]
        k += 1
    out.append(ops.gemm_problem(dzs[1], acts[0], dims[1], dims[0], rows, ops.GE_ATOMIC, gW[0], a_mn=True,
                                b_mn=True, split_k=split, ones_out=gb[0], wait=ready[0]))
    return out
mg = merged(int(os.environ.get("WSPLIT", "9")))
def run_mg():
    dep.zero_(); ops.gemm_group(mg)
# each argument: comma-separated VAR=value settings applied for that measurement ("-" = none)
for cfg in (sys.argv[1:] or ["-"]):
    sets = [kv.split("=") for kv in cfg.split(",") if "=" in kv]
    for k_, v_ in sets: os.environ[k_] = v_
    dbg = int(os.environ.get("ABN_GEMM_DBG", "0"))
    t_dep = timeit(run_dep) if dbg == 0 else float("nan")
    t_nodep = timeit(lambda: ops.gemm_group(fw_nodep))
    t_same = timeit(lambda: ops.gemm_group(same))
    t_one = timeit(lambda: ops.gemm_group(one))
    t_dg = timeit(run_dg) if dbg == 0 else float("nan")
    t_dgn = timeit(lambda: ops.gemm_group(dg_nodep)) if dbg == 0 else float("nan")
    t_wg = timeit(lambda: ops.gemm_group(wg)) if dbg == 0 else float("nan")
    t_mg = timeit(run_mg) if dbg == 0 else float("nan")
    print("%-40s fwd chain %5.1f nodep %5.1f | 4x500 %5.1f 1x500 %5.1f | dgrad chain %5.1f nodep %5.1f | wgrad %5.1f | bwd merged %5.1f us" %
          (cfg, t_dep, t_nodep, t_same, t_one, t_dg, t_dgn, t_wg, t_mg), flush=True)
    for k_, v_ in sets: os.environ.pop(k_)
