"""torchrun check of the data-parallel step: NVLink peer-memory fused all-reduce+optimizer
(ABN_DP_P2P=1) against the NCCL all-reduce path, and rank-to-rank equality of the weights."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.model import SiameseNetwork

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)
B = 8192
g = torch.Generator(device=dev).manual_seed(100 + rank)
feat = torch.randn(200000, 280, device=dev, generator=g)
n_fp = 500000
idx1 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32, generator=g)
idx2 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32, generator=g)
y = (torch.randint(0, 2, (n_fp,), device=dev, generator=g) * 2 - 1).to(torch.int8)

def run(p2p, steps=12, graph=True, opt="adadelta", no_comm=False):
    os.environ["ABN_DP_P2P"] = p2p if isinstance(p2p, str) else ("1" if p2p else "0")
    torch.manual_seed(0)
    net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
                         activation_layer="sigmoid", precision="bf16").to(dev)
    if opt == "adadelta":
        step = SiameseTrainStep(net, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
    else:
        step = SiameseTrainStep(net, ("coscos2", 0.0, False), "sgd", lr=1e-4, momentum=0.9)
    mode = {0: "push2", 1: "push", 2: "ll"}[int(step._dp_push.one_shot)] if step._dp_push is not None else (
        "read" if step._dp is not None else "nccl")
    if rank == 0:
        print("   exchange: %s" % mode, flush=True)
    if no_comm:
        step._allreduce = lambda: None
    sel = step.gather_buffers(B)
    losses = []
    for i in range(steps):
        sel.copy_(torch.arange(i * B, (i + 1) * B, device=dev))
        losses.append(float(step.step_gather(feat, idx1, idx2, y, B, graph=graph)))
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(50):
        sel.copy_(torch.arange(i * B, (i + 1) * B, device=dev))
        step.step_gather(feat, idx1, idx2, y, B, graph=graph)
    e1.record()
    host_dt = (time.perf_counter() - t0) / 50
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / 50 * 1e-3
    if rank == 0:
        print("   host enqueue %.1f us/step, device %.1f us/step" % (host_dt * 1e6, dt * 1e6), flush=True)
    w = step.bucket.trained_param.clone()
    return losses, w, dt

l0, w0, dt0 = run(False, no_comm=True)
if rank == 0:
    print("no communication at all: us/step %.1f" % (dt0 * 1e6), flush=True)
for p2p in (("ll", "push", "push2", "0") if world > 2 else ("ll", "push", "push2", "1", "0")):
    losses, w, dt = run(p2p, opt="sgd")
    # every rank must hold the same weights
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    same = all(torch.equal(ws[0], x) for x in ws)
    if rank == 0:
        print("p2p=%s  us/step %.1f  ranks identical: %s  losses %s" % (p2p, dt * 1e6, same, ["%.1f" % l for l in losses[:6]]), flush=True)
    if p2p == "ll": w_p2p, l_p2p = w, losses
    elif p2p == "0": w_nccl, l_nccl = w, losses
rel = float((w_p2p - w_nccl).norm() / w_nccl.norm())
if rank == 0:
    print("weights p2p vs nccl: rel diff %.3e; loss[5] %.2f vs %.2f" % (rel, l_p2p[5], l_nccl[5]), flush=True)
# every rank draws its OWN initialisation (different seeds): the engine broadcasts rank 0's at
# construction, so the replicas must still be identical after training
os.environ.pop("ABN_DP_P2P", None)
torch.manual_seed(1000 + rank)
net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
                     activation_layer="sigmoid", precision="bf16").to(dev)
step = SiameseTrainStep(net, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
tot = step.sweep_table(feat, (idx1, idx2, y), B, 8)
w = step.bucket.param.clone()
ws = [torch.empty_like(w) for _ in range(world)]
dist.all_gather(ws, w)
if rank == 0:
    print("unseeded init, 8 steps: replicas identical: %s" % all(torch.equal(ws[0], x) for x in ws), flush=True)
dist.destroy_process_group()
