"""Device-side timeline of the gradient-exchange kernel (torchrun, N ranks): phase durations of
abn_dp_push_step averaged over the last steps, and the step time with / without the exchange."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from abnet3_b200 import _lib
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.model import SiameseNetwork
rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)
B = 8192
g = torch.Generator(device=dev).manual_seed(100 + rank)
feat = torch.randn(500000, 280, device=dev, generator=g)
n_fp = 4_000_000
idx1 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32, generator=g)
idx2 = torch.randint(0, feat.shape[0], (n_fp,), device=dev, dtype=torch.int32, generator=g)
y = (torch.randint(0, 2, (n_fp,), device=dev, generator=g) * 2 - 1).to(torch.int8)
table = (idx1, idx2, y)
torch.manual_seed(0)
net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
                     activation_layer="sigmoid").to(dev)
eng = SiameseTrainStep(net, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
trace = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
_lib.check(_lib.lib().abn_dp_set_trace(trace.data_ptr()))
eng.sweep_table(feat, table, B, 40)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
eng.sweep_table(feat, table, B, 400)
b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / 400 * 1e3
t = trace.view(64, 8).cpu().double()
mode = {0: "two-shot", 1: "one-shot", 2: "ll"}[int(eng._dp_push.one_shot)] if eng._dp_push is not None else "nccl"
d = (t[:, 1:] - t[:, :-1]) / 1e3
names = ["push", "fence+ticket+raise", "wait pushed", "reduce+update(+push params)", "fence+ticket+raise", "wait updated", "bf16+zero"]
if mode == "one-shot":
    names = ["push all", "fence+ticket+raise", "wait pushed"]
    d = d[:, :3]
if mode == "ll":
    names = ["push (LL)", "spin+reduce+update+push params", "spin+store others' slices"]
    d = d[:, :3]
ok = (t[:, 0] > 0)
print("\nrank %d %s: step %.1f us | exchange kernel phases (us, mean over %d steps): %s | total %.1f" % (
    rank, mode, us, int(ok.sum()), " | ".join("%s %.1f" % (n, float(d[ok, k].mean())) for k, n in enumerate(names)),
    float((t[ok, len(names)] - t[ok, 0]).mean() / 1e3)), flush=True)
dist.destroy_process_group()
