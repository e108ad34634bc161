"""Round-2 ncu target: one invocation of every hot kernel at bench-like sizes, few launches.

  alignment   100 k same pairs of ONE size class (tokens of 50-58 frames -> class (4,4)), stacked fast
              path and generic kernels: align_stack_kernel<4,4> / align_class_kernel<4,4> + dtw_skew_kernel<2>
  long tokens 2 000 pairs of 400-frame tokens: long_tile_kernel<true,6> + dtw_band_kernel
  training    3 eager steps of the C3 step (8192 frame pairs): gather, forward chain, loss, dgrad chain,
              wgrad group, optimizer

    ncu --set full --clock-control none --import-source on -k regex:... -o gpurun_out/r2_full python tools/prof_r2.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import ops, synth
from abnet3_b200.engine import SiameseTrainStep
from abnet3_b200.model import SiameseNetwork

dev = "cuda"
c = synth.make_corpus(20000, seed=0, device=dev, len_range=(50, 58))
pairs = synth.make_same_pairs(c, 100_000, seed=1)
for stack in (7, 0):
    res = ops.align_pairs(c.feat, pairs, stack=stack)
torch.cuda.synchronize()
print("short pairs", pairs.shape[0], "mean path", float(res.path_len.float().mean()))
d1, d2, _ = ops.compact_paths(res)

cl = synth.make_corpus(600, seed=3, device=dev, len_range=(400, 400), tokens_per_file=100)
pl = synth.make_same_pairs(cl, 2000, seed=4)
rl = ops.align_pairs(cl.feat, pl, stack=7)
torch.cuda.synchronize()
print("long pairs", pl.shape[0], "valid", int(rl.valid.sum()))

torch.manual_seed(0)
net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
                     activation_layer="sigmoid").to(dev)
eng = SiameseTrainStep(net, ("coscos2", 0.0, False), "adadelta", lr=0.1, momentum=None)
y = torch.ones(d1.numel(), dtype=torch.int8, device=dev)
y[1::2] = -1
perm = torch.randperm(d1.numel(), device=dev)
table = (d1[perm].contiguous(), d2[perm].contiguous(), y)
os.environ["ABN_PIPELINE"] = "0"
tot = eng.sweep_table(c.feat, table, 8192, 3, graph=False)
torch.cuda.synchronize()
print("train loss", float(tot) / 3)
