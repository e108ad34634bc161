"""Phases of the C3 end-to-end step of bench.py (host table + pair lists -> epoch), wall clock with syncs."""
import os, sys, time, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import synth, utils
from abnet3_b200.dataloader import FramesDataLoader
from abnet3_b200.trainer import TrainerSiamese
from abnet3_b200.model import SiameseNetwork
from abnet3_b200.loss import coscos2
dev = torch.device("cuda", 0)
cap = int(sys.argv[1]) if len(sys.argv) > 1 else 300
corpus = synth.make_corpus(40000, seed=0, device=dev)
P = 1_000_000
toks = [synth.make_same_pairs(corpus, P, seed=1), synth.make_diff_pairs(corpus, P, seed=100),
        synth.make_same_pairs(corpus, P // 100, seed=500), synth.make_diff_pairs(corpus, P // 100, seed=600)]
host_feat = torch.empty(corpus.feat.shape, dtype=torch.float32, pin_memory=True); host_feat.copy_(corpus.feat)
flat = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t) for t in toks]
file_off = corpus.file_off.tolist()
torch.manual_seed(0)
net = SiameseNetwork(input_dim=280, num_hidden_layers=2, hidden_dim=500, output_dim=100, p_dropout=0.0,
                     activation_layer="sigmoid").to(dev)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
trainer = None
for it in range(3):
    t0 = T()
    tab = utils.FeatureTable.from_host(host_feat, file_off, device=dev)
    t1 = T()
    dt = [h.to(dev, non_blocking=True) for h in flat]
    ld = FramesDataLoader.from_tokens(tab, {"train": (dt[0], dt[1]), "dev": (dt[2], dt[3])}, batch_size=8192,
                                      exact_numpy_shuffle=False)
    o2 = ld.epoch_table
    ld.epoch_table = lambda train_mode=True: (lambda r: r[:4] + (min(r[4], cap),))(o2(train_mode))
    with contextlib.redirect_stdout(sys.stderr):
        if trainer is None:
            trainer = TrainerSiamese(network=net, loss=coscos2(avg=False), optimizer_type="adadelta", lr=0.1,
                                     momentum=None, cuda=True, dataloader=ld, log_dir="/tmp/abn_runs")
        trainer.dataloader = ld
        t2 = T()
        ld.load_data()
        t3 = T()
        trainer.optimize_model(do_training=True)
        t4 = T()
    print("iter %d: upload+wrap %.1f ms | loader %.1f | load_data (align train+dev, shuffle) %.1f | epoch(%d batches) %.1f"
          % (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), cap, 1e3 * (t4 - t3)), flush=True)
