import torch, sys
sys.path.insert(0, '/root/repo')
from abnet3_b200 import ops, synth
corpus = synth.make_corpus(160, cluster_size=8, tokens_per_file=80, seed=0)
pairs = synth.make_same_pairs(corpus, 64, seed=1)
feat = corpus.feat.cuda()
try:
    res = ops.align_pairs(feat, pairs.cuda())
    torch.cuda.synchronize()
    print("ok", res.path_len[:8])
except Exception as e:
    print("ERR", e)
