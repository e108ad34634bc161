"""Diagnostic (GPU box): accuracy of error-compensated TF32 tensor-core dot
products (hi/lo split, 3 products) against fp64, next to plain fp32, on
token-like data.  Decides whether the distance stage may run on tcgen05."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from abnet3_b200 import synth

torch.manual_seed(0)
c = synth.make_corpus(4000, seed=5, device="cuda")
X = c.feat[:8192].contiguous()
Y = (0.9 * X + 0.45 * c.feat[8192:16384]).contiguous()      # cos ~ 0.9 on the diagonal
ex = (X.double() @ Y.double().T)
scale = X.double().norm(dim=1)[:, None] * Y.double().norm(dim=1)[None, :]

def rep(name, G):
    e = (G.double() - ex) / scale
    d = torch.diagonal(e)
    print("%-34s all: mean %+.2e rms %.2e max %.2e | diag(cos~.9): mean %+.2e rms %.2e max %.2e" % (
        name, e.mean(), e.pow(2).mean().sqrt(), e.abs().max(), d.mean(), d.pow(2).mean().sqrt(), d.abs().max()))

torch.backends.cuda.matmul.allow_tf32 = False
rep("fp32 cuBLAS (allow_tf32=False)", X @ Y.T)
torch.backends.cuda.matmul.allow_tf32 = True
rep("tf32 x1", X @ Y.T)

def split_trunc(A):
    hi = (A.view(torch.int32) & ~0x1FFF).view(torch.float32)
    return hi, A - hi
def split_round(A):
    i = A.view(torch.int32)
    hi = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    return hi, A - hi

for nm, sp in (("trunc", split_trunc), ("round", split_round)):
    xh, xl = sp(X); yh, yl = sp(Y)
    G = xh @ yh.T + (xh @ yl.T + xl @ yh.T)
    rep("tf32 x3 (%s split)" % nm, G)
    G4 = G + xl @ yl.T
    rep("tf32 x4 (%s split)" % nm, G4)
    # K split in 7 chunks of 40, partial GEMMs summed in fp32 (what a chunked kernel does)
    Gc = torch.zeros_like(G)
    for k in range(0, 280, 40):
        s = slice(k, k + 40)
        Gc += xh[:, s] @ yh[:, s].T + (xh[:, s] @ yl[:, s].T + xl[:, s] @ yh[:, s].T)
    rep("tf32 x3 (%s) 7 K-chunks fp32-summed" % nm, Gc)
# bf16 x 3-way split (6 products)
torch.backends.cuda.matmul.allow_tf32 = False
def bsplit(A):
    a0 = A.bfloat16(); r = A - a0.float(); a1 = r.bfloat16(); a2 = (r - a1.float()).bfloat16()
    return a0, a1, a2
x0, x1, x2 = bsplit(X); y0, y1, y2 = bsplit(Y)
mm = lambda a, b: (a @ b.T).float()
G = mm(x0, y0) + (mm(x0, y1) + mm(x1, y0)) + (mm(x0, y2) + mm(x2, y0) + mm(x1, y1))
rep("bf16 x6 (outputs rounded to bf16!)", G)
