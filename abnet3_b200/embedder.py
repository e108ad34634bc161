"""Embedders of /root/reference/abnet3/embedder.py on the sm_100a kernels.

``EmbedderSiamese`` / ``EmbedderSiameseMultitask`` keep the reference's constructor
(`embedder.py:37-47`) and ``embed()`` flow (`:61-100`, `:110-148`): load the network, eval
mode, read the feature file, embed every item, write the embeddings with the items' names and
times.  What differs is how the frames are embedded:

* the reference runs ``network(feat, feat)`` -- BOTH siamese branches on the same rows -- and
  keeps the first output (`embedder.py:91`, `:134`); the branches share their weights, so one
  pass of ``forward_once`` gives the same rows and that is what runs here;
* a bf16 network streams every item through ONE launch per chunk of rows of the forward
  kernel with the activations resident in shared memory (``abn_mlp_forward_fused``; layers
  wider than its 512-feature slab go layer by layer through ``abn_gemm_bf16_group``), bf16
  weight copies made once; an fp32 network uses the fp32 kernels (``abn_linear_forward``);
* ``batch_size`` rows per launch (the reference's knob against GPU memory, default 5000; the
  chunking is ``np.array_split``'s, `embedder.py:84-85`).

Feature containers: the reference reads / writes `h5features` files; that package is used when
importable, otherwise a ``.npz`` archive (``<item>`` arrays + ``times/<item>``) or an in-memory
``{item: array}`` dict is read and a ``.npz`` archive is written (see utils.read_feats).
There is no CPU path: the network and the kernels run on an sm_100 GPU.
"""
import numpy as np
import torch

from . import ops
from .engine import _trained_layers
from .model import PRECISIONS


def _load_features(feature_path):
    """-> (items, times, feats) in file order (abnet3/embedder.py:75-80)."""
    if isinstance(feature_path, dict):
        items = list(feature_path)
        feats = [np.asarray(feature_path[k]) for k in items]
        times = [0.0025 + 0.01 * np.arange(f.shape[0]) for f in feats]
        return items, times, feats
    if str(feature_path).endswith(".npz"):
        z = np.load(feature_path)
        items = [k for k in z.files if not k.startswith("times/")]
        feats = [z[k] for k in items]
        times = [z["times/" + k] if "times/" + k in z.files else 0.0025 + 0.01 * np.arange(f.shape[0])
                 for k, f in zip(items, feats)]
        return items, times, feats
    try:
        import h5features
    except ImportError:
        raise ImportError("reading %r needs the `h5features` package (not installed); pass a .npz "
                          "archive or an {item: array} dict" % (feature_path,))
    with h5features.Reader(feature_path, 'features') as fh:
        features = fh.read()
    return features.items(), features.labels(), features.features()


def _write_features(output_path, items, times, embeddings):
    """abnet3/embedder.py:98-100 (h5features when importable, else a .npz archive)."""
    if output_path is None:
        return
    if not str(output_path).endswith(".npz"):
        try:
            import h5features
            data = h5features.Data(items, times, embeddings, check=True)
            with h5features.Writer(output_path) as fh:
                fh.write(data, 'features')
            return
        except ImportError:
            output_path = str(output_path) + ".npz"
    arrays = {k: e for k, e in zip(items, embeddings)}
    arrays.update({"times/" + k: np.asarray(t) for k, t in zip(items, times)})
    np.savez(output_path, **arrays)


class _ForwardChain(object):
    """Inference-only bf16 chain of a (multitask) siamese network: bf16 weight copies made once,
    chunk buffers reused, one abn_mlp_forward_fused launch per chunk."""

    def __init__(self, network, max_rows):
        trunk, heads = network.inference_layers()         # (eval-mode BatchNorm folded in)
        layers = [(W.data, b.data, act) for W, b, act in trunk]
        self.head_dim = 0
        if heads:       # the two heads side by side: one [2d, hidden] layer
            Ws = [h[0][0].data for h in heads]
            bs = [h[0][1].data for h in heads]
            layers.append((torch.cat(Ws, 0).contiguous(), torch.cat(bs, 0).contiguous(), heads[0][0][2]))
            self.head_dim = Ws[0].shape[0]
        dev = layers[0][0].device
        self.rows = max_rows
        self.d_in = layers[0][0].shape[1]
        self.xb = torch.zeros((max_rows, ops.pad_row(self.d_in + 1)), dtype=torch.bfloat16, device=dev)
        self.wb, self.bias, self.act, self.outs = [], [], [], []
        for l, (W, b, act) in enumerate(layers):
            n_out, n_in = W.shape
            wb = torch.zeros((n_out, ops.pad_row(n_in)), dtype=torch.bfloat16, device=dev)
            ops.cast_bf16(W.contiguous(), wb, None)
            self.wb.append(wb)
            self.bias.append(b.contiguous())
            self.act.append(act)
            last = l == len(layers) - 1
            self.outs.append(torch.empty((max_rows, n_out), dtype=torch.float32, device=dev) if last else
                             torch.zeros((max_rows, ops.pad_row(n_out + 1)), dtype=torch.bfloat16, device=dev))
        self.dims = [(W.shape[1], W.shape[0]) for W, _, _ in layers]
        fits = all(n_in <= ops.MLP_MAX_WIDTH and n_out + 1 <= ops.MLP_MAX_WIDTH for n_in, n_out in self.dims)
        self.fused = None
        if fits and len(layers) <= ops.MLP_MAX_LAYERS:
            self.fused = ops.mlp_layers([(self.wb[l], self.dims[l][0], self.bias[l], self.act[l], self.outs[l],
                                          l < len(layers) - 1) for l in range(len(layers))])

    def __call__(self, x):
        """x: fp32 CUDA [n <= max_rows, d_in] -> fp32 [n, n_last] (a view of the chunk buffer)."""
        n = x.shape[0]
        ops.cast_bf16(x, self.xb[:n], None)
        if self.fused is not None:
            ops.mlp_forward_fused(self.xb, n, self.fused)
        else:
            h = self.xb
            for l, (n_in, n_out) in enumerate(self.dims):
                ops.gemm_group([ops.gemm_problem(h, self.wb[l], n, n_out, n_in, ops.GE_BIAS_ACT, self.outs[l],
                                                 act=self.act[l], bias=self.bias[l],
                                                 ones_col=(l < len(self.dims) - 1))])
                h = self.outs[l]
        return self.outs[-1][:n]


class EmbedderBuilder:
    """abnet3/embedder.py:19-51 (same parameters)."""

    def __init__(self, network=None, network_path=None, feature_path=None,
                 output_path=None, cuda=True, batch_size=5000):
        if network is None:
            raise ValueError("network is None.")
        self.network = network
        self.network_path = network_path
        self.feature_path = feature_path
        self.output_path = output_path
        self.cuda = cuda
        self.batch_size = batch_size

    def embed(self):
        raise NotImplementedError('Unimplemented embed for class:',
                                  self.__class__.__name__)

    # -- shared machinery ---------------------------------------------------
    def _prepare(self):
        if self.network_path is not None:
            self.network.load_network(self.network_path)
        self.network.eval()
        if not self.cuda:
            raise RuntimeError("abnet3_b200 embedders run on an sm_100 GPU only (there is no CPU path)")
        self.network.cuda()
        self.network._check_supported()

    def _embed_items(self, feats):
        """[array [n_i, d]] -> [fp32 array [n_i, n_out_total]], chunked like the reference
        (n_batches = len // batch_size + 1, np.array_split; abnet3/embedder.py:84-85)."""
        dev = next(self.network.parameters()).device
        bf16 = PRECISIONS[self.network.precision] == 1
        longest = max((-(-len(f) // (len(f) // self.batch_size + 1)) for f in feats if len(f)), default=1)
        chain = _ForwardChain(self.network, longest) if bf16 else None
        out = []
        with torch.no_grad():
            for feat in feats:
                if feat.dtype != np.float32:
                    feat = feat.astype(np.float32)
                if len(feat) == 0:
                    out.append(np.zeros((0, 0), dtype=np.float32))
                    continue
                n_batches = len(feat) // self.batch_size + 1
                pieces = []
                for b_feat in np.array_split(feat, n_batches):
                    if len(b_feat) == 0:
                        continue
                    x = torch.from_numpy(np.ascontiguousarray(b_feat)).to(dev, non_blocking=True)
                    if chain is not None:
                        emb = chain(x)
                    else:
                        emb = self.network.forward_once(x)
                        if isinstance(emb, tuple):
                            emb = torch.cat(emb, 1)
                    pieces.append(emb.cpu().numpy().copy())
                out.append(np.vstack(pieces))
        return out


class EmbedderSiamese(EmbedderBuilder):
    """abnet3/embedder.py:54-100: one embedding per frame, written to ``output_path``."""

    def embed(self):
        self._prepare()
        items, times, feats = _load_features(self.feature_path)
        embeddings = self._embed_items(feats)
        _write_features(self.output_path, items, times, embeddings)
        return dict(zip(items, embeddings))


class EmbedderSiameseMultitask(EmbedderBuilder):
    """abnet3/embedder.py:103-148: speaker and phone embeddings, written to
    ``output_path + '.spk'`` and ``output_path + '.phn'``."""

    def embed(self):
        self._prepare()
        items, times, feats = _load_features(self.feature_path)
        both = self._embed_items(feats)
        d = self.network.output_layer_spk[0].weight.shape[0]
        spk = [e[:, :d] for e in both]
        phn = [e[:, d:] for e in both]
        if self.output_path is not None:
            _write_features(str(self.output_path) + '.spk', items, times, spk)
            _write_features(str(self.output_path) + '.phn', items, times, phn)
        return dict(zip(items, spk)), dict(zip(items, phn))
