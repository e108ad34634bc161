"""Fused siamese training step: the inner loop of
/root/reference/abnet3/trainer.py:226-256 (``forward -> loss -> zero_grad ->
backward -> optimizer.step``) as a fixed sequence of sm_100a kernels over
pre-allocated buffers, plus the one thing the reference does not have: data
parallelism -- one NCCL all-reduce of a single flat gradient bucket per step.

All parameters of the network are re-pointed at views of ONE flat float32
buffer (and their ``.grad`` at views of one flat gradient buffer), so

* the backward kernels write dW / db straight into the bucket (no autograd
  accumulation kernels, no ``zero_grad``: every slot is overwritten each step;
  parameters the reference never trains -- the multitask ``hidden_layers_spk`` /
  ``hidden_layers_phn`` stacks -- sit outside the trained range);
* the all-reduce is one call on one contiguous tensor (2.77 MB for the
  280-500-500-500-100 network): latency bound, so it is never split;
* the optimizer is one elementwise kernel (abn_optimizer_step) over the bucket.

``state_dict`` / ``.pth`` interchange is unaffected: the module tree and the
parameter shapes are those of the reference.
"""
import torch
import torch.distributed as dist

from . import ops
from .model import SiameseNetwork, SiameseMultitaskNetwork, PRECISIONS

OPTIMIZERS = ("sgd", "adadelta", "adam")


def _trained_layers(network):
    """-> (trunk [(W, b, act)], heads [[(W, b, act)], ...])"""
    if isinstance(network, SiameseNetwork):
        return network.layer_specs(), []
    if isinstance(network, SiameseMultitaskNetwork):
        act = network.activation_layer
        trunk = [(network.input_emb[0].weight, network.input_emb[0].bias, act)]
        for m in network.hidden_layers_shared:
            if isinstance(m, torch.nn.Linear):
                trunk.append((m.weight, m.bias, act))
        heads = [[(network.output_layer_spk[0].weight, network.output_layer_spk[0].bias, act)],
                 [(network.output_layer_phn[0].weight, network.output_layer_phn[0].bias, act)]]
        return trunk, heads
    raise TypeError("unsupported network class %s" % type(network).__name__)


class FlatBucket(object):
    """Re-point the parameters at views of one flat buffer (trained parameters
    first, never-trained ones after) and give them flat gradient views."""

    def __init__(self, network, trained):
        dev = next(network.parameters()).device
        trained_ids = {id(p) for p in trained}
        rest = [p for p in network.parameters() if id(p) not in trained_ids]
        order = list(trained) + rest
        n_trained = sum(p.numel() for p in trained)
        total = sum(p.numel() for p in order)
        self.param = torch.empty(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.n_trained = n_trained
        o = 0
        with torch.no_grad():
            for p in order:
                n = p.numel()
                self.param[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.param[o:o + n].view_as(p.data)
                if id(p) in trained_ids:
                    p.grad = self.grad[o:o + n].view_as(p.data)
                o += n

    @property
    def trained_param(self):
        return self.param[:self.n_trained]

    @property
    def trained_grad(self):
        return self.grad[:self.n_trained]


class SiameseTrainStep(object):
    """One training step of a (multitask) siamese network on device tensors.

    loss_spec: ("coscos2" | "cosmargin", margin, avg) for SiameseNetwork, or
    ((kind_spk, margin, avg), (kind_phn, margin, avg), weight) for the multitask
    network (abnet3/loss.py:165-182).
    """

    def __init__(self, network, loss_spec, optimizer_type="sgd", lr=0.001, momentum=0.9,
                 process_group=None):
        if optimizer_type not in OPTIMIZERS:
            raise ValueError("fused step supports %s, got %r" % (OPTIMIZERS, optimizer_type))
        network._check_supported()
        self.network = network
        self.loss_spec = loss_spec
        self.kind, self.lr, self.momentum = optimizer_type, float(lr), float(momentum or 0.0)
        self.trunk, self.heads = _trained_layers(network)
        trained = []
        for W, b, _ in self.trunk + [l for h in self.heads for l in h]:
            trained += [W, b]
        self.bucket = FlatBucket(network, trained)
        dev = self.bucket.param.device
        n = self.bucket.n_trained
        self.state0 = torch.zeros(n, dtype=torch.float32, device=dev)
        self.state1 = torch.zeros(n, dtype=torch.float32, device=dev) \
            if optimizer_type != "sgd" else None
        self.loss_buf = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_count = 0
        self.precision = PRECISIONS[network.precision]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and \
            dist.is_initialized() else 1
        self._rows = -1
        self.wgrad_split_k = 8
        if self.precision == 1:
            self._alloc_bf16_weights()

    # ---- bf16 tensor-core path: operand copies of the weights -----------------
    def _all_layers(self):
        return self.trunk + [l for h in self.heads for l in h]

    def _alloc_bf16_weights(self):
        dev = self.bucket.param.device
        self.wb, self.wtb = {}, {}
        for W, _, _ in self._all_layers():
            n_out, n_in = W.shape
            self.wb[id(W)] = torch.zeros((n_out, ops.pad8(n_in)), dtype=torch.bfloat16, device=dev)
            self.wtb[id(W)] = torch.zeros((n_in, ops.pad8(n_out)), dtype=torch.bfloat16, device=dev)
        self.refresh_bf16_weights()

    def refresh_bf16_weights(self):
        """bf16 W and W^T operand copies of the fp32 master weights (after every
        optimizer step, or after load_state_dict)."""
        for W, _, _ in self._all_layers():
            ops.cast_bf16(W.data, self.wb[id(W)], self.wtb[id(W)])

    # buffers for a given number of rows (2B): activations and their gradients
    def _reserve(self, rows):
        if rows == self._rows:
            return
        dev = self.bucket.param.device

        def buf(width):
            return torch.empty((rows, width), dtype=torch.float32, device=dev)

        self.acts = [buf(W.shape[0]) for W, _, _ in self.trunk]
        self.dacts = [buf(W.shape[0]) for W, _, _ in self.trunk]
        self.head_acts = [[buf(W.shape[0]) for W, _, _ in h] for h in self.heads]
        self.head_dacts = [[buf(W.shape[0]) for W, _, _ in h] for h in self.heads]
        if self.precision == 1:
            ldm = ops.pad8(rows)

            def b16(r, c):
                return torch.zeros((r, c), dtype=torch.bfloat16, device=dev)

            d_in = self.trunk[0][0].shape[1]
            self.xb, self.xbT = b16(rows, ops.pad8(d_in)), b16(d_in, ldm)
            mk = lambda layers: [(b16(rows, ops.pad8(W.shape[0])), b16(W.shape[0], ldm))
                                 for W, _, _ in layers]
            self.actb = mk(self.trunk)              # (a bf16, a^T bf16) per trunk layer
            self.dzb = mk(self.trunk)               # (dz bf16, dz^T bf16)
            self.head_actb = [mk(h) for h in self.heads]
            self.head_dzb = [mk(h) for h in self.heads]
        self._rows = rows

    def _forward_bf16(self, x):
        rows = x.shape[0]
        ops.cast_bf16(x, self.xb, self.xbT)
        hb = self.xb
        for l, (W, b, act) in enumerate(self.trunk):
            n_out, n_in = W.shape
            ops.gemm_bf16_tn(hb, self.wb[id(W)], rows, n_out, n_in, ops.EPI_BIAS_ACT, b.data, act,
                             out_f32=self.acts[l], out_bf16=self.actb[l][0],
                             outT_bf16=self.actb[l][1])
            hb = self.actb[l][0]
        outs = []
        for hi, head in enumerate(self.heads):
            gb = hb
            for l, (W, b, act) in enumerate(head):
                n_out, n_in = W.shape
                ops.gemm_bf16_tn(gb, self.wb[id(W)], rows, n_out, n_in, ops.EPI_BIAS_ACT, b.data,
                                 act, out_f32=self.head_acts[hi][l],
                                 out_bf16=self.head_actb[hi][l][0],
                                 outT_bf16=self.head_actb[hi][l][1])
                gb = self.head_actb[hi][l][0]
            outs.append(self.head_acts[hi][-1])
        return outs if self.heads else self.acts[-1]

    def _layer_backward_bf16(self, rows, W, b, act, y, dy, dzb, in_bT, dx, dx_accumulate):
        """One layer: dz (bf16 + transposed), db, dgrad into dx (fp32), wgrad into W.grad."""
        n_out, n_in = W.shape
        ops.act_backward_bf16(y, dy, act, dz=dzb[0], dzT=dzb[1], db=b.grad)
        if dx is not None:
            ops.gemm_bf16_tn(dzb[0], self.wtb[id(W)], rows, n_in, n_out,
                             ops.EPI_ATOMIC if dx_accumulate else ops.EPI_STORE, out_f32=dx)
        ops.gemm_bf16_tn(dzb[1], in_bT, n_out, n_in, rows, ops.EPI_ATOMIC, out_f32=W.grad,
                         split_k=self.wgrad_split_k)

    def _backward_bf16(self, x):
        rows = x.shape[0]
        self.bucket.trained_grad.zero_()       # db / dW are accumulated with atomics
        if self.heads:
            first = True
            for hi, head in enumerate(self.heads):
                for l in reversed(range(len(head))):
                    W, b, act = head[l]
                    in_bT = self.head_actb[hi][l - 1][1] if l > 0 else self.actb[-1][1]
                    dx = self.head_dacts[hi][l - 1] if l > 0 else self.dacts[-1]
                    self._layer_backward_bf16(rows, W, b, act, self.head_acts[hi][l],
                                              self.head_dacts[hi][l], self.head_dzb[hi][l], in_bT,
                                              dx, l == 0 and not first)
                first = False
        for l in reversed(range(len(self.trunk))):
            W, b, act = self.trunk[l]
            in_bT = self.actb[l - 1][1] if l > 0 else self.xbT
            self._layer_backward_bf16(rows, W, b, act, self.acts[l], self.dacts[l], self.dzb[l],
                                      in_bT, self.dacts[l - 1] if l > 0 else None, False)

    def forward(self, x):
        """x [rows, input_dim] -> embeddings (last trunk act, or the two heads)."""
        self._reserve(x.shape[0])
        if self.precision == 1:
            return self._forward_bf16(x)
        h = x
        for l, (W, b, act) in enumerate(self.trunk):
            h = ops.linear_forward(h, W.data, b.data, act, self.precision, out=self.acts[l])
        outs = []
        for hi, head in enumerate(self.heads):
            g = h
            for l, (W, b, act) in enumerate(head):
                g = ops.linear_forward(g, W.data, b.data, act, self.precision,
                                       out=self.head_acts[hi][l])
            outs.append(g)
        return outs if self.heads else h

    def _loss_and_seed(self, out, n, labels):
        """loss into self.loss_buf, d(loss)/d(embeddings) into the dact buffers."""
        self.loss_buf.zero_()
        if not self.heads:
            kind, margin, avg = self.loss_spec
            de = self.dacts[-1]
            ops.pair_loss(out[:n], out[n:], labels[0], kind, margin, 1.0 / n if avg else 1.0,
                          loss_out=self.loss_buf, grads=(de[:n], de[n:]))
            return
        spec_spk, spec_phn, weight = self.loss_spec
        for hi, (spec, w, y) in enumerate(((spec_spk, weight, labels[0]),
                                           (spec_phn, 1.0 - weight, labels[1]))):
            kind, margin, avg = spec
            de = self.head_dacts[hi][-1]
            ops.pair_loss(out[hi][:n], out[hi][n:], y, kind, margin,
                          w * (1.0 / n if avg else 1.0), loss_out=self.loss_buf,
                          grads=(de[:n], de[n:]))

    def backward(self, x):
        if self.precision == 1:
            return self._backward_bf16(x)
        trunk = self.trunk
        if self.heads:
            first = True
            for hi, head in enumerate(self.heads):
                for l in reversed(range(len(head))):
                    W, b, act = head[l]
                    xin = self.head_acts[hi][l - 1] if l > 0 else self.acts[-1]
                    dx = self.head_dacts[hi][l - 1] if l > 0 else self.dacts[-1]
                    ops.linear_backward(xin, W.data, self.head_acts[hi][l], self.head_dacts[hi][l],
                                        act, self.precision, dW=W.grad, db=b.grad, accumulate=False,
                                        dx=dx, accumulate_dx=(l == 0 and not first))
                first = False
        for l in reversed(range(len(trunk))):
            W, b, act = trunk[l]
            xin = self.acts[l - 1] if l > 0 else x
            ops.linear_backward(xin, W.data, self.acts[l], self.dacts[l], act, self.precision,
                                need_dx=(l > 0), dW=W.grad, db=b.grad, accumulate=False,
                                dx=self.dacts[l - 1] if l > 0 else None)

    # ---- CUDA graphs ------------------------------------------------------------
    # The step is ~40 small launches on fixed buffers; replaying them as two graphs
    # (forward+loss+backward | optimizer) removes the launch latency that otherwise
    # dominates a 0.1-0.3 ms step.  The NCCL all-reduce stays between the two graphs.
    def input_buffers(self, n):
        """Static (x [2n, D], labels...) buffers of the graphed step: producers
        (abn_gather_batch) may write straight into them."""
        if getattr(self, "_static_n", None) != n:
            dev = self.bucket.param.device
            d_in = self.trunk[0][0].shape[1]
            self._sx = torch.empty((2 * n, d_in), dtype=torch.float32, device=dev)
            self._sy = [torch.empty(n, dtype=torch.float32, device=dev)
                        for _ in range(2 if self.heads else 1)]
            self._static_n = n
            self._graph_fb = self._graph_opt = None
            self._eager_warm = 0
        return (self._sx,) + tuple(self._sy)

    def _fwd_loss_bwd(self):
        out = self.forward(self._sx)
        self._loss_and_seed(out, self._static_n, self._sy)
        self.backward(self._sx)

    def _optimizer(self, scale):
        ops.optimizer_step(self.bucket.trained_param, self.bucket.trained_grad, self.state0,
                           self.state1, self.kind, self.lr, self.momentum, scale, 1)
        if self.precision == 1:
            self.refresh_bf16_weights()

    def step_graphed(self, x, n, *labels):
        """Same as step(do_training=True) through CUDA graphs (SGD / Adadelta)."""
        bufs = self.input_buffers(n)
        for dst, src in zip(bufs, (x,) + tuple(labels)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        avg = self.loss_spec[2] if not self.heads else self.loss_spec[0][2]
        scale = 1.0 / self.world if (self.world > 1 and avg) else 1.0
        if self._graph_fb is None:
            if self._eager_warm < 2:           # warm-up: lazy attribute setup must not be captured
                self._eager_warm += 1
                self._fwd_loss_bwd()
                if self.world > 1:
                    dist.all_reduce(self.bucket.trained_grad, op=dist.ReduceOp.SUM, group=self.group)
                self._optimizer(scale)
                self.step_count += 1
                return self.loss_buf
            torch.cuda.synchronize()
            self._graph_fb, self._graph_opt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_fb):
                self._fwd_loss_bwd()
            with torch.cuda.graph(self._graph_opt):
                self._optimizer(scale)
        self._graph_fb.replay()
        if self.world > 1:
            dist.all_reduce(self.bucket.trained_grad, op=dist.ReduceOp.SUM, group=self.group)
        self._graph_opt.replay()
        self.step_count += 1
        return self.loss_buf

    def step(self, x, n, *labels, do_training=True, graph=False):
        """x = [X1; X2] as one [2n, D] batch; labels float32 [n] (y) or
        (y_spk, y_phn).  Returns the loss as a 1-element device tensor that is
        overwritten by the next step.  graph=True replays CUDA graphs (not for Adam,
        whose bias correction changes every step)."""
        if graph and do_training and self.kind != "adam":
            return self.step_graphed(x, n, *labels)
        out = self.forward(x)
        self._loss_and_seed(out, n, labels)
        if not do_training:
            return self.loss_buf
        self.backward(x)
        grad = self.bucket.trained_grad
        scale = 1.0
        if self.world > 1:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=self.group)
            avg = self.loss_spec[2] if not self.heads else self.loss_spec[0][2]
            scale = 1.0 / self.world if avg else 1.0
        self.step_count += 1
        ops.optimizer_step(self.bucket.trained_param, grad, self.state0, self.state1, self.kind,
                           self.lr, self.momentum, scale, self.step_count)
        if self.precision == 1:
            self.refresh_bf16_weights()
        return self.loss_buf
