"""Fused siamese training step: the inner loop of
/root/reference/abnet3/trainer.py:226-256 (``forward -> loss -> zero_grad ->
backward -> optimizer.step``) as a fixed sequence of sm_100a kernels over
pre-allocated buffers, plus the one thing the reference does not have: data
parallelism -- one NCCL all-reduce of a single flat gradient bucket per step.

All parameters of the network are re-pointed at views of ONE flat float32
buffer (and their ``.grad`` at views of one flat gradient buffer), so

* the backward kernels write dW / db straight into the bucket (no autograd
  accumulation kernels; parameters the reference never trains -- the multitask
  ``hidden_layers_spk`` / ``hidden_layers_phn`` stacks -- sit outside the trained
  range);
* the all-reduce is one call on one contiguous tensor (2.77 MB for the
  280-500-500-500-100 network): latency bound, so it is never split;
* the optimizer is one elementwise kernel (abn_optimizer_step) over the bucket.

Two kernel paths, selected by ``network.precision``:

``fp32``  SIMT GEMMs (abn_linear_forward / backward): the 1e-4 parity path.
``bf16``  tcgen05 tensor cores (abn_mlp_forward_fused / abn_mlp_dgrad_fused /
          abn_gemm_bf16_group): activations and dz live in bf16 in their natural
          row-major layout, the dgrad epilogue applies act'(y) of the layer below
          and emits that layer's dz directly, the two multitask heads run as one
          200-wide layer; master weights, loss, embeddings and gradients stay fp32.
          This is the default on CUDA (``network.precision``).

``state_dict`` / ``.pth`` interchange is unaffected: the module tree and the
parameter shapes are those of the reference.
"""
import os

import torch
import torch.distributed as dist

from . import ops
from .model import SiameseNetwork, SiameseMultitaskNetwork, PRECISIONS

OPTIMIZERS = ("sgd", "adadelta", "adam")
GEMM_CTAS = 148                # persistent tcgen05 GEMM CTAs on one B200 (one per SM)


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
    first, in the given order; never-trained ones after) and give the trained
    ones flat gradient views."""

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
        self.offset = {}
        o = 0
        with torch.no_grad():
            for p in order:
                n = p.numel()
                self.param[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.param[o:o + n].view_as(p.data)
                if id(p) in trained_ids:
                    p.grad = self.grad[o:o + n].view_as(p.data)
                self.offset[id(p)] = o
                o += n

    @property
    def trained_param(self):
        return self.param[:self.n_trained]

    @property
    def trained_grad(self):
        return self.grad[:self.n_trained]


class _Layer(object):
    """One GEMM layer of the tensor-core chain: fp32 master views + bf16 operands."""
    __slots__ = ("W", "b", "gW", "gb", "act", "n_out", "n_in", "wb")


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
        network._check_supported(training=True, fused=True)     # whatever the mode at construction
        self.network = network
        self.loss_spec = loss_spec
        self.kind, self.lr, self.momentum = optimizer_type, float(lr), float(momentum or 0.0)
        self.trunk, self.heads = _trained_layers(network)
        trained = []
        for W, b, _ in self.trunk:
            trained += [W, b]
        if self.heads:      # weights adjacent, biases adjacent: the two heads form one 2d-wide layer
            trained += [h[0][0] for h in self.heads] + [h[0][1] for h in self.heads]
        self.bucket = FlatBucket(network, trained)
        dev = self.bucket.param.device
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and \
            dist.is_initialized() else 1
        if self.world > 1:
            # every rank starts from rank 0's initialisation (each process draws its own
            # xavier weights unless the caller seeded them identically)
            dist.broadcast(self.bucket.param, src=dist.get_global_rank(process_group, 0)
                           if process_group is not None else 0, group=process_group)
        n = self.bucket.n_trained
        self.state0 = torch.zeros(n, dtype=torch.float32, device=dev)
        self.state1 = torch.zeros(n, dtype=torch.float32, device=dev) \
            if optimizer_type != "sgd" else None
        self._zbuf = torch.zeros(1, dtype=torch.int32, device=dev)     # [loss | dependency counters]
        self.loss_buf = self._zbuf[:1].view(torch.float32)
        self.step_count = 0
        self.precision = PRECISIONS[network.precision]
        # Dropout (abnet3/model.py:111, :136-141; the reference's default is p = 0.1): active while
        # the network is in train() mode, as in the reference; the keep masks are a function of a
        # device-side {seed, step} the kernels re-evaluate (include/abnet3_b200.h, abn_dropout)
        self.p_drop = float(getattr(network, "p_dropout", 0.0) or 0.0)
        self._drop_state = ops.dropout_state(dev) if self.p_drop > 0 else None
        self._rows = -1
        self._plans = {}
        self._graphs = {}
        self._warm = {}
        self._cursor = torch.zeros(2, dtype=torch.int64, device=dev)
        self._loss_acc = torch.zeros(1, dtype=torch.float64, device=dev)
        self._loss_cleared = False
        self._grads_clean = False
        self._dp = None
        self._dp_push = None
        if self.precision == 1:
            self._build_chain()
            # The gradient exchange is fused into the optimizer kernel over NVLink peer memory
            # (CUDA IPC, device-side step flags, graph-replayable).  It is latency bound at this
            # bucket size (2.8 MB); measured us/step (tools/dp_trace.py, 133 us without exchange):
            #   2 GPUs: flag-in-data two-shot ("ll") 154 | two-shot push 163 | one-shot push 168 | NCCL 173
            #   4 GPUs: ll 165 | two-shot push ~165
            #   8 GPUs: ll 177 | two-shot push 166 | one-shot push 197 | NCCL 206
            #           (mixed forms, tried and removed: ll slices + plain parameters 179, plain
            #           slices + ll parameters 171 -- at 8 GPUs the flag-in-data stores to 7 peers
            #           cost more than the fence + flag hop they replace)
            # auto: ll at world == 2 (no fences, no flag round trips; twice the bytes), two-shot push
            # beyond.  ABN_DP_P2P = ll | push | push2 | 1 (reads) | 0 (NCCL all-reduce in the graph).
            p2p = os.environ.get("ABN_DP_P2P", "auto")
            if self.world > 1 and p2p in ("push", "push2", "ll", "auto") and \
                    self.bucket.n_trained % 4 == 0:
                try:        # write-only exchange fused with the optimizer
                    one_shot = p2p == "push"
                    use_ll = p2p == "ll" or (p2p == "auto" and self.world <= 2)
                    self._dp_push = ops.dp_push_setup(self.bucket.param, self.bucket.n_trained, self.group,
                                                      one_shot=one_shot, ll=use_ll)
                except Exception as exc:
                    import warnings
                    warnings.warn("peer-memory data parallelism unavailable (%s); using NCCL" % exc)
                    self._dp_push = None
            elif self.world > 1 and p2p == "1":
                try:
                    self._dp = ops.dp_setup(self.bucket.grad, self.group)
                except Exception as exc:       # buffers not shareable: the NCCL all-reduce remains
                    import warnings
                    warnings.warn("peer-memory data parallelism unavailable (%s); using NCCL" % exc)
                    self._dp = None

    # -------------------------------------------------------------- dropout ---
    def _drop_on(self):
        return self.p_drop > 0 and self.network.training

    def _drop(self, layer):
        """ctypes abn_dropout of one layer for the current pass (None when inactive)."""
        if not self._drop_on():
            return None
        return ops.dropout_spec(self._drop_state, self.p_drop, layer)

    def _drop_advance(self):
        """Next step, next masks (after the last kernel that evaluates this step's)."""
        if self._drop_on():
            self._drop_state[1:].add_(1)

    def dropout_masks(self, rows):
        """Keep masks of every layer at the CURRENT step, in forward order (tests: replayed in
        the oracle).  bf16 path: the two multitask heads are one 2d-wide layer."""
        if self.precision == 1:
            widths = [L.n_out for L in self.chain]
        else:
            widths = [W.shape[0] for W, _, _ in self.trunk] + [h[0][0].shape[0] for h in self.heads]
        return [ops.dropout_mask(self._drop_state, self.p_drop, l, rows, w)
                for l, w in enumerate(widths)]

    # ------------------------------------------------------------ fp32 path ---
    # Buffers, GEMM problem lists and CUDA graphs of one batch size form a "plan"; plans are
    # cached per row count (LRU), so alternating batch sizes (train / dev sweeps, the ragged
    # token batches of OriginalDataLoader) do not rebuild or invalidate each other.
    _PLAN_ATTRS = ("acts", "dacts", "head_acts", "head_dacts", "xb", "actb", "dzb", "out_last",
                   "_zbuf", "loss_buf", "_dep", "_fwd_problems", "_dgrad_problems", "_fwd_fused",
                   "_fwd_rows", "_dgrad_fused", "_wgrad_split", "_backward_groups", "_gy",
                   "_graphs", "_warm", "_gsel", "_sx", "_sy", "_static_n", "_graph_fb",
                   "_graph_opt", "_eager_warm", "_pipe", "_fwd_fused_drop", "_dgrad_fused_drop")
    MAX_PLANS = 6

    def _reserve(self, rows):
        if rows == self._rows:
            return
        if self._rows >= 0:                       # park the active plan
            self._plans[self._rows] = {k: getattr(self, k) for k in self._PLAN_ATTRS
                                       if hasattr(self, k)}
        plan = self._plans.pop(rows, None)
        if plan is not None:
            for k, v in plan.items():
                setattr(self, k, v)
            self._rows = rows
            return
        while len(self._plans) >= self.MAX_PLANS:
            self._plans.pop(next(iter(self._plans)))
        dev = self.bucket.param.device
        self._graphs, self._warm = {}, {}
        self._gsel = self._sx = self._sy = self._static_n = None
        self._graph_fb = self._graph_opt = None
        self._pipe = None
        self._eager_warm = 0
        self._gy = [torch.empty(max(rows // 2, 1), dtype=torch.float32, device=dev)
                    for _ in range(2 if self.heads else 1)]

        def buf(width):
            return torch.empty((rows, width), dtype=torch.float32, device=dev)

        if self.precision == 0:
            zb = torch.zeros(1, dtype=torch.int32, device=dev)
            self._zbuf, self.loss_buf = zb, zb[:1].view(torch.float32)
            self.acts = [buf(W.shape[0]) for W, _, _ in self.trunk]
            self.dacts = [buf(W.shape[0]) for W, _, _ in self.trunk]
            self.head_acts = [[buf(W.shape[0]) for W, _, _ in h] for h in self.heads]
            self.head_dacts = [[buf(W.shape[0]) for W, _, _ in h] for h in self.heads]
        else:
            self._reserve_bf16(rows)
        self._rows = rows

    def _forward_fp32(self, x):
        h = x
        nt = len(self.trunk)
        for l, (W, b, act) in enumerate(self.trunk):
            h = ops.linear_forward(h, W.data, b.data, act, 0, out=self.acts[l], drop=self._drop(l))
        outs = []
        for hi, head in enumerate(self.heads):
            g = h
            for l, (W, b, act) in enumerate(head):
                g = ops.linear_forward(g, W.data, b.data, act, 0, out=self.head_acts[hi][l],
                                       drop=self._drop(nt + hi))
            outs.append(g)
        return outs if self.heads else h

    def _backward_fp32(self, x):
        trunk = self.trunk
        if self.heads:
            first = True
            for hi, head in enumerate(self.heads):
                for l in reversed(range(len(head))):
                    W, b, act = head[l]
                    xin = self.head_acts[hi][l - 1] if l > 0 else self.acts[-1]
                    dx = self.head_dacts[hi][l - 1] if l > 0 else self.dacts[-1]
                    ops.linear_backward(xin, W.data, self.head_acts[hi][l], self.head_dacts[hi][l],
                                        act, 0, dW=W.grad, db=b.grad, accumulate=False,
                                        dx=dx, accumulate_dx=(l == 0 and not first),
                                        drop=self._drop(len(trunk) + hi))
                first = False
        for l in reversed(range(len(trunk))):
            W, b, act = trunk[l]
            xin = self.acts[l - 1] if l > 0 else x
            ops.linear_backward(xin, W.data, self.acts[l], self.dacts[l], act, 0,
                                need_dx=(l > 0), dW=W.grad, db=b.grad, accumulate=False,
                                dx=self.dacts[l - 1] if l > 0 else None, drop=self._drop(l))
        self._drop_advance()

    # ------------------------------------------------- bf16 tensor-core path ---
    # Every contraction is one problem of the persistent grouped tcgen05 GEMM
    # (abn_gemm_bf16_group) on bf16 arrays in their natural layout: activations and dz
    # [rows, features] (+ a column of ones inside the row padding: the wgrad GEMM's extra
    # output column is the bias gradient), weights [n_out, n_in].
    def _build_chain(self):
        """Trunk layers, then (multitask) the two heads as ONE layer whose weight
        is the [2d, hidden] block they occupy side by side in the flat bucket."""
        dev = self.bucket.param.device
        P, G, off = self.bucket.param, self.bucket.grad, self.bucket.offset
        self.chain = []
        entries = []

        def add(W_view, b_view, gW, gb, act, w_off, b_off):
            L = _Layer()
            L.W, L.b, L.gW, L.gb, L.act = W_view, b_view, gW, gb, act
            L.n_out, L.n_in = W_view.shape
            L.wb = torch.zeros((L.n_out, ops.pad_row(L.n_in)), dtype=torch.bfloat16, device=dev)
            self.chain.append(L)
            entries.append((w_off, L.n_out * L.n_in, L.wb, L.n_in))
            entries.append((b_off, L.n_out, None, 0))

        for W, b, act in self.trunk:
            add(W.data, b.data, W.grad, b.grad, act, off[id(W)], off[id(b)])
        if self.heads:
            Ws = [h[0][0] for h in self.heads]
            bs = [h[0][1] for h in self.heads]
            d, hid = Ws[0].shape
            ow, ob = off[id(Ws[0])], off[id(bs[0])]
            assert off[id(Ws[1])] == ow + d * hid and off[id(bs[1])] == ob + d
            add(P[ow:ow + 2 * d * hid].view(2 * d, hid), P[ob:ob + 2 * d],
                G[ow:ow + 2 * d * hid].view(2 * d, hid), G[ob:ob + 2 * d], self.heads[0][0][2],
                ow, ob)
            self.head_dim = d
        self._segments = ops.param_segments(entries)
        self._grads_clean = False
        self.refresh_bf16_weights()
        # load_state_dict / load_network copy into the fp32 masters in place: redo the bf16
        # operand copies (the fused optimizer only keeps them current during training)
        import weakref
        me = weakref.ref(self)

        def _after_load(module, incompatible_keys):
            eng = me()
            if eng is not None and eng.network is module:
                eng.refresh_bf16_weights()

        self._load_hook = self.network.register_load_state_dict_post_hook(_after_load)

    def refresh_bf16_weights(self):
        """bf16 operand copies of the fp32 master weights (after load_state_dict; the
        fused optimizer keeps them current during training)."""
        for L in self.chain:
            ops.cast_bf16(L.W, L.wb, None)

    def _reserve_bf16(self, rows):
        dev = self.bucket.param.device

        def b16(r, c, ones_at=None):
            t = torch.zeros((r, c), dtype=torch.bfloat16, device=dev)
            if ones_at is not None:
                t[:, ones_at] = 1.0
            return t

        d_in = self.chain[0].n_in
        self.xb = b16(rows, ops.pad_row(d_in + 1), d_in)
        # hidden activations and every layer's dz: bf16; the embeddings: fp32
        self.actb = [b16(rows, ops.pad_row(L.n_out + 1)) for L in self.chain[:-1]]
        self.dzb = [b16(rows, ops.pad_row(L.n_out)) for L in self.chain]
        n_last = self.chain[-1].n_out
        self.out_last = torch.empty((rows, n_last), dtype=torch.float32, device=dev)
        self.acts = [None] * (len(self.chain) - 1) + [self.out_last]
        last = len(self.chain) - 1
        # Layers are chained INSIDE a launch: a problem's tiles signal per 256-row block when
        # their output rows are in global memory, the next layer's tiles wait for their block
        # only -- no grid-wide barrier, no launch per layer (groups of up to 4 problems).
        tiles_m = (rows + 255) // 256
        n_dep = 2 * len(self.chain) + 2
        zbuf = torch.zeros(1 + n_dep * tiles_m, dtype=torch.int32, device=dev)
        self._zbuf = zbuf                  # one contiguous block: the gather kernel clears it
        self.loss_buf = zbuf[:1].view(torch.float32)
        self._dep = zbuf[1:].view(n_dep, tiles_m)
        self._fwd_problems, self._dgrad_problems, wgrad = [], [], []
        hb = self.xb
        G = ops.GEMM_MAX_GROUP
        for l, L in enumerate(self.chain):
            out = self.actb[l] if l < last else self.out_last
            self._fwd_problems.append(ops.gemm_problem(
                hb, L.wb, rows, L.n_out, L.n_in, ops.GE_BIAS_ACT, out, act=L.act, bias=L.b,
                ones_col=(l < last),
                signal=self._dep[l] if (l < last and (l + 1) % G != 0) else None,
                wait=self._dep[l - 1] if (l > 0 and l % G != 0) else None))
            wgrad.append((L, hb))
            hb = out
        # Forward pass in ONE launch with the activations resident in shared memory
        # (abn_mlp_forward_fused) when every layer fits its 512-feature slab; same bits as the
        # chained launch, which serves everything else (ABN_FWD_FUSED=0 forces it).
        self._fwd_fused = None
        self._fwd_rows = rows
        fits = all(L.n_in <= ops.MLP_MAX_WIDTH and L.n_out + 1 <= ops.MLP_MAX_WIDTH for L in self.chain)
        self._fwd_fused_drop = None
        if fits and len(self.chain) <= ops.MLP_MAX_LAYERS and os.environ.get("ABN_FWD_FUSED", "1") != "0":
            self._fwd_fused = ops.mlp_layers(
                [(L.wb, L.n_in, L.b, L.act, self.actb[l] if l < last else self.out_last, l < last)
                 for l, L in enumerate(self.chain)])
            if self.p_drop > 0:
                self._fwd_fused_drop = ops.mlp_layers(
                    [(L.wb, L.n_in, L.b, L.act, self.actb[l] if l < last else self.out_last, l < last,
                      (self._drop_state, self.p_drop, l)) for l, L in enumerate(self.chain)])
        # dz of the layer below = (dz W) * act'(its output); every dgrad problem signals its row
        # blocks: the next dgrad problem AND the weight gradients of the layer below wait on them
        n_d = 0
        base = len(self.chain)
        dz_ready = {last: None}            # layer -> counters that say "dz of this layer is written"
        for l in range(last, 0, -1):
            L = self.chain[l]
            self._dgrad_problems.append(ops.gemm_problem(
                self.dzb[l], L.wb, rows, L.n_in, L.n_out, ops.GE_DACT, self.dzb[l - 1], b_mn=True,
                act=self.chain[l - 1].act, yprev=self.actb[l - 1],
                signal=self._dep[base + n_d], wait=dz_ready[l]))
            dz_ready[l - 1] = self._dep[base + n_d]
            n_d += 1
        # dz chain in ONE launch with dz resident in shared memory (abn_mlp_dgrad_fused) when the
        # layers fit its slab (ABN_BWD_FUSED=0: the grouped GEMM instead)
        self._dgrad_fused = None
        self._dgrad_fused_drop = None
        if fits and last >= 1 and last <= ops.MLP_MAX_LAYERS and os.environ.get("ABN_BWD_FUSED", "1") != "0":
            self._dgrad_fused = ops.mlp_dlayers(
                [(self.chain[l].wb, self.chain[l].n_in, self.chain[l - 1].act, self.actb[l - 1],
                  self.dzb[l - 1]) for l in range(last, 0, -1)])
            if self.p_drop > 0:
                self._dgrad_fused_drop = ops.mlp_dlayers(
                    [(self.chain[l].wb, self.chain[l].n_in, self.chain[l - 1].act, self.actb[l - 1],
                      self.dzb[l - 1], (self._drop_state, self.p_drop, l - 1))
                     for l in range(last, 0, -1)])
        if self.p_drop > 0 and (self._fwd_fused is None or (last >= 1 and self._dgrad_fused is None)):
            raise NotImplementedError(
                "dropout with p > 0 on the bf16 path is implemented by the fused chain kernels only "
                "(layer widths <= %d, at most %d layers); use precision='fp32' for this network"
                % (ops.MLP_MAX_WIDTH, ops.MLP_MAX_LAYERS))
        G = ops.GEMM_MAX_GROUP
        merge = (self._dgrad_fused is None and 2 * len(self.chain) - 1 <= G and
                 os.environ.get("ABN_BWD_MERGE", "1") != "0")
        # weight + bias gradients of ALL layers, split over the batch: ONE wave of CTA pairs when
        # they have a launch of their own (step: split 5 146.2 us | 9 149.8 | 10 147.9 | 6 159.1),
        # ~3.5 waves inside the merged backward launch (measured; ABN_WGRAD_SPLIT overrides)
        tiles = sum(((L.n_out + 255) // 256) * ((L.n_in + 1 + 255) // 256) for L, _ in wgrad)
        waves4 = 14 if merge else 4
        split = max(1, min((rows + 63) // 64, int(os.environ.get("ABN_WGRAD_SPLIT", "0")) or
                           max(1, (waves4 * (GEMM_CTAS // 2) // 4) // max(1, tiles))))
        self._wgrad_split = split

        def wgrad_problem(l, wait):
            L, xin = wgrad[l]
            return ops.gemm_problem(self.dzb[l], xin, L.n_out, L.n_in, rows, ops.GE_ATOMIC, L.gW,
                                    a_mn=True, b_mn=True, split_k=split, ones_out=L.gb, wait=wait)

        # Grouped-GEMM launches of the backward pass.  Merged: dgrad(l), wgrad(l), dgrad(l-1), ...
        # in ONE launch -- the weight-gradient tiles (whose dz is a layer older) fill the gaps the
        # dgrad chain's dependencies leave.  Otherwise the dgrad chain(s) (unless the fused dgrad
        # kernel runs), then the wgrad group(s).
        self._backward_groups = []
        if merge:
            merged = []
            for k, l in enumerate(range(last, 0, -1)):
                merged.append(self._dgrad_problems[k])
                merged.append(wgrad_problem(l, dz_ready[l]))
            merged.append(wgrad_problem(0, dz_ready[0]))
            self._backward_groups.append(merged)
        else:
            if self._dgrad_fused is None:
                # a wait may only name a counter signalled inside the same group
                for i in range(0, len(self._dgrad_problems), G):
                    grp = self._dgrad_problems[i:i + G]
                    grp[0].wait = None
                    self._backward_groups.append(grp)
            probs = [wgrad_problem(l, None) for l in range(len(wgrad))]
            for i in range(0, len(probs), G):
                self._backward_groups.append(probs[i:i + G])
        # Pipelined sweeps (sweep_table): the gather of batch k + 1 runs beside the step of batch
        # k, so the input operand, its labels and the [loss | counters] block exist twice.  Only
        # the layer-0 weight-gradient problem reads x: the second set of backward groups differs
        # from the first in that operand alone.
        self._pipe = None
        if self._fwd_fused is not None and self._dgrad_fused is not None and not merge:
            xb2 = b16(rows, ops.pad_row(d_in + 1), d_in)
            zbuf2 = torch.zeros_like(self._zbuf)
            gy2 = [torch.empty_like(t) for t in self._gy]
            xb1 = self.xb
            wgrad[0] = (wgrad[0][0], xb2)
            probs2 = [wgrad_problem(l, None) for l in range(len(wgrad))]
            wgrad[0] = (wgrad[0][0], xb1)
            groups2 = [probs2[i:i + G] for i in range(0, len(probs2), G)]
            self._pipe = [dict(xb=self.xb, _zbuf=self._zbuf, _gy=self._gy,
                               _backward_groups=self._backward_groups),
                          dict(xb=xb2, _zbuf=zbuf2, _gy=gy2, _backward_groups=groups2)]

    def _forward_bf16(self, x):
        if x is not None:       # fp32 batch -> bf16 A operand (the gather can also write it directly)
            ops.cast_bf16(x, self.xb, None)
        if not self._loss_cleared:          # (the gather kernel clears loss + counters together)
            self._dep.zero_()
        if self._fwd_fused is not None:
            ops.mlp_forward_fused(self.xb, self._fwd_rows,
                                  self._fwd_fused_drop if self._drop_on() else self._fwd_fused)
        else:
            G = ops.GEMM_MAX_GROUP
            for i in range(0, len(self._fwd_problems), G):
                ops.gemm_group(self._fwd_problems[i:i + G])
        if self.heads:
            d = self.head_dim
            return [self.out_last[:, :d], self.out_last[:, d:]]
        return self.out_last

    def _backward_bf16(self, x):
        if self._dp is not None:
            ops.dp_grad_reset(self.bucket.trained_grad, self._dp)    # once no peer reads it any more
        elif not self._grads_clean:
            self.bucket.trained_grad.zero_()       # dW / db are accumulated with reds
        self._grads_clean = False
        if self._dgrad_fused is not None:
            ops.mlp_dgrad_fused(self.dzb[-1], self._fwd_rows,
                                self._dgrad_fused_drop if self._drop_on() else self._dgrad_fused)
        self._drop_advance()
        for grp in self._backward_groups:
            ops.gemm_group(grp)

    # -------------------------------------------------------------- common ---
    def forward(self, x):
        """x [rows, input_dim] -> embeddings (last layer, or the two heads)."""
        self._reserve(x.shape[0])
        return self._forward_bf16(x) if self.precision == 1 else self._forward_fp32(x)

    def backward(self, x):
        return self._backward_bf16(x) if self.precision == 1 else self._backward_fp32(x)

    def _loss_and_seed(self, out, n, labels):
        """loss into self.loss_buf, d(loss)/d(embeddings) into the seed buffers."""
        if not self._loss_cleared:
            self.loss_buf.zero_()
        self._loss_cleared = False
        if self.precision == 1:
            return self._loss_and_seed_bf16(out, n, labels)
        if not self.heads:
            kind, margin, avg = self.loss_spec
            de = self.dacts[-1]
            ops.pair_loss(out[:n], out[n:], labels[0], kind, margin, 1.0 / n if avg else 1.0,
                          loss_out=self.loss_buf, grads=(de[:n], de[n:]))
            return
        spec_spk, spec_phn, weight = self.loss_spec
        d = out[0].shape[1]
        for hi, (spec, w, y) in enumerate(((spec_spk, weight, labels[0]),
                                           (spec_phn, 1.0 - weight, labels[1]))):
            kind, margin, avg = spec
            de = self.head_dacts[hi][-1]
            ops.pair_loss(out[hi][:n], out[hi][n:], y, kind, margin,
                          w * (1.0 / n if avg else 1.0), loss_out=self.loss_buf,
                          grads=(de[:n], de[n:]))

    def _loss_and_seed_bf16(self, out, n, labels):
        """Tensor-core path: the loss kernel writes the output layer's dz (bf16) directly."""
        dz = self.dzb[-1]
        act = self.chain[-1].act
        drop = self._drop(len(self.chain) - 1)
        if not self.heads:
            kind, margin, avg = self.loss_spec
            ops.pair_loss_dz(out[:n], out[n:], labels[0], dz[:n], dz[n:], kind, margin,
                             1.0 / n if avg else 1.0, act, loss_out=self.loss_buf, drop=drop,
                             row2_offset=n)
            return
        spec_spk, spec_phn, weight = self.loss_spec
        d = self.head_dim
        for hi, (spec, w, y) in enumerate(((spec_spk, weight, labels[0]),
                                           (spec_phn, 1.0 - weight, labels[1]))):
            kind, margin, avg = spec
            dzh = dz[:, hi * d:(hi + 1) * d]
            ops.pair_loss_dz(out[hi][:n], out[hi][n:], y, dzh[:n], dzh[n:], kind, margin,
                             w * (1.0 / n if avg else 1.0), act, loss_out=self.loss_buf, drop=drop,
                             row2_offset=n, col_offset=hi * d)

    def _grad_scale(self):
        """Ranks SUM their gradients; a loss averaged over the batch (``avg=True``) is averaged
        over the global batch by scaling the sum with 1 / world."""
        if not self.heads:
            avg = self.loss_spec[2]
        else:
            avg = self.loss_spec[0][2]
            if self.world > 1 and bool(self.loss_spec[1][2]) != bool(avg):
                raise ValueError("data-parallel multitask training needs the same `avg` flag "
                                 "on both head losses")
        return 1.0 / self.world if (self.world > 1 and avg) else 1.0

    def _allreduce(self):
        """Sum the gradient bucket over the ranks -- unless the optimizer kernel does it
        itself over NVLink peer memory (self._dp)."""
        if self.world > 1 and self._dp is None and self._dp_push is None:
            dist.all_reduce(self.bucket.trained_grad, op=dist.ReduceOp.SUM, group=self.group)

    def _optimizer(self, scale, step):
        if self.precision == 1 and self._dp_push is not None:
            # reduce-scatter + update + all-gather of the parameters as NVLink writes, bf16
            # weight copies and the gradient reset: one kernel
            ops.dp_push_step(self.bucket.grad, self.state0, self.state1, self.kind, self.lr,
                             self.momentum, scale, step, self._segments, self._dp_push)
            self._grads_clean = True
            return
        if self.precision == 1 and self._dp is not None:
            # all-reduce fused into the update: every rank reads its peers' buckets directly
            ops.dp_optimizer_step(self.bucket.param, self.state0, self.state1, self.kind, self.lr,
                                  self.momentum, scale, step, self._segments, self._dp)
            return
        if self.precision == 1:     # update + bf16 operand copies + gradient reset in one kernel
            ops.optimizer_step_fused(self.bucket.param, self.bucket.grad, self.state0, self.state1,
                                     self.kind, self.lr, self.momentum, scale, step,
                                     self._segments, zero_grad=True)
            self._grads_clean = True
            return
        ops.optimizer_step(self.bucket.trained_param, self.bucket.trained_grad, self.state0,
                           self.state1, self.kind, self.lr, self.momentum, scale, step)

    # ---- CUDA graphs ----------------------------------------------------------
    # The step is a handful of small launches on fixed buffers; replaying them as one
    # graph removes the launch latency that otherwise dominates a sub-millisecond step.
    def input_buffers(self, n):
        """Static (x [2n, D], labels...) buffers of the graphed step: producers
        (abn_gather_batch) may write straight into them."""
        self._reserve(2 * n)
        if getattr(self, "_static_n", None) != n or self._sx is None:
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

    def step_graphed(self, x, n, *labels):
        """Same as step(do_training=True) through CUDA graphs (SGD / Adadelta)."""
        self._check_trainable()
        bufs = self.input_buffers(n)
        for dst, src in zip(bufs, (x,) + tuple(labels)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        scale = self._grad_scale()
        if self._graph_fb is None:
            if self._eager_warm < 2:           # lazy one-time setup must not be captured
                self._eager_warm += 1
                self._fwd_loss_bwd()
                if self.world > 1:
                    self._allreduce()
                self._optimizer(scale, 1)
                self.step_count += 1
                return self.loss_buf
            torch.cuda.synchronize()
            self._graph_fb, self._graph_opt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_fb):
                self._fwd_loss_bwd()
            with torch.cuda.graph(self._graph_opt):
                self._optimizer(scale, 1)
        self._graph_fb.replay()
        if self.world > 1:
            self._allreduce()
        self._graph_opt.replay()
        self.step_count += 1
        return self.loss_buf

    # ---- batch generation fused in front (tensor-core path) --------------------------
    # A frame-pair table is (idx1, idx2 int32 global rows, y int8) or, for the multitask
    # network, (idx1, idx2, y_spk, y_phn): what FramesDataLoader keeps on the device.
    def gather_buffers(self, n):
        """Static ``sel`` [n] int64 buffer of step_gather: write the batch's positions
        in the frame-pair table into it (e.g. ``sel.copy_(perm[lo:lo + n])``)."""
        self._reserve(2 * n)
        if getattr(self, "_gsel", None) is None or self._gsel.numel() != n:
            self._gsel = torch.zeros(n, dtype=torch.int64, device=self.bucket.param.device)
        return self._gsel

    def _check_table(self, table):
        want = 4 if self.heads else 3
        if len(table) != want:
            raise ValueError("frame-pair table of %d arrays expected (idx1, idx2, %s)"
                             % (want, "y_spk, y_phn" if self.heads else "y"))

    def _check_trainable(self):
        """Training-mode restrictions of the kernels, re-checked at every training step (the
        network may have been built or put in eval() before the engine was)."""
        self.network._check_supported(training=True)

    def _fuse_loss(self):
        """The table-driven step computes the loss inside the forward chain kernel (interleaved
        pair rows, ops.mlp_forward_loss_fused) when the network is a single-head chain of at most
        128 outputs and dropout is off; the embeddings then never reach HBM."""
        return (self._fwd_fused is not None and not self.heads and self.chain[-1].n_out <= 128
                and not self._drop_on() and os.environ.get("ABN_FUSE_LOSS", "1") != "0")

    def _fwd_loss(self, n):
        """forward + loss + dz of the output layer on the gathered operand set."""
        self._loss_cleared = True
        if self._fuse_loss():
            kind, margin, avg = self.loss_spec
            ops.mlp_forward_loss_fused(self.xb, self._fwd_rows, self._fwd_fused, self._gy[0],
                                       self.dzb[-1], kind, margin, 1.0 / n if avg else 1.0,
                                       loss_out=self.loss_buf)
            self._loss_cleared = False
            return
        out = self._forward_bf16(None)
        self._loss_and_seed(out, n, self._gy)

    def _table_fwd_loss(self, feat, table, n, sel, train):
        """gather -> forward -> loss [-> backward] of one batch of the table."""
        self._reserve(2 * n)
        ys = table[2:]
        two = len(ys) > 1
        # the gather kernel writes the first layer's bf16 operand, adds the previous step's loss
        # to the sweep accumulator and clears loss + dependency counters
        ops.gather_batch_bf16(feat, table[0], table[1], ys[0], sel, n, self.xb, y_out=self._gy[0],
                              zero=self._zbuf, y2=ys[1] if two else None,
                              y2_out=self._gy[1] if two else None,
                              cursor=None if sel is not None else self._cursor,
                              loss_acc=self._loss_acc, interleave=self._fuse_loss())
        self._fwd_loss(n)
        if train:
            self.backward(None)
        else:
            self._drop_advance()

    def _table_step(self, feat, table, n, sel=None, train=True, graph=True):
        if self.precision != 1:
            raise ValueError("the gather-fused step serves the bf16 tensor-core path "
                             "(network.precision == 'bf16')")
        self._check_table(table)
        if train:
            self._check_trainable()
        self._reserve(2 * n)
        scale = self._grad_scale()
        key = (feat.data_ptr(),) + tuple(t.data_ptr() for t in table) + \
            (None if sel is None else sel.data_ptr(), bool(train), self._drop_on())
        use_graph = graph and (self.kind != "adam" or not train)
        g = self._graphs.get(key) if use_graph else None
        if g is not None:
            g[0].replay()
            if g[1] is not None:                    # NCCL all-reduce between two graphs
                self._allreduce()
                g[1].replay()
            self.step_count += int(train)
            return self.loss_buf
        if use_graph and self._warm.get(key, 0) >= 2:
            torch.cuda.synchronize()
            fused_ar = self.world > 1 and os.environ.get("ABN_GRAPH_ALLREDUCE", "1") == "1"
            g_main, g_opt = torch.cuda.CUDAGraph(), None
            if not train or self.world == 1 or fused_ar:
                # ONE graph: gather .. backward, the all-reduce (NCCL capture, or fused into
                # the optimizer kernel over peer memory), optimizer
                with torch.cuda.graph(g_main):
                    self._table_fwd_loss(feat, table, n, sel, train)
                    if train:
                        if self.world > 1:
                            self._allreduce()
                        self._optimizer(scale, 1)
            else:
                g_opt = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_main):
                    self._table_fwd_loss(feat, table, n, sel, train)
                with torch.cuda.graph(g_opt):
                    self._optimizer(scale, 1)
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = (g_main, g_opt)
            return self._table_step(feat, table, n, sel, train, graph)
        self._warm[key] = self._warm.get(key, 0) + 1
        self._table_fwd_loss(feat, table, n, sel, train)
        if train:
            if self.world > 1:
                self._allreduce()
            self.step_count += 1
            self._optimizer(scale, self.step_count)
        return self.loss_buf

    def step_gather(self, feat, idx1, idx2, y, n, graph=True, y2=None, do_training=True):
        """One training step on the batch ``sel`` (gather_buffers) of the device-resident
        frame-pair table (idx1, idx2, y int8 [, y2 = y_phn when y = y_spk]) over ``feat``:
        gather -> forward -> loss -> backward [-> all-reduce] -> optimizer, replayed as a
        CUDA graph after two eager steps.  What FramesDataLoader.load_batch +
        TrainerSiamese.optimize_model do per batch (abnet3/dataloader.py:673-684,
        abnet3/trainer.py:226-243)."""
        sel = self.gather_buffers(n)
        table = (idx1, idx2, y) if y2 is None else (idx1, idx2, y, y2)
        return self._table_step(feat, table, n, sel, do_training, graph)

    def sweep_table(self, feat, table, n, n_batches, start=0, do_training=True, graph=True):
        """One sweep of abnet3/trainer.py:229-243 over ``n_batches`` consecutive batches of
        ``n`` rows of an (already shuffled) device-resident frame-pair table, starting at
        row ``start``.  The batch position lives on the device (the gather kernel advances
        it), so after two eager steps the sweep is ``n_batches`` replays of one CUDA graph
        with no other per-batch work.  Returns the summed loss as a float64 device tensor
        [1] (read it once per sweep; the reference synchronises on every step, :242)."""
        if n_batches <= 0:
            return torch.zeros(1, dtype=torch.float64, device=self.bucket.param.device)
        if start + n * n_batches > table[0].numel():
            raise ValueError("the sweep runs past the end of the frame-pair table")
        self._reserve(2 * n)
        self._cursor.copy_(torch.tensor([start, 0], dtype=torch.int64), non_blocking=True)
        self._loss_acc.zero_()
        if self._can_pipeline(do_training):
            return self._sweep_pipelined(feat, table, n, n_batches, do_training, graph)
        self.loss_buf.zero_()
        for _ in range(n_batches):
            self._table_step(feat, table, n, None, do_training, graph)
        return self._loss_acc + self.loss_buf.double()

    # ---- pipelined sweep: the next batch is gathered beside the current step ------------
    # The gather (16 384 random 1 120-byte rows: HBM latency bound, ~12 us) depends on nothing
    # the step computes, and the forward / dgrad chain kernels leave 20 of the 148 SMs idle
    # (64 row blocks on 74 CTA pairs): batch k + 1 is gathered on a side stream, into the second
    # operand buffer, while batch k trains.  Inside the CUDA graph that is a fork / join.
    def _can_pipeline(self, train):
        if os.environ.get("ABN_PIPELINE", "1") == "0" or self.precision != 1:
            return False
        if getattr(self, "_pipe", None) is None:
            return False
        if train and self.kind == "adam":
            return False
        # the NCCL all-reduce between two graphs keeps the plain sequence
        if train and self.world > 1 and self._dp is None and self._dp_push is None and \
                os.environ.get("ABN_GRAPH_ALLREDUCE", "1") != "1":
            return False
        return True

    def _use_parity(self, p):
        for k, v in self._pipe[p].items():
            setattr(self, k, v)
        self.loss_buf = self._zbuf[:1].view(torch.float32)

    def _pipe_gather(self, feat, table, n, p):
        """Gather the batch at the device cursor into operand set p (adds that set's previous
        loss to the sweep accumulator and clears it)."""
        q = self._pipe[p]
        ys = table[2:]
        two = len(ys) > 1
        ops.gather_batch_bf16(feat, table[0], table[1], ys[0], None, n, q["xb"], y_out=q["_gy"][0],
                              zero=q["_zbuf"], y2=ys[1] if two else None,
                              y2_out=q["_gy"][1] if two else None, cursor=self._cursor,
                              loss_acc=self._loss_acc, table_rows=table[0].numel(),
                              interleave=self._fuse_loss())

    def _pipe_step(self, feat, table, n, p, train, scale, step):
        """Step on operand set p; the gather of the NEXT batch into set 1 - p runs on the side
        stream meanwhile."""
        main = torch.cuda.current_stream()
        side = self._side_stream()
        # fork at the very start of the step: the gather then runs beside the forward chain (a
        # fork later in the sequence cuts the programmatic-dependent-launch chain of the main
        # stream and gains nothing: 134.5 us/step here against 141-143, tools/time_sweep.py)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self._pipe_gather(feat, table, n, 1 - p)
        self._use_parity(p)
        self._fwd_loss(n)
        if train:
            self.backward(None)
            if self.world > 1:
                self._allreduce()
            self._optimizer(scale, step)
        else:
            self._drop_advance()
        main.wait_stream(side)

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.bucket.param.device)
        return self._side

    def _sweep_pipelined(self, feat, table, n, n_batches, train, graph):
        self._check_table(table)
        if train:
            self._check_trainable()
        scale = self._grad_scale()
        for q in self._pipe:
            q["_zbuf"][:1].zero_()
        key = ("pipe", feat.data_ptr()) + tuple(t.data_ptr() for t in table) + \
            (bool(train), self._drop_on())
        self._pipe_gather(feat, table, n, 0)                  # batch 0 -> set 0
        for b in range(n_batches):
            p = b & 1
            g = self._graphs.get(key + (p,)) if graph else None
            if g is not None:
                g[0].replay()
            elif graph and self._warm.get(key + (p,), 0) >= 1:
                torch.cuda.synchronize()
                g_main = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_main):
                    self._pipe_step(feat, table, n, p, train, scale, 1)
                if len(self._graphs) >= 12:
                    self._graphs.pop(next(iter(self._graphs)))
                self._graphs[key + (p,)] = (g_main, None)
                g_main.replay()
            else:
                self._warm[key + (p,)] = self._warm.get(key + (p,), 0) + 1
                self._pipe_step(feat, table, n, p, train, scale, self.step_count + 1)
            self.step_count += int(train)
        total = self._loss_acc + self._pipe[0]["_zbuf"][:1].view(torch.float32).double() + \
            self._pipe[1]["_zbuf"][:1].view(torch.float32).double()
        self._use_parity(0)
        return total

    def step(self, x, n, *labels, do_training=True, graph=False):
        """x = [X1; X2] as one [2n, D] batch; labels float32 [n] (y) or
        (y_spk, y_phn).  Returns the loss as a 1-element device tensor that is
        overwritten by the next step.  graph=True replays CUDA graphs (not for Adam,
        whose bias correction changes every step)."""
        if do_training:
            self._check_trainable()
        if graph and do_training and self.kind != "adam":
            return self.step_graphed(x, n, *labels)
        out = self.forward(x)
        self._loss_and_seed(out, n, labels)
        if not do_training:
            self._drop_advance()
            return self.loss_buf
        self.backward(x)
        if self.world > 1:
            self._allreduce()
        self.step_count += 1
        self._optimizer(self._grad_scale(), self.step_count)
        return self.loss_buf
