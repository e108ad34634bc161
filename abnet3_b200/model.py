"""Siamese embedders of /root/reference/abnet3/model.py on the sm_100a kernels.

``SiameseNetwork`` (:82-208) and ``SiameseMultitaskNetwork`` (:211-376) keep
the reference's constructor arguments, module tree (hence ``state_dict`` key
names: ``input_emb.0.*``, ``hidden_layers.{0,3,..}.*``, ``output_layer.0.*`` /
``hidden_layers_shared.*``, ``output_layer_spk.0.*``, ``output_layer_phn.0.*``,
and the never-applied ``hidden_layers_spk/phn`` stacks), initialisation,
``forward`` / ``forward_once`` / ``whoami`` / ``save_network`` /
``load_network``.  Every ``Linear -> Dropout(p=0) -> activation`` block runs as
ONE kernel (abn_linear_forward: GEMM + bias + activation) and its backward as
abn_linear_backward; both branches of the siamese pair go through the layers
as a single 2B-row batch.

Dropout (the reference's default is p = 0.1) runs inside the kernels' epilogues in train()
mode: the keep mask is a counter-based hash of (seed, step, layer, row, col), re-evaluated by
the backward kernels (include/abnet3_b200.h, abn_dropout).  ``batch_norm=True`` works in eval
mode (embedding a trained network: the running statistics are folded into W and b); BatchNorm
with batch statistics (training mode) raises instead of silently falling back.
"""
import torch
import torch.nn as nn

from . import ops

activation_functions = {'relu': nn.ReLU,
                        'sigmoid': nn.Sigmoid,
                        'tanh': nn.Tanh,
                        'softmax': nn.Softmax,
                        }

init_functions = {'xavier_uni': nn.init.xavier_uniform_,
                  'xavier_normal': nn.init.xavier_normal_,
                  'orthogonal': nn.init.orthogonal_}

PRECISIONS = {"fp32": 0, "bf16": 1}


class _LinearActFn(torch.autograd.Function):
    """y = act(x W^T + b) with the fused backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, precision, drop=None):
        x = x.contiguous()
        ctx.drop = drop
        if drop is not None:
            # dropout between the GEMM and the activation: the fp32 kernels evaluate the mask in
            # their epilogues (the tensor-core version lives in the fused training engine)
            y = ops.linear_forward(x, weight, bias, act, 0, drop=drop)
        elif precision == 1:
            # tensor-core forward (tcgen05): bf16 operand copies made on the fly; the
            # training engine keeps them resident instead (abnet3_b200.engine)
            m, n_in = x.shape
            n_out = weight.shape[0]
            xb = torch.empty((m, ops.pad_row(n_in)), dtype=torch.bfloat16, device=x.device)
            wb = torch.empty((n_out, ops.pad_row(n_in)), dtype=torch.bfloat16, device=x.device)
            ops.cast_bf16(x, xb)
            ops.cast_bf16(weight.detach(), wb)
            y = torch.empty((m, n_out), dtype=torch.float32, device=x.device)
            ops.gemm_group([ops.gemm_problem(xb, wb, m, n_out, n_in, ops.GE_BIAS_ACT, y, act=act,
                                             bias=bias.detach())])
        else:
            y = ops.linear_forward(x, weight, bias, act, 0)
        ctx.save_for_backward(x, weight, y)
        ctx.act, ctx.precision = act, precision
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        dz = dy.contiguous().clone()          # overwritten with dy * act'(y)
        # autograd backward always takes the fp32 kernels (the fused engine has the
        # tensor-core backward)
        dx, dW, db = ops.linear_backward(x, weight, y, dz, ctx.act, 0,
                                         need_dx=ctx.needs_input_grad[0], drop=ctx.drop)
        return dx, dW, db, None, None, None


def _joint(input1, input2):
    """Both branches as one [2B, D] batch; no copy when the two inputs are the
    halves of one buffer (what the dataloaders of this package yield)."""
    if (input1.is_contiguous() and input2.is_contiguous() and input1.dtype == input2.dtype
            and input2.data_ptr() == input1.data_ptr() + input1.numel() * input1.element_size()
            and not input1.requires_grad and not input2.requires_grad
            and input1._base is not None and input1._base is input2._base):
        base = input1._base
        off = (input1.data_ptr() - base.data_ptr()) // input1.element_size()
        flat = base.reshape(-1)[off:off + 2 * input1.numel()]
        return flat.view(2 * input1.shape[0], input1.shape[1])
    return torch.cat([input1, input2], 0)


class NetworkBuilder(nn.Module):
    """Generic network class (abnet3/model.py:31-79)."""

    def __init__(self, *args, **kwargs):
        super(NetworkBuilder, self).__init__()

    def forward_once(self, *args, **kwargs):
        raise NotImplementedError('Unimplemented forward_once for class:',
                                  self.__class__.__name__)

    def forward(self, *args, **kwargs):
        raise NotImplementedError('Unimplemented forward for class:',
                                  self.__class__.__name__)

    def whoami(self, *args, **kwargs):
        return {'params': self.__dict__, 'class_name': self.__class__.__name__}

    def init_weight_method(self, layer):
        # abnet3/model.py:172-177
        if isinstance(layer, nn.Linear):
            init_func = init_functions[self.type_init]
            init_func(layer.weight.data,
                      gain=nn.init.calculate_gain(self.activation_layer))
            layer.bias.data.fill_(0.0)

    # -- kernel path helpers ------------------------------------------------
    def _check_supported(self, training=None, fused=False):
        training = self.training if training is None else training
        if self.batch_norm and training:
            raise NotImplementedError(
                "batch_norm=True in training mode (batch statistics) is not implemented by the "
                "sm_100a kernels; in eval mode the running statistics are folded into the layer")

    # -- dropout (abnet3/model.py:136-141: Linear -> Dropout(p) -> act) -------------
    def _drop_for_call(self, device):
        """One forward pass in train() mode = one set of masks: a SNAPSHOT of the {seed, step}
        state that the forward and the (later) backward kernels of this pass both read; the
        network's own counter moves on."""
        if not (self.training and self.p_dropout > 0):
            return None
        st = self.__dict__.get("_drop_state")
        if st is None or st.device != device:
            st = ops.dropout_state(device)
            self.__dict__["_drop_state"] = st
        snap = st.clone()
        st[1:].add_(1)
        return snap

    def _drop_spec(self, snap, layer):
        return None if snap is None else ops.dropout_spec(snap, float(self.p_dropout), layer)

    @staticmethod
    def _blocks(seq):
        """[(Linear, BatchNorm1d or None)] of a Sequential of `Linear -> Dropout -> [BatchNorm1d]
        -> [act]` groups (abnet3/model.py:132-141)."""
        mods = list(seq)
        out = []
        for k, m in enumerate(mods):
            if isinstance(m, nn.Linear):
                bn = None
                for nxt in mods[k + 1:]:
                    if isinstance(nxt, nn.Linear):
                        break
                    if isinstance(nxt, nn.BatchNorm1d):
                        bn = nxt
                        break
                out.append((m, bn))
        return out

    @staticmethod
    def _effective(lin, bn):
        """(weight, bias) of the layer as the kernels see it.  Eval-mode BatchNorm1d is the
        affine map (z - mean) / sqrt(var + eps) * gamma + beta on the layer's pre-activation
        (the Dropout between them is the identity in eval mode): folded into W and b."""
        if bn is None:
            return lin.weight, lin.bias
        s = torch.rsqrt(bn.running_var + bn.eps)
        if bn.weight is not None:
            s = s * bn.weight
        W = (lin.weight * s.unsqueeze(1)).contiguous()
        b = (lin.bias - bn.running_mean) * s
        if bn.bias is not None:
            b = b + bn.bias
        return W.detach(), b.detach().contiguous()

    def _block(self, x, seq, act, snap=None, layer=0):
        """Run one `Linear -> Dropout -> [BatchNorm1d] -> [act]` Sequential through the kernel."""
        lin, bn = self._blocks(seq)[0]
        W, b = self._effective(lin, bn)
        return _LinearActFn.apply(x, W, b, act, PRECISIONS[self.precision],
                                  self._drop_spec(snap, layer))

    def _stack(self, x, seq, act, snap=None, layer0=0):
        for k, (lin, bn) in enumerate(self._blocks(seq)):
            W, b = self._effective(lin, bn)
            x = _LinearActFn.apply(x, W, b, act, PRECISIONS[self.precision],
                                   self._drop_spec(snap, layer0 + k))
        return x

    def _as_input(self, x):
        if not x.is_cuda:
            raise RuntimeError("abnet3_b200 networks run on an sm_100 GPU only: move the "
                               "input and the network to CUDA (there is no CPU path)")
        return x if x.dtype == torch.float32 else x.float()


def _layer(n_in, n_out, p_dropout, batch_norm, act_cls):
    mods = [nn.Linear(n_in, n_out), nn.Dropout(p=p_dropout)]
    if batch_norm:
        mods.append(nn.BatchNorm1d(n_out))
    if act_cls is not None:
        mods.append(act_cls())
    return mods


class SiameseNetwork(NetworkBuilder):
    """abnet3/model.py:82-208.  Same parameters; ``precision`` selects the kernel path of
    the layers: 'bf16' (default) = tcgen05 tensor cores with fp32 accumulation and fp32
    master weights, 'fp32' = the SIMT kernels (the 1e-4 parity path)."""

    def __init__(self, input_dim=None, num_hidden_layers=None, hidden_dim=None,
                 output_dim=None, p_dropout=0.1, batch_norm=False,
                 type_init='xavier_uni', activation_layer=None,
                 output_path=None, last_non_linearity="default", precision="bf16"):
        super(SiameseNetwork, self).__init__()
        assert activation_layer in ('relu', 'sigmoid', 'tanh')
        assert type_init in ('xavier_uni', 'xavier_normal', 'orthogonal')
        assert type(input_dim) == int, 'input dim should be int'
        assert type(hidden_dim) == int, 'hidden dim should be int'
        assert type(num_hidden_layers) == int, 'num hidden lay should be int'
        assert type(output_dim) == int, 'output dim should be int'
        assert precision in PRECISIONS

        self.input_dim = input_dim
        self.num_hidden_layers = num_hidden_layers
        self.hidden_dim = hidden_dim
        self.output_dim = output_dim
        self.activation_layer = activation_layer
        self.batch_norm = batch_norm
        self.type_init = type_init
        self.last_non_linearity = last_non_linearity
        self.p_dropout = p_dropout
        self.precision = precision

        activation = activation_functions[activation_layer]
        self.input_emb = nn.Sequential(
            *_layer(input_dim, hidden_dim, p_dropout, batch_norm, activation))
        hidden = []
        for _ in range(num_hidden_layers):
            hidden += _layer(hidden_dim, hidden_dim, p_dropout, batch_norm, activation)
        self.hidden_layers = nn.Sequential(*hidden)
        if last_non_linearity == "default":
            last = activation
        elif last_non_linearity is None:
            last = None
        else:
            last = activation_functions[last_non_linearity]
        self.output_layer = nn.Sequential(
            *_layer(hidden_dim, output_dim, p_dropout, batch_norm, last))
        self.output_path = output_path
        self.apply(self.init_weight_method)

    def _last_act(self):
        if self.last_non_linearity == "default":
            return self.activation_layer
        if self.last_non_linearity is None:
            return "none"
        if self.last_non_linearity == "softmax":
            raise NotImplementedError("softmax output is not implemented by the kernels")
        return self.last_non_linearity

    def forward_once(self, x):
        """abnet3/model.py:179-186"""
        self._check_supported()
        h = self._as_input(x)
        snap = self._drop_for_call(h.device)
        h = self._block(h, self.input_emb, self.activation_layer, snap, 0)
        h = self._stack(h, self.hidden_layers, self.activation_layer, snap, 1)
        return self._block(h, self.output_layer, self._last_act(), snap, 1 + self.num_hidden_layers)

    def forward(self, input1, input2):
        """abnet3/model.py:188-196: shared weights on both inputs (one 2B batch)."""
        n = input1.shape[0]
        out = self.forward_once(_joint(self._as_input(input1), self._as_input(input2)))
        return out[:n], out[n:]

    def inference_layers(self):
        """-> (trunk [(W, b, act)], heads []) with eval-mode BatchNorm folded in."""
        out = []
        for seq, act in ((self.input_emb, self.activation_layer),
                         (self.hidden_layers, self.activation_layer),
                         (self.output_layer, self._last_act())):
            for lin, bn in self._blocks(seq):
                W, b = self._effective(lin, bn)
                out.append((W.detach(), b.detach(), act))
        return out, []

    def layer_specs(self):
        """[(weight, bias, act)] in forward order, for the fused training engine."""
        specs = [(self.input_emb[0].weight, self.input_emb[0].bias, self.activation_layer)]
        for m in self.hidden_layers:
            if isinstance(m, nn.Linear):
                specs.append((m.weight, m.bias, self.activation_layer))
        specs.append((self.output_layer[0].weight, self.output_layer[0].bias, self._last_act()))
        return specs

    def save_network(self, epoch=''):
        torch.save(self.state_dict(), self.output_path + str(epoch) + '.pth')

    def load_network(self, network_path=None):
        self.load_state_dict(torch.load(network_path))


class SiameseMultitaskNetwork(NetworkBuilder):
    """abnet3/model.py:211-376: shared trunk, speaker and phone heads.  The
    ``hidden_layers_spk`` / ``hidden_layers_phn`` stacks are built (they are in
    the reference's state_dict, :293-309) and never applied (:346-354)."""

    def __init__(self, input_dim=None, num_hidden_layers_shared=None,
                 num_hidden_layers_spk=None,
                 num_hidden_layers_phn=None,
                 hidden_dim=None,
                 output_dim=None, p_dropout=0.1, batch_norm=False,
                 type_init='xavier_uni', activation_layer=None,
                 output_path=None, precision="bf16"):
        super(SiameseMultitaskNetwork, self).__init__()
        assert activation_layer in ('relu', 'sigmoid', 'tanh')
        assert type_init in ('xavier_uni', 'xavier_normal', 'orthogonal')
        assert type(input_dim) == int, 'input dim should be int'
        assert type(hidden_dim) == int, 'hidden dim should be int'
        assert type(num_hidden_layers_shared) == int
        assert type(num_hidden_layers_spk) == int
        assert type(num_hidden_layers_phn) == int
        assert type(output_dim) == int, 'output dim should be int'
        assert precision in PRECISIONS

        self.input_dim = input_dim
        self.num_hidden_layers_shared = num_hidden_layers_shared
        self.num_hidden_layers_spk = num_hidden_layers_spk
        self.num_hidden_layers_phn = num_hidden_layers_phn
        self.hidden_dim = hidden_dim
        self.output_dim = output_dim
        self.activation_layer = activation_layer
        self.batch_norm = batch_norm
        self.type_init = type_init
        self.p_dropout = p_dropout
        self.precision = precision

        activation = activation_functions[activation_layer]
        self.input_emb = nn.Sequential(
            *_layer(input_dim, hidden_dim, p_dropout, batch_norm, activation))

        def stack(n):
            mods = []
            for _ in range(n):
                mods += _layer(hidden_dim, hidden_dim, p_dropout, batch_norm, activation)
            return nn.Sequential(*mods)

        self.hidden_layers_shared = stack(num_hidden_layers_shared)
        self.hidden_layers_spk = stack(num_hidden_layers_spk)
        self.hidden_layers_phn = stack(num_hidden_layers_phn)
        self.output_layer_spk = nn.Sequential(
            *_layer(hidden_dim, output_dim, p_dropout, batch_norm, activation))
        self.output_layer_phn = nn.Sequential(
            *_layer(hidden_dim, output_dim, p_dropout, batch_norm, activation))
        self.output_path = output_path
        self.apply(self.init_weight_method)

    def forward_once(self, x):
        """abnet3/model.py:346-354"""
        self._check_supported()
        h = self._as_input(x)
        snap = self._drop_for_call(h.device)
        nt = 1 + self.num_hidden_layers_shared
        h = self._block(h, self.input_emb, self.activation_layer, snap, 0)
        h = self._stack(h, self.hidden_layers_shared, self.activation_layer, snap, 1)
        output_spk = self._block(h, self.output_layer_spk, self.activation_layer, snap, nt)
        output_phn = self._block(h, self.output_layer_phn, self.activation_layer, snap, nt + 1)
        return output_spk, output_phn

    def inference_layers(self):
        """-> (trunk [(W, b, act)], heads [[(W, b, act)], [(W, b, act)]]) with eval-mode
        BatchNorm folded in."""
        act = self.activation_layer
        trunk = []
        for seq in (self.input_emb, self.hidden_layers_shared):
            for lin, bn in self._blocks(seq):
                W, b = self._effective(lin, bn)
                trunk.append((W.detach(), b.detach(), act))
        heads = []
        for seq in (self.output_layer_spk, self.output_layer_phn):
            lin, bn = self._blocks(seq)[0]
            W, b = self._effective(lin, bn)
            heads.append([(W.detach(), b.detach(), act)])
        return trunk, heads

    def forward(self, input1, input2):
        """abnet3/model.py:356-364: returns (spk1, phn1, spk2, phn2)."""
        n = input1.shape[0]
        spk, phn = self.forward_once(_joint(self._as_input(input1), self._as_input(input2)))
        return spk[:n], phn[:n], spk[n:], phn[n:]

    def save_network(self, epoch=''):
        torch.save(self.state_dict(), self.output_path + str(epoch) + '.pth')

    def load_network(self, network_path=None):
        self.load_state_dict(torch.load(network_path))
