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

Not supported by the kernels (they raise instead of silently falling back):
``batch_norm=True`` and dropout with p > 0 in training mode.
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
    def forward(ctx, x, weight, bias, act, precision):
        x = x.contiguous()
        if precision == 1:
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
                                         need_dx=ctx.needs_input_grad[0])
        return dx, dW, db, None, None


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
    def _check_supported(self, training=None):
        training = self.training if training is None else training
        if self.batch_norm:
            raise NotImplementedError(
                "batch_norm=True is not implemented by the sm_100a kernels (no fallback)")
        if training and self.p_dropout > 0:
            raise NotImplementedError(
                "dropout with p > 0 in training mode is not implemented by the sm_100a "
                "kernels (use p_dropout=0 as in test/data/buckeye.yaml, or eval mode)")

    def _block(self, x, seq, act):
        """Run one `Linear -> Dropout -> [act]` Sequential through the kernel."""
        lin = seq[0]
        return _LinearActFn.apply(x, lin.weight, lin.bias, act, PRECISIONS[self.precision])

    def _stack(self, x, seq, act):
        for m in seq:
            if isinstance(m, nn.Linear):
                x = _LinearActFn.apply(x, m.weight, m.bias, act, PRECISIONS[self.precision])
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
        h = self._block(h, self.input_emb, self.activation_layer)
        h = self._stack(h, self.hidden_layers, self.activation_layer)
        return self._block(h, self.output_layer, self._last_act())

    def forward(self, input1, input2):
        """abnet3/model.py:188-196: shared weights on both inputs (one 2B batch)."""
        n = input1.shape[0]
        out = self.forward_once(_joint(self._as_input(input1), self._as_input(input2)))
        return out[:n], out[n:]

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
        h = self._block(h, self.input_emb, self.activation_layer)
        h = self._stack(h, self.hidden_layers_shared, self.activation_layer)
        output_spk = self._block(h, self.output_layer_spk, self.activation_layer)
        output_phn = self._block(h, self.output_layer_phn, self.activation_layer)
        return output_spk, output_phn

    def forward(self, input1, input2):
        """abnet3/model.py:356-364: returns (spk1, phn1, spk2, phn2)."""
        n = input1.shape[0]
        spk, phn = self.forward_once(_joint(self._as_input(input1), self._as_input(input2)))
        return spk[:n], phn[:n], spk[n:], phn[n:]

    def save_network(self, epoch=''):
        torch.save(self.state_dict(), self.output_path + str(epoch) + '.pth')

    def load_network(self, network_path=None):
        self.load_state_dict(torch.load(network_path))
