"""Pair losses of /root/reference/abnet3/loss.py on the fused sm_100a kernel.

``coscos2`` (:37-67), ``cosmargin`` (:70-105) and ``weighted_loss_multi``
(:140-182) keep the reference's constructor arguments and ``forward``
signatures and return a scalar tensor that supports ``.backward()``.  One
kernel launch (abn_pair_loss) produces the loss value AND both embedding
gradients; autograd only scales them by the incoming gradient.
"""
import torch
import torch.nn as nn

from . import ops


class _PairLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e1, e2, y, kind, margin, scale):
        e1c = e1.contiguous().float()
        e2c = e2.contiguous().float()
        yf = y.to(device=e1c.device, dtype=torch.float32).contiguous()
        need = e1.requires_grad or e2.requires_grad
        loss, de1, de2 = ops.pair_loss(e1c, e2c, yf, kind, margin, scale, need_grad=need)
        ctx.save_for_backward(de1, de2) if need else None
        ctx.need = need
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        if not ctx.need:
            return None, None, None, None, None, None
        de1, de2 = ctx.saved_tensors
        return de1 * grad_out, de2 * grad_out, None, None, None, None


class LossBuilder(nn.Module):
    """Generic loss class (abnet3/loss.py:15-34)."""

    def __init__(self, *args, **kwargs):
        super(LossBuilder, self).__init__(*args, **kwargs)

    def forward(self, *args, **kwargs):
        raise NotImplementedError('Unimplemented forward for class:',
                                  self.__class__.__name__)

    def whoami(self, *args, **kwargs):
        return {'params': self.__dict__, 'class_name': self.__class__.__name__}


class coscos2(LossBuilder):
    """same pairs: (1 - cos) / 2; different pairs: cos^2; summed, divided by
    the batch size when ``avg`` (abnet3/loss.py:46-67)."""

    def __init__(self, avg=True, *args, **kwargs):
        super(coscos2, self).__init__(*args, **kwargs)
        self.avg = avg

    def forward(self, input1, input2, y):
        assert input1.size() == input2.size(), 'Input not the same size'
        scale = 1.0 / input1.size()[0] if self.avg else 1.0
        return _PairLossFn.apply(input1, input2, y, "coscos2", 0.0, scale)


class cosmargin(LossBuilder):
    """same pairs: 1 - cos; different pairs: max(cos - margin, 0)
    (abnet3/loss.py:85-105)."""

    def __init__(self, avg=True, margin=0.5, *args, **kwargs):
        super(cosmargin, self).__init__(*args, **kwargs)
        self.margin = margin
        self.avg = avg
        assert (margin >= 0 and margin <= 1)

    def forward(self, input1, input2, y, avg=True):
        assert input1.size() == input2.size(), 'Input not the same size'
        scale = 1.0 / input1.size()[0] if self.avg else 1.0
        return _PairLossFn.apply(input1, input2, y, "cosmargin", float(self.margin), scale)


class weighted_loss_multi(LossBuilder):
    """weight * loss_spk + (1 - weight) * loss_phn (abnet3/loss.py:165-182);
    argument order (spk1, phn1, spk2, phn2, y_spk, y_phn) as the reference."""

    def __init__(self, avg=True, loss_phn=None, loss_spk=None,
                 weight=0.5, *args, **kwargs):
        super(weighted_loss_multi, self).__init__(*args, **kwargs)
        assert type(weight) is float
        assert (weight >= 0 and weight <= 1)
        self.weight = weight
        self.avg = avg
        self.loss_phn = loss_phn
        self.loss_spk = loss_spk

    def forward(self, emb_spk1, emb_phn1, emb_spk2, emb_phn2, y_spk, y_phn):
        output_spk = self.loss_spk(emb_spk1, emb_spk2, y_spk)
        output_phn = self.loss_phn(emb_phn1, emb_phn2, y_phn)
        return self.weight * output_spk + (1.0 - self.weight) * output_phn
