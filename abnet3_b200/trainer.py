"""Training loop of /root/reference/abnet3/trainer.py for the siamese hot path.

``TrainerSiamese`` (:211-256) and ``TrainerSiameseMultitask`` (:259-279) keep
the reference's constructor arguments and control flow -- epoch-0 evaluation
pass, per-epoch train sweep + dev sweep, early stopping on the SUMMED dev loss
(:154-171, :256), save-best checkpoints with the reference's file names -- and
run every step through the fused kernels (abnet3_b200.engine).

New: data parallelism.  Under ``torchrun`` (one process per GPU,
``torch.distributed`` initialised with NCCL) every rank starts from rank 0's
weights, aligns and trains on its own shard of the pair lists
(``dataloader.shard(rank, world)``, applied by the trainer) and the flat gradient
bucket is exchanged once per step; ranks agree ONCE per epoch on the number of
steps (the minimum over ranks) so the exchange never deadlocks.

With a ``FramesDataLoader`` and the bf16 engine (the default) a sweep is
``engine.sweep_table``: the batch position lives on the device and every step is
one replay of a CUDA graph (gather -> forward -> loss -> backward -> exchange ->
optimizer); nothing is copied or synchronised per batch.

The loss is accumulated on the device and read back once per sweep (the
reference synchronises on ``loss.data[0]`` every step, :242).
"""
import copy
import pickle
import time
from pathlib import Path

import torch
import torch.distributed as dist
import torch.optim as optim

from .engine import SiameseTrainStep, OPTIMIZERS
from .loss import coscos2, cosmargin, weighted_loss_multi
from .model import NetworkBuilder, SiameseNetwork, SiameseMultitaskNetwork


def shard_pairs(pairs, rank, world_size):
    """Round-robin shard of a pair list (pairs are independent; DTW alignment
    needs no communication)."""
    return pairs[rank::world_size]


def all_ranks_have_batch(has_batch, device, world):
    """True while EVERY rank still has a batch: ranks must issue the same number
    of gradient all-reduces, so an epoch ends at the shortest shard."""
    if world <= 1:
        return bool(has_batch)
    more = torch.tensor([1 if has_batch else 0], device=device)
    dist.all_reduce(more, op=dist.ReduceOp.MIN)
    return int(more.item()) == 1


def agree_on_batches(n_batches, device, world):
    """The number of training steps of this epoch on EVERY rank: the minimum over the ranks'
    batch counts, ONE reduction per epoch (each step holds one gradient exchange, so ranks must
    run the same number; the reference already drops the tail, abnet3/dataloader.py:708)."""
    if world <= 1:
        return int(n_batches)
    t = torch.tensor([int(n_batches)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item())


def reduce_grads(parameters, avg, world):
    """torch.optim fallback under data parallelism: sum the gradients over the ranks (mean for
    a loss that averages over its batch), as the fused engine does inside its step."""
    if world <= 1:
        return
    for p in parameters:
        if p.grad is not None:
            dist.all_reduce(p.grad)
            if avg:
                p.grad.div_(world)


def _loss_spec(loss):
    if isinstance(loss, coscos2):
        return ("coscos2", 0.0, bool(loss.avg))
    if isinstance(loss, cosmargin):
        return ("cosmargin", float(loss.margin), bool(loss.avg))
    if isinstance(loss, weighted_loss_multi):
        return (_loss_spec(loss.loss_spk), _loss_spec(loss.loss_phn), float(loss.weight))
    return None


class TrainerBuilder:
    """abnet3/trainer.py:32-208"""

    def __init__(self, network=None, loss=None,
                 num_epochs=200, patience=20,
                 optimizer_type='sgd', lr=0.001, momentum=0.9, cuda=True,
                 seed=0, dataloader=None, log_dir=None,
                 feature_generator=None,
                 checkpoints=False):
        if not cuda or not torch.cuda.is_available():
            raise RuntimeError("abnet3_b200 trains on an sm_100 GPU only (cuda=True); "
                               "there is no CPU path")
        self.network = network
        self.loss = loss
        self.num_epochs = num_epochs
        self.patience = patience
        self.lr = lr
        self.momentum = momentum
        self.best_epoch = 0
        self.seed = seed
        self.cuda = cuda
        self.statistics_training = {}
        self.train_losses = []
        self.dev_losses = []
        self.dataloader = dataloader
        self.feature_generator = feature_generator
        self.checkpoints = checkpoints
        self.loss.cuda()
        self.network.cuda()
        if log_dir is None:
            self.log_dir = Path('./runs/%s' % time.strftime('%m-%d-%Hh%M-%S'))
        else:
            self.log_dir = Path(log_dir) / ('%s' % time.strftime('%m-%d-%Hh%M-%S'))
        assert optimizer_type in ('sgd', 'adadelta', 'adam', 'adagrad',
                                  'RMSprop', 'LBFGS')
        self.optimizer_type = optimizer_type
        self.rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        spec = _loss_spec(self.loss)
        self.engine = None
        if optimizer_type in OPTIMIZERS and spec is not None and isinstance(
                self.network, (SiameseNetwork, SiameseMultitaskNetwork)):
            self.engine = SiameseTrainStep(self.network, spec, optimizer_type, lr, momentum)
            self.optimizer = None
        else:
            # optimizers outside the fused kernel: torch.optim on the same parameters
            cls = {'sgd': optim.SGD, 'adadelta': optim.Adadelta, 'adam': optim.Adam,
                   'adagrad': optim.Adagrad, 'RMSprop': optim.RMSprop,
                   'LBFGS': optim.LBFGS}[optimizer_type]
            kw = {'lr': self.lr}
            if optimizer_type == 'sgd':
                kw['momentum'] = self.momentum or 0.0
            self.optimizer = cls(self.network.parameters(), **kw)
            if self.world > 1:          # same starting point on every rank
                for p in self.network.parameters():
                    dist.broadcast(p.data, src=0)
        if self.world > 1 and self.dataloader is not None and hasattr(self.dataloader, 'shard'):
            # every rank aligns and trains on its own part of the pair lists
            if getattr(self.dataloader, '_shard', None) is None:
                self.dataloader.shard(self.rank, self.world)

    def params(self):
        params = copy.copy(self.__dict__)
        for k in ('dataloader', 'feature_generator', 'engine', 'optimizer'):
            params.pop(k, None)
        return params

    def whoami(self):
        whoami = {
            'params': {k: v for k, v in self.params().items()
                       if isinstance(v, (int, float, str, bool, type(None)))},
            'network': {'class_name': self.network.__class__.__name__},
            'loss': {'class_name': self.loss.__class__.__name__},
            'class_name': self.__class__.__name__,
            'dataloader': self.dataloader.whoami() if self.dataloader is not None else None,
        }
        return whoami

    def save_whoami(self):
        pickle.dump(self.whoami(), open(self.network.output_path + '.params', "wb"))

    def optimize_model(self, do_training=True):
        raise NotImplementedError('Unimplemented optimize_model for class:',
                                  self.__class__.__name__)

    def _writers(self):
        try:
            from tensorboardX import SummaryWriter
        except ImportError:
            return None, None
        return (SummaryWriter(log_dir=str(self.log_dir / 'train_loss')),
                SummaryWriter(log_dir=str(self.log_dir / 'dev_loss')))

    def train(self):
        """abnet3/trainer.py:117-173"""
        self.patience_dev = 0
        self.best_dev = None
        self.train_losses = []
        self.dev_losses = []
        self.network.eval()
        main = self.rank == 0
        if main and self.network.output_path:
            self.network.save_network()
        train_writer, dev_writer = self._writers() if main else (None, None)

        self.optimize_model(do_training=False)              # epoch-0 evaluation pass
        if train_writer:
            train_writer.add_scalar('loss', self.train_losses[-1], 0)
            dev_writer.add_scalar('loss', self.dev_losses[-1], 0)
        if main and self.checkpoints and self.network.output_path:
            self.network.save_network(epoch=0)
        for key in self.statistics_training.keys():
            self.statistics_training[key] = 0

        for epoch in range(self.num_epochs):
            dev_loss = self.optimize_model(do_training=True)
            if train_writer:
                train_writer.add_scalar('loss', self.train_losses[-1], epoch + 1)
                dev_writer.add_scalar('loss', self.dev_losses[-1], epoch + 1)
            if self.best_dev is None or dev_loss < self.best_dev:
                self.best_dev = dev_loss
                self.patience_dev = 0
                if main and self.network.output_path:
                    print('Saving best model so far, epoch {}... '.format(epoch + 1), end='',
                          flush=True)
                    if self.checkpoints:
                        self.network.save_network(epoch=epoch + 1)
                    self.network.save_network()
                    self.save_whoami()
                    print("Done.")
                self.best_epoch = epoch
            else:
                self.patience_dev += 1
                if self.patience_dev > self.patience:
                    if main:
                        print("No improvements after {} iterations, "
                              "stopping now".format(self.patience))
                        print('Finished Training')
                    break
        if main:
            print('Saving best checkpoint network')

    def pretty_print_losses(self, train_loss, dev_loss):
        print("  training loss:\t\t{:.6f}".format(train_loss))
        print("  dev loss:\t\t\t{:.6f}".format(dev_loss))


class TrainerSiamese(TrainerBuilder):
    """abnet3/trainer.py:211-256"""
    N_LABELS = 1

    def __init__(self, *args, **kwargs):
        super(TrainerSiamese, self).__init__(*args, **kwargs)
        assert isinstance(self.network, NetworkBuilder)

    def give_batch_to_network(self, batch):
        """Forward + loss through the autograd surface (abnet3/trainer.py:211-224)."""
        X_batch1, X_batch2, y_batch = batch
        emb_batch1, emb_batch2 = self.network(X_batch1.cuda(), X_batch2.cuda())
        return self.loss(emb_batch1, emb_batch2, y_batch.cuda())

    def _joint_batch(self, batch):
        from .model import _joint
        x1, x2 = batch[0].cuda(), batch[1].cuda()
        labels = [t.cuda().float().contiguous() for t in batch[2:]]
        return _joint(x1.float(), x2.float()), x1.shape[0], labels

    def _reduce_grads(self):
        reduce_grads(self.network.parameters(), bool(getattr(self.loss, 'avg', False)), self.world)

    def _agreed_batches(self, n_batches):
        return agree_on_batches(n_batches, 'cuda', self.world)

    def _sweep_table(self, train_mode, do_training):
        """Sweep over a device-resident frame-pair table (FramesDataLoader.epoch_table):
        the engine gathers every batch itself and replays one CUDA graph per step."""
        import os
        trace = os.environ.get("ABN_TRACE_SWEEP") == "1"
        if trace:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        feat, table, bs, first_row, n_batches = self.dataloader.epoch_table(train_mode=train_mode)
        training = train_mode and do_training
        if training:
            n_batches = self._agreed_batches(n_batches)
        if trace:
            torch.cuda.synchronize()
            t1 = time.perf_counter()
        total = self.engine.sweep_table(feat, table, bs, n_batches, start=first_row,
                                        do_training=training)
        out = float(total.item())
        if trace:
            import sys
            t2 = time.perf_counter()
            sys.stderr.write("[sweep train=%d] epoch_table %.1f ms | %d batches %.1f ms\n"
                             % (train_mode, 1e3 * (t1 - t0), n_batches, 1e3 * (t2 - t1)))
        return out, n_batches

    def _sweep(self, train_mode, do_training):
        """One pass over the dataloader; returns (summed loss, number of batches)."""
        if (self.engine is not None and self.engine.precision == 1
                and hasattr(self.dataloader, 'epoch_table')):
            return self._sweep_table(train_mode, do_training)
        total = torch.zeros(1, dtype=torch.float32, device='cuda')
        n_batches = 0
        training = train_mode and do_training
        it = self.dataloader.batch_iterator(train_mode=train_mode)
        while True:
            batch = next(it, None)
            if self.world > 1 and training:
                # generator loaders do not know their length in advance
                if not all_ranks_have_batch(batch is not None, 'cuda', self.world):
                    break
            if batch is None:
                break
            if self.engine is not None:
                x, n, labels = self._joint_batch(batch)
                loss = self.engine.step(x, n, *labels, do_training=training)
                total += loss
            elif training and self.optimizer_type == 'LBFGS':
                def closure():
                    self.optimizer.zero_grad()
                    out = self.give_batch_to_network(batch)
                    out.backward()
                    self._reduce_grads()
                    return out
                total += self.optimizer.step(closure).detach()
            else:
                loss = self.give_batch_to_network(batch)
                self.optimizer.zero_grad()
                if training:
                    loss.backward()
                    self._reduce_grads()
                    self.optimizer.step()
                total += loss.detach()
            n_batches += 1
        return float(total.item()), n_batches

    def optimize_model(self, do_training=True):
        """abnet3/trainer.py:226-256"""
        self.network.train()
        train_loss, num_batches_train = self._sweep(True, do_training)
        self.network.eval()
        dev_loss, num_batches_dev = self._sweep(False, False)
        self.last_sweep = {'train_batches': num_batches_train, 'dev_batches': num_batches_dev}
        if self.world > 1:
            t = torch.tensor([train_loss, num_batches_train, dev_loss, num_batches_dev],
                             dtype=torch.float64, device='cuda')
            dist.all_reduce(t)
            train_loss, num_batches_train, dev_loss, num_batches_dev = t.tolist()
        self.train_losses.append(train_loss / max(num_batches_train, 1))
        self.dev_losses.append(dev_loss / max(num_batches_dev, 1))
        if self.rank == 0:
            self.pretty_print_losses(self.train_losses[-1], self.dev_losses[-1])
        return dev_loss


class TrainerSiameseMultitask(TrainerSiamese):
    """abnet3/trainer.py:259-279"""
    N_LABELS = 2

    def __init__(self, *args, **kwargs):
        super(TrainerSiameseMultitask, self).__init__(*args, **kwargs)
        assert type(self.network) == SiameseMultitaskNetwork

    def give_batch_to_network(self, batch):
        X_batch1, X_batch2, y_spk_batch, y_phn_batch = batch
        emb = self.network(X_batch1.cuda(), X_batch2.cuda())
        emb_spk1, emb_phn1, emb_spk2, emb_phn2 = emb
        return self.loss(emb_spk1, emb_phn1, emb_spk2, emb_phn2,
                         y_spk_batch.cuda(), y_phn_batch.cuda())
