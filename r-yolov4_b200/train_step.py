"""One optimizer step of the reference's training loop (train.py:184-202) executed natively:

    imgs, targets -> Yolo.forward(training=True) -> fused loss value+gradient -> conv-stack backward
                  -> ONE NCCL all-reduce of the flat fp32 gradient buffer (data parallel) -> SGD(momentum, nesterov)

Nothing here goes through torch autograd or torch.optim; torch.distributed only carries the all-reduce.
The reference's optimizer is SGD(lr, momentum=0.937, nesterov=True) (train.py:156) and it has no SyncBN, so
BatchNorm statistics stay per rank.
"""
import torch
import torch.distributed as dist

from . import ops
from .dist import allreduce_mean
from .model import blocks


class TrainStep:
    def __init__(self, model, compute_loss, lr=0.01, momentum=0.937, nesterov=True, weight_decay=0.0,
                 process_group=None, optimizer="SGD"):
        if optimizer not in ("SGD", "Adam"):
            raise NotImplementedError("The specified optimizer is not implemented.")      # train.py:158
        self.model, self.crit, self.optimizer = model, compute_loss, optimizer
        self.lr, self.momentum, self.nesterov, self.wd = lr, momentum, nesterov, weight_decay
        self.nstep = 0
        model.autograd = False
        self.flat, self.grad = model.flatten_parameters()
        self.buf = torch.zeros_like(self.flat)
        model.enable_fused_pack()
        self.first = True
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        if self.world > 1:   # identical replicas
            dist.broadcast(self.flat, 0, group=self.pg)

    def zero_grad(self):
        self.grad.zero_()

    def forward_backward(self, imgs, targets):
        """Accumulates d loss / d params into the flat gradient buffer; returns the device loss items."""
        levels = self.model(imgs, training=True)
        items, dlevels = self.crit.value_and_grad(levels, targets)
        self.model.backward(dlevels)
        return items

    def step(self):
        if self.world > 1:
            allreduce_mean(self.grad, self.pg)
        self.nstep += 1
        if self.optimizer == "Adam":                 # torch.optim.Adam(lr) defaults (train.py:154)
            if self.first:
                self.buf2 = torch.zeros_like(self.flat)
            ops.adam_step(self.flat, self.grad, self.buf, self.buf2, self.lr, self.nstep, weight_decay=self.wd)
        else:
            ops.sgd_step(self.flat, self.grad, self.buf, self.lr, self.momentum, self.wd, self.nesterov, self.first)
        self.first = False
        blocks.WEIGHT_EPOCH[0] += 1          # BN-eval affine caches are stale now
        self.model.repack_weights()          # refresh every bf16 operand copy in one launch (re-keys the staleness check)

    def __call__(self, imgs, targets):
        self.zero_grad()
        items = self.forward_backward(imgs, targets)
        self.step()
        return items

    def train_batch(self, imgs, targets, schedule, epoch, batch):
        """One iteration of the reference's inner loop (train.py:183-202) under a `schedule.Schedule`: warm-up lr /
        accumulate interpolation, loss.backward() accumulating into the flat gradient buffer, and optimizer.step() +
        zero_grad() only when global_step % accumulate == 0.  Returns (loss items, stepped)."""
        self.lr, _, do_step = schedule.batch(epoch, batch)
        items = self.forward_backward(imgs, targets)
        if do_step:
            self.step()
            self.zero_grad()
        return items, do_step
