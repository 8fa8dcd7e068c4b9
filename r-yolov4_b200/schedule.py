"""Learning-rate / gradient-accumulation schedule of the reference training loop (train.py:150-163,189-202,219-220),
as host-side state that drives TrainStep:

    nbs = 64; accumulate = max(round(nbs / batch_size), 1)                         train.py:150-151
    nw  = max(int(epochs * iters_per_epoch * warmup_prop), 1000)                   train.py:160
    lf  = one_cycle(1, lrf, epochs)  (LambdaLR, stepped once per epoch)            train.py:36-38,161-162,220
    warm-up (global_step <= nw): accumulate and lr are re-interpolated every batch train.py:189-193
    optimizer.step() + zero_grad() when global_step % accumulate == 0              train.py:200-202
"""
import math

import numpy as np


def one_cycle(y1=0.0, y2=1.0, steps=100):
    """Sinusoidal ramp from y1 to y2 over `steps` (train.py:36-38)."""
    return lambda x: ((1 - math.cos(x * math.pi / steps)) / 2) * (y2 - y1) + y1


class Schedule:
    def __init__(self, epochs, iters_per_epoch, batch_size, lr, lrf=0.1, warmup_prop=0.05, nbs=64):
        self.epochs, self.ipe, self.bs, self.nbs = int(epochs), int(iters_per_epoch), batch_size, nbs
        self.initial_lr = lr
        self.accumulate = max(round(nbs / batch_size), 1)
        self.nw = max(int((epochs * iters_per_epoch) * warmup_prop), 1000)
        self.lf = one_cycle(1, lrf, int(epochs))
        self.lr = lr * self.lf(0)                  # LambdaLR sets lr = initial_lr * lf(0) at construction
        self._sched_epoch = 0

    def global_step(self, epoch, batch):
        return self.ipe * epoch + batch + 1        # train.py:184

    def batch(self, epoch, batch):
        """Call before each batch's forward: returns (lr, accumulate, do_step) for it."""
        gs = self.global_step(epoch, batch)
        if gs <= self.nw:
            xi = [0, self.nw]
            self.accumulate = max(1, np.interp(gs, xi, [1, self.nbs / self.bs]).round())
            self.lr = float(np.interp(gs, xi, [0.0, self.initial_lr * self.lf(epoch)]))
        return self.lr, self.accumulate, gs % self.accumulate == 0

    def epoch_end(self):
        """scheduler.step() (train.py:220): lr = initial_lr * lf(epochs completed)."""
        self._sched_epoch += 1
        self.lr = self.initial_lr * self.lf(self._sched_epoch)
        return self.lr
