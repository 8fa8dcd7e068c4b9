"""One optimizer step of the reference's training loop (train.py:184-202) executed natively:

    imgs, targets -> Yolo.forward(training=True) -> fused loss value+gradient -> conv-stack backward
                  -> ONE NCCL all-reduce of the flat fp32 gradient buffer (data parallel) -> SGD(momentum, nesterov)

Nothing here goes through torch autograd or torch.optim; torch.distributed only carries the all-reduce.
The reference's optimizer is SGD(lr, momentum=0.937, nesterov=True) (train.py:156) and it has no SyncBN, so
BatchNorm statistics stay per rank.

Data parallel (world > 1): the flat gradient buffer is cut into a few contiguous BUCKETS.  The backward pass walks the
tape in reverse, so the tail of the buffer (neck, deep backbone: most of the bytes) is final long before the 800x800
layers are done; as soon as the last tape entry feeding a bucket has launched, its weight-gradient scratches are folded
in and its ncclAllReduce(AVG) is queued on a communication stream, overlapping the rest of the backward pass.
"""
import torch
import torch.distributed as dist

from . import ops
from .dist import allreduce_mean, plan_buckets
from .model import blocks


class TrainStep:
    def __init__(self, model, compute_loss, lr=0.01, momentum=0.937, nesterov=True, weight_decay=0.0,
                 process_group=None, optimizer="SGD", n_buckets=4):
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
        self.n_buckets = n_buckets
        self._plan = self._comm = None
        self._works = []
        self._graph = None
        self._lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=self.flat.device)
        if self.world > 1:   # identical replicas
            dist.broadcast(self.flat, 0, group=self.pg)

    def zero_grad(self):
        self.grad.zero_()

    # ---- bucketed, backward-overlapped gradient all-reduce ---------------------------------------------------
    def _tape_params(self, tape):
        """For every tape entry (forward order): ids of the parameters whose gradient it finalises."""
        out = []
        neck = self.model.neck
        for e in tape:
            ids = []
            if e[0] == "head":
                conv = e[1].conv[0]
                ids += [id(conv.weight), id(conv.bias)]
                if e[4] is not None:                                  # yolov7 implicit head (model/neck.py:201,208,215)
                    j = {id(getattr(neck, f"conv{4 + i}")): i for i in (1, 2, 3)}[id(e[1])]
                    ids += [id(getattr(neck, f"ia{j}").implicit), id(getattr(neck, f"im{j}").implicit)]
            elif e[0] == "conv":
                mod = e[1]
                ids += [id(mod.conv[0].weight), id(mod.conv[1].weight), id(mod.conv[1].bias)]
            elif e[0] == "repconv":
                mod = e[1]
                for seq in (mod.rbr_dense, mod.rbr_1x1):
                    ids += [id(seq[0].weight), id(seq[1].weight), id(seq[1].bias)]
            out.append(ids)
        return out

    def _make_plan(self, tape):
        model = self.model
        per_entry = self._tape_params(tape)
        n = len(tape)
        ready = {}                                                    # id(param) -> reverse tape index that finalises it
        for fi, ids in enumerate(per_entry):
            for i in ids:
                ready[i] = max(ready.get(i, -1), n - 1 - fi)
        params = list(model.parameters())
        spans = [model._flat_offsets[id(p)] for p in params]
        final = [ready.get(id(p), -1) for p in params]
        buckets = plan_buckets([(o, o + (c + 3) // 4 * 4) for o, c in spans], final, self.n_buckets)
        plan = {}
        for lo, hi, when, members in buckets:
            ids = {id(params[i]) for i in members}
            plan.setdefault(max(when, 0), []).append(dict(lo=lo, hi=hi, sub=model.wgrad_subtable(ids)))
        return plan

    def _flush(self, bucket, sink, reduce):
        """Queue one bucket on the communication stream: wait for its producers (weight-gradient GEMMs on the side
        stream, BatchNorm / bias gradients on the main stream), fold the K-major scratches into the flat OIHW gradient,
        all-reduce the slice."""
        main = torch.cuda.current_stream()
        ev_main, ev_side = torch.cuda.Event(), torch.cuda.Event()
        ev_main.record(main)
        ev_side.record(sink.side)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ev_main)
            self._comm.wait_event(ev_side)
            if bucket["sub"] is not None:
                self.model.unpack_wgrads_sub(bucket["sub"])
            if reduce:
                self._works.append(dist.all_reduce(self.grad[bucket["lo"]:bucket["hi"]], op=dist.ReduceOp.AVG,
                                                   group=self.pg, async_op=True))
        done = torch.cuda.Event()
        done.record(self._comm)
        self._works.append(done)

    def forward_backward(self, imgs, targets, reduce=True):
        """Accumulates d loss / d params into the flat gradient buffer; returns the device loss items.
        world > 1: with `reduce` the gradient buckets are all-reduced (averaged) as the backward pass releases them;
        pass reduce=False for the non-stepping micro-batches of a gradient-accumulation window (train.py:200)."""
        levels = self.model(imgs, training=True)
        items, dlevels = self.crit.value_and_grad(levels, targets)
        if self.world == 1 or self.n_buckets <= 1 or not reduce or dist.get_backend(self.pg) != "nccl":
            self.model._bucketed_unpack = False
            self.model.backward(dlevels)
            self._reduced = False
            return items
        if self._plan is None:
            self._plan = self._make_plan(self.model.last_ctx.tape)
            self._comm = torch.cuda.Stream(self.flat.device)
        plan = self._plan

        def on_entry(i, sink):
            for b in plan.get(i, ()):
                self._flush(b, sink, True)

        self.model._bucketed_unpack = True
        try:
            from .model.backward import run_backward
            run_backward(self.model, self.model.last_ctx, list(dlevels), self.model._grad_views, on_entry=on_entry)
        finally:
            self.model._bucketed_unpack = False
        main = torch.cuda.current_stream()
        for w in self._works:                       # join: every bucket folded and reduced before the optimizer reads it
            if isinstance(w, torch.cuda.Event):
                main.wait_event(w)
            else:
                w.wait()
        self._works = []
        self._reduced = True
        return items

    def step(self):
        if self.world > 1 and not getattr(self, "_reduced", False):
            allreduce_mean(self.grad, self.pg)
        self._reduced = False
        self.nstep += 1
        if self.optimizer == "Adam":                 # torch.optim.Adam(lr) defaults (train.py:154)
            if self.first:
                self.buf2 = torch.zeros_like(self.flat)
            ops.adam_step(self.flat, self.grad, self.buf, self.buf2, self.lr, self.nstep, weight_decay=self.wd)
        elif self._capturing:
            ops.sgd_step_lrdev(self.flat, self.grad, self.buf, self._lr_dev, self.momentum, self.wd, self.nesterov)
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

    # ---- whole-step CUDA graph -------------------------------------------------------------------------------
    _capturing = False

    def capture(self, imgs, targets, target_capacity=None, warmup=2):
        """Capture ONE WHOLE optimizer step — zero_grad, train-mode forward, fused loss value + gradient, conv-stack
        backward (side-stream weight-gradient GEMMs included), SGD, bf16 weight repack — into a single CUDA graph
        (SURVEY.md §7 step 5): `replay()` is then one cudaGraphLaunch per step with no host synchronisation; the ~580
        ctypes launches, tensor-map encodes and allocator calls of the eager path happen once, here.

        imgs: [B,3,S,S] example batch (shape is frozen).  targets: example label rows; the graph is captured for
        `target_capacity` rows (default: 2x the example): replay() pads shorter label sets with image index -1, which
        the assignment kernel drops (csrc/assign.cuh), so a variable number of boxes per batch needs no re-capture.
        The learning rate lives in a device float (`set_lr`).  SGD only (Adam's bias correction is host arithmetic).
        Single-GPU path; with world > 1 use the eager step (the bucketed all-reduce is host-driven)."""
        if self.optimizer != "SGD":
            raise NotImplementedError("capture(): SGD only")
        if self.world > 1:
            raise NotImplementedError("capture(): single-GPU path (the bucketed all-reduce is host-driven)")
        cols = targets.shape[1]
        cap = int(target_capacity or max(64, 2 * targets.shape[0]))
        assert targets.shape[0] <= cap
        dev = self.flat.device
        self._g_imgs = torch.empty_like(imgs)
        self._g_tg = torch.full((cap, cols), -1.0, dtype=torch.float32, device=dev)
        self._g_cap = cap
        self._stage(imgs, targets)
        self.crit.sync_items = False
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # eager warm-up on a side stream (workspaces, tables, momentum)
            for _ in range(max(1, warmup)):
                self(self._g_imgs, self._g_tg)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.set_lr(self.lr)
        g = torch.cuda.CUDAGraph()
        self._capturing = True
        try:
            with torch.cuda.graph(g):
                self._g_items = self(self._g_imgs, self._g_tg)
        finally:
            self._capturing = False
        self._graph = g
        return self

    def set_lr(self, lr):
        self.lr = float(lr)
        self._lr_dev.fill_(self.lr)

    def _stage(self, imgs, targets):
        self._g_imgs.copy_(imgs, non_blocking=True)
        n = targets.shape[0]
        if n > self._g_cap:
            raise ValueError(f"{n} label rows exceed the captured capacity {self._g_cap}: capture() again")
        self._g_tg[:n].copy_(targets, non_blocking=True)
        if n < self._g_cap:
            self._g_tg[n:, 0].fill_(-1.0)

    def replay(self, imgs, targets):
        """One captured step on new data: two device copies into the graph's input buffers + one graph launch.
        Returns the device loss items (valid until the next replay)."""
        assert self._graph is not None, "capture() first"
        if self.model._pack_signature() != self.model._pack_sig:     # load_state_dict / init since the last step
            self.model.repack_weights()
        self._stage(imgs, targets)
        self._graph.replay()
        self.nstep += 1
        blocks.WEIGHT_EPOCH[0] += 1          # weights and running statistics moved without any torch-visible write
        blocks.BN_EPOCH[0] += 1
        self.model._pack_sig = self.model._pack_signature()          # the graph's last node repacked the operands
        return self._g_items

    def train_batch(self, imgs, targets, schedule, epoch, batch):
        """One iteration of the reference's inner loop (train.py:183-202) under a `schedule.Schedule`: warm-up lr /
        accumulate interpolation, loss.backward() accumulating into the flat gradient buffer, and optimizer.step() +
        zero_grad() only when global_step % accumulate == 0.  Returns (loss items, stepped)."""
        self.lr, _, do_step = schedule.batch(epoch, batch)
        items = self.forward_backward(imgs, targets, reduce=do_step)
        if do_step:
            self.step()
            self.zero_grad()
        return items, do_step
