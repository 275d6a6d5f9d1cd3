"""Fused training step over the flat parameter buffer, with batch-sharded data parallelism.

Semantics of the reference's step (LetsFaceItGlow.training_step + configure_optimizers,
lets_face_it_glow.py:39-72, final_model.yaml: Adam lr 1e-5 betas (0.9, 0.9999), gradient_clip_val 20):
loss = mean_t mean_b NLL bits; backward; clip_grad_norm_(20); Adam.  Here the whole step is a handful of
stream-ordered launches with no host synchronisation: forward + backward through the C ABI, one NCCL
all-reduce of the flat fp32 gradient when world_size > 1 (the only exchange the path has),
and a fused clip+Adam over the flat buffer.
"""
from __future__ import annotations

import math

import torch

from . import _cabi as cabi

LN2 = math.log(2.0)


class Trainer:
    def __init__(self, model, lr=None, betas=None, eps=None, max_norm=None, process_group=None, dropout=True, probe_seed=1234):
        hp = model.hparams
        self.model = model
        self.eng = model.engine()
        adam = hp.Optim["args"]["adam"]
        self.base_lr = float(lr if lr is not None else hp.lr)
        self.lr = self.base_lr
        self.epoch = 0
        # configure_optimizers (lets_face_it_glow.py:61-72) returns get_scheduler(Optim.Schedule) (utils.py:65-82): StepLR per epoch
        sched = (hp.Optim.get("Schedule") or {}) if isinstance(getattr(hp, "Optim", None), dict) else {}
        self._sched_name = sched.get("name") or None
        self._sched_args = (sched.get("args") or {}).get(self._sched_name, {}) if self._sched_name else {}
        if self._sched_name not in (None, "step"):
            raise NotImplementedError("Unimplemented Scheduler!")  # utils.py:80 (multiplicative / lambda are unused by the shipped yamls)
        self.betas = tuple(betas if betas is not None else adam["betas"])
        self.eps = float(eps if eps is not None else adam["eps"])
        self.max_norm = float(max_norm if max_norm is not None else (hp.gradient_clip_val or 0))
        self.dropout = dropout
        dev = self.eng.theta.device
        n = self.eng.n_theta
        self.m = torch.zeros(n, device=dev)
        self.v = torch.zeros(n, device=dev)
        self.gflat = self.eng.new_flat_grad()
        self.scratch = torch.zeros(2, device=dev)
        self.step_count = 0
        self.pg = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        # the probe decision of training_step must be the same on every rank: a generator of its own,
        # seeded identically everywhere, instead of the process-global `random` whose state differs between ranks
        import random
        self._probe_rng = random.Random(probe_seed) if self.world > 1 else random
        self._dnll = {}
        # LetsFaceItGlow.__init__ (lets_face_it_glow.py:28-36): state of the mismatched-NLL probe
        self.last_missmatched_nll = float("inf")
        self.missmatched_modalities, self.missmatched_nll_name = None, None
        if getattr(hp, "Train", None) and hp.Train.get("use_negative_nll_loss"):
            from .glow.utils import get_mismatched_modalities
            self.missmatched_modalities, self.missmatched_nll_name = get_mismatched_modalities(hp)
        # data parallel: the flow-step bucket of the gradient is reduced on a side stream while the encoder backward runs
        self.overlap = self.world > 1 and dev.type == "cuda"
        if self.overlap:
            self._side = torch.cuda.Stream(device=dev)
            self._ev = torch.cuda.Event()
            self._ev.record()  # creates the underlying cudaEvent
            self._bucket = flow_grad_range(self.eng.blocks)

    def broadcast_parameters(self, src=0):
        """Replicas start identical.  (The ActNorm data-dependent init is shared by `step` itself: it broadcasts rank 0's
        parameters right after the init, whatever the caller did before.)"""
        if self.world > 1:
            torch.distributed.broadcast(self.eng.theta, src, group=self.pg)

    def epoch_end(self):
        """StepLR(step_size, gamma) stepped once per epoch, as Lightning steps the scheduler `configure_optimizers` returns
        (lets_face_it_glow.py:61-72, utils.py:65-82; final_model.yaml: step_size 3, gamma 0.73)."""
        self.epoch += 1
        if self._sched_name == "step":
            self.lr = self.base_lr * float(self._sched_args.get("gamma", 0.1)) ** (self.epoch // int(self._sched_args["step_size"]))
        return self.lr

    def training_step(self, batch, masks=None):
        """`LetsFaceItGlow.training_step` (lets_face_it_glow.py:39-55) on the fused step: with `use_negative_nll_loss`, once
        the last mismatched NLL is positive, one step in ten (python `random`, as the reference) trains on a batch whose
        interlocutor modalities are shuffled across sequences (`derange_batch`) with the loss scaled by -0.1.  The `and`
        chain is evaluated in the reference's order, so `random.random()` is consumed on the same steps; the only host
        synchronisation is the read-back of the probe's NLL on those (10 %) steps.  Returns (loss, deranged)."""
        from .glow.utils import derange_batch

        hp = self.model.hparams
        if (hp.Train["use_negative_nll_loss"] and self.last_missmatched_nll > 0 and self._probe_rng.random() < 0.1
                and self.missmatched_modalities):
            loss = self.step(derange_batch(batch, self.missmatched_modalities), masks, loss_scale=-0.1)
            probe = loss.detach().clone()
            if self.world > 1:  # every rank must see the same `last_missmatched_nll`: mean over the global batch
                torch.distributed.all_reduce(probe, op=torch.distributed.ReduceOp.SUM, group=self.pg)
                probe /= self.world
            self.last_missmatched_nll = -float(probe)  # "Loss/missmatched_nll" (:51-52)
            return loss * -0.1, True
        return self.step(batch, masks), False

    def step(self, batch, masks=None, loss_scale=1.0):
        """One optimizer step on this rank's shard of sequences.  Returns the (device) loss of the shard (unscaled);
        `loss_scale` multiplies the loss the gradient is taken of (the -0.1 of the mismatched-NLL probe)."""
        self.step_count += 1
        return self._step_body(batch, masks, loss_scale, None)

    def adam_scalars(self):
        """(lr, 1 - beta1^step, sqrt(1 - beta2^step)) of the CURRENT step count, in double as torch.optim.Adam computes them."""
        b1, b2 = float(self.betas[0]), float(self.betas[1])
        return self.lr, 1.0 - b1 ** self.step_count, math.sqrt(1.0 - b2 ** self.step_count)

    def _step_body(self, batch, masks, loss_scale, hyper):
        """The launches of one step.  `hyper` (device, 3 floats = `adam_scalars()`): the per-step scalars are read on the device
        (`lfi_clip_adam_dev`), so that the identical launch sequence can be captured once and replayed (GraphedStep)."""
        model = self.model
        eng = self.eng = model.engine()  # picks up a changed model.gemm_mode / a model moved to another device
        x0 = batch["p1_face"]
        B, T = x0.shape[0], x0.shape[1]
        Tp = T - eng.start_ts
        if masks is None and self.dropout and model.training:
            masks = model._masks(Tp, B, eng.theta.device)
        if model.training and not all(l.actnorm.inited for l in model.glow.flow.layers):
            model._ddi(eng, batch, masks)
            if self.world > 1:  # the init is data dependent: every replica takes rank 0's (the other shards' statistics are dropped)
                torch.distributed.broadcast(eng.theta, 0, group=self.pg)
        z, nll = eng.train_forward(batch, masks)
        key = (Tp, B, float(loss_scale))
        if key not in self._dnll:
            self._dnll[key] = torch.full((Tp, B), float(loss_scale) / (Tp * B), device=eng.theta.device)
        self.gflat.zero_()
        g = self.gflat[:eng.n_theta]
        if self.overlap:
            L = cabi.lib()
            L.lfi_set_grad_ready_event(self._ev.cuda_event)
            try:
                eng.train_backward(z, self._dnll[key], self.gflat)
            finally:
                L.lfi_set_grad_ready_event(None)
            lo, hi = self._bucket
            with torch.cuda.stream(self._side):
                self._side.wait_event(self._ev)  # recorded inside the backward call, before the encoder backward
                work = torch.distributed.all_reduce(g[lo:hi], op=torch.distributed.ReduceOp.SUM, group=self.pg, async_op=True)
            # ActNorm / 1x1-conv and encoder gradients are final only now.  The flat layout puts them behind the flow-step bucket
            # as one contiguous tail (engine.py: _layout), so this is ONE small all-reduce (2.4 MB) instead of two.
            for a0, a1 in ((0, lo), (hi, eng.n_theta)):
                if a1 > a0:
                    torch.distributed.all_reduce(g[a0:a1], op=torch.distributed.ReduceOp.SUM, group=self.pg)
            work.wait()
        else:
            eng.train_backward(z, self._dnll[key], self.gflat)
            allreduce_flat_gradient(g, self.world, self.pg)  # sum; averaged by grad_scale below
        loss = nll.mean() - eng.logdet_const() / LN2  # before the update: the parameter-only log-det term belongs to THIS step's theta
        L = cabi.lib()
        if hyper is None:
            cabi.check(L.lfi_clip_adam(eng.theta.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), eng.n_theta,
                                       self.lr, self.betas[0], self.betas[1], self.eps, self.max_norm, 1.0 / self.world,
                                       self.step_count, self.scratch.data_ptr(), cabi.stream_ptr()), "lfi_clip_adam")
        else:
            cabi.check(L.lfi_clip_adam_dev(eng.theta.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), eng.n_theta,
                                           hyper.data_ptr(), self.betas[0], self.betas[1], self.eps, self.max_norm, 1.0 / self.world,
                                           self.scratch.data_ptr(), cabi.stream_ptr()), "lfi_clip_adam_dev")
        return loss

    def graphed(self, batch, masks=None, warmup=3):
        """A `GraphedStep` over batches of the shape of `batch`."""
        return GraphedStep(self, batch, masks, warmup)

    def grad_norm(self):
        """Global gradient norm of the last step (after the all-reduce average, before clipping)."""
        return torch.sqrt(self.scratch[0]) / self.world


class GraphedStep:
    """One training step (frame-dropout masks, forward + NLL, backward, clip, Adam: every launch of `Trainer.step`, side streams
    included) captured ONCE as a CUDA graph and replayed per step ("make the whole optimizer step
    graph-capturable").  What makes the step capturable: no host synchronisation anywhere in it, every buffer owned by the
    caller at a fixed address (grow-only workspace, flat theta / grad / Adam state), and the only scalars that change from
    step to step - learning rate and Adam's bias corrections - read from device memory (`lfi_clip_adam_dev`).  The frame
    dropout draws from torch's CUDA generator, which torch advances per replay.

    `step(batch)` copies the batch into the captured input buffers, uploads the three scalars (pinned, stream ordered) and
    replays; it returns the loss tensor of the captured step (overwritten by the next replay).  `warmup` eager steps run first
    (they ARE optimizer steps: ActNorm data-dependent init, workspace growth and the library's per-shape caches happen there).
    Single process only: with world_size > 1 the eager, stream-ordered `Trainer.step` is the path (its NCCL calls overlap the
    backward through events that are handed to the library per call)."""

    def __init__(self, trainer, batch, masks=None, warmup=3):
        if trainer.world > 1:
            raise NotImplementedError("GraphedStep: single-process only; use Trainer.step under torch.distributed")
        tr = self.tr = trainer
        dev = tr.eng.theta.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedStep needs a CUDA device (no CPU path)")
        self.static = {k: v.to(dev).float().contiguous().clone() for k, v in batch.items() if torch.is_tensor(v)}
        self.masks = masks
        self.hyper = torch.zeros(3, device=dev)
        self._host = [torch.zeros(3).pin_memory() for _ in range(4)]
        self._host_ev = [torch.cuda.Event() for _ in range(4)]
        self._i = 0
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):
                tr.step(self.static, masks)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = tr._step_body(self.static, masks, 1.0, self.hyper)

    def step(self, batch=None):
        tr = self.tr
        if batch is not None:
            for k, dst in self.static.items():
                src = batch[k]
                if src.data_ptr() != dst.data_ptr():
                    dst.copy_(src, non_blocking=True)
        tr.step_count += 1
        j = self._i
        self._i = (j + 1) % len(self._host)
        self._host_ev[j].synchronize()  # the copy that last used this pinned slot has been executed
        lr, bc1, bc2s = tr.adam_scalars()
        self._host[j][0], self._host[j][1], self._host[j][2] = lr, bc1, bc2s
        self.hyper.copy_(self._host[j], non_blocking=True)
        self._host_ev[j].record()
        self.graph.replay()
        return self.loss


class HostFeed:
    """End-to-end driver over host batches (what a DataLoader with pinned memory hands to `LetsFaceItGlow.training_step`):
    every step's inputs are copied host -> device on a copy stream into one of two device buffers while the previous step
    computes, and every step's loss is read back (device -> host) one step late, so that neither the copy nor the read-back
    leaves the GPU idle.  `step(host_batch)` returns the loss (python float) of the PREVIOUS call (None on the first);
    `flush()` returns the last one."""

    def __init__(self, trainer):
        self.tr = trainer
        dev = trainer.eng.theta.device
        self.dev = dev
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.bufs = [None, None]
        self.copied = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0
        self.pending = None  # (pinned host scalar, event) of the previous step's loss
        self._loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._loss_ev = [torch.cuda.Event(), torch.cuda.Event()]

    def step(self, host_batch, masks=None):
        j = self.i & 1
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            if self.i >= 2:
                self.copy_stream.wait_event(self.consumed[j])  # the step that last read this buffer is done
            if self.bufs[j] is None or any(self.bufs[j][k].shape != v.shape for k, v in host_batch.items()):
                self.bufs[j] = {k: torch.empty(v.shape, dtype=torch.float32, device=self.dev) for k, v in host_batch.items()}
            for k, v in host_batch.items():
                self.bufs[j][k].copy_(v, non_blocking=True)
            self.copied[j].record(self.copy_stream)
        cur.wait_event(self.copied[j])
        loss = self.tr.step(self.bufs[j], masks)
        self.consumed[j].record(cur)
        self._loss_host[j].copy_(loss.detach().reshape(1), non_blocking=True)
        self._loss_ev[j].record(cur)
        prev = self.flush() if self.pending is not None else None
        self.pending = j
        self.i += 1
        return prev

    def flush(self):
        """Loss of the most recent step whose read-back has not been returned yet."""
        if self.pending is None:
            return None
        j, self.pending = self.pending, None
        self._loss_ev[j].synchronize()
        return float(self._loss_host[j][0])


class ResidentFeed:
    """End-to-end driver over an HBM-resident corpus (`lets_face_it_b200.data.ResidentWindows`):
    per step the host sends only the window table entries of the batch (8 bytes per sequence instead of 55 KB), the batch is
    gathered on the device, and the loss is read back one step late as in `HostFeed`."""

    def __init__(self, trainer, windows):
        self.tr, self.ds = trainer, windows
        self.bufs = {}
        self.pending = None
        self._loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0

    def step(self, index, masks=None):
        j = self.i & 1
        batch = self.ds.batch(index, out=self.bufs)
        loss = self.tr.step(batch, masks)
        self._loss_host[j].copy_(loss.detach().reshape(1), non_blocking=True)
        self._loss_ev[j].record(torch.cuda.current_stream(self.ds.device))
        prev = self.flush() if self.pending is not None else None
        self.pending = j
        self.i += 1
        return prev

    def flush(self):
        if self.pending is None:
            return None
        j, self.pending = self.pending, None
        self._loss_ev[j].synchronize()
        return float(self._loss_host[j][0])


def allreduce_flat_gradient(g, world, group=None):
    """The one exchange of the path: SUM all-reduce of the flat fp32 gradient over the ranks.
    The mean over the global batch is restored by `grad_scale = 1/world` in the fused clip+Adam (every rank holds an
    equal shard and the reference loss is a batch mean, models.py:555), so clipping sees the global-batch gradient."""
    if world > 1:
        torch.distributed.all_reduce(g, op=torch.distributed.ReduceOp.SUM, group=group)
    return g


# flat-buffer blocks whose gradients are final before the encoder backward (engine.py: _STEP_BLOCKS_F, contiguous)
_FLOW_BUCKET = ("wc", "bc", "w_ih", "b_ih", "w_hh", "b_hh", "wf", "bf", "lf")


def flow_grad_range(blocks):
    """[lo, hi) of the flat gradient covered by the flow-step weight blocks (`blocks`: name -> (offset, numel per step,
    steps), Engine.blocks).  The blocks must be adjacent in the flat layout (only alignment padding between them)."""
    spans = sorted((blocks[n][0], blocks[n][0] + blocks[n][1] * blocks[n][2]) for n in _FLOW_BUCKET if n in blocks)
    if not spans:
        return 0, 0
    lo, hi = spans[0][0], spans[-1][1]
    inside = {n for n, (o, m, k) in blocks.items() if lo <= o < hi}
    if inside != {n for n in _FLOW_BUCKET if n in blocks}:
        raise RuntimeError("flat layout changed: blocks %s lie inside the flow-step bucket" % sorted(inside - set(_FLOW_BUCKET)))
    return lo, hi


def shard_batch(batch, rank, world):
    """Contiguous, equal shards of the sequence batch; B must divide evenly so the mean over
    the global batch equals the mean of the per-rank means."""
    out = {}
    for k, v in batch.items():
        B = v.shape[0]
        if B % world:
            raise ValueError("batch size %d is not divisible by world size %d" % (B, world))
        n = B // world
        out[k] = v[rank * n:(rank + 1) * n]
    return out
