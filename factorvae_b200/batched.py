"""Date-batched, date-sharded ELBO step -- the extension that sits next to the reference's per-date
loop (train_model.py:17-30; the reference has no date batch and no data parallelism).

Semantics: for a global batch of B dates,
    L = (1/B) sum_d [ mean_i (yhat_di - y_di)^2 + KL_d ],
i.e. B=1 is the reference step and grad L is the mean of B reference per-date gradients.

Multi-GPU: dates are independent (every cross-sectional op is within one date), so each rank takes a
contiguous block of dates, runs the same kernel chain on its block and the ranks exchange exactly ONE
buffer per step: the flat fp32 gradient with the loss riding in its tail (<= 2.2 MB, latency bound on
NVLink/NVSwitch).  Noise is keyed by the GLOBAL unit index, so results do not depend on the sharding.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import engine


def shard_dates(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [d0, d1) of the B global dates owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(B, world)
    d0 = rank * base + min(rank, rem)
    return d0, d0 + base + (1 if rank < rem else 0)


class DateShardedStep:
    def __init__(self, layout: engine.ParamLayout, flat: torch.Tensor, precision: str = "bf16", group=None,
                 seed: int = 42, collective: str = "auto"):
        """collective: "p2p" = the one-kernel all-reduce over NVLink peer memory (p2p.P2PAllReduce), "nccl" = one
        torch.distributed all-reduce, "auto" = p2p when the group runs NCCL on GPUs with peer access, else nccl."""
        self.layout, self.flat, self.precision, self.group, self.seed = layout, flat, precision, group, seed
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.p2p = None
        if self.world > 1 and collective in ("auto", "p2p") and flat.is_cuda and dist.get_backend(group) == "nccl":
            try:
                from .p2p import P2PAllReduce
                self.p2p = P2PAllReduce(layout.total + 4, flat.device, group)
            except Exception:
                if collective == "p2p":
                    raise
                self.p2p = None
            # all ranks must agree (a rank without peer access would wait for an NCCL call the others never issue)
            ok = torch.tensor([1 if self.p2p is not None else 0], device=flat.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                if self.p2p is not None:
                    self.p2p.close()
                self.p2p = None
        # gradient buffer with a 4-float tail: [total] = loss (written by the kernels), rest padding
        self.gradbuf = torch.zeros(layout.total + 4, dtype=torch.float32, device=flat.device)
        self.workspace: Optional[torch.Tensor] = None
        self._outs: Dict = {}           # output tensors of the step, allocated once per batch shape and reused
        self.step_index = 0
        self._dev: Dict[str, torch.Tensor] = {}

    @property
    def grad(self) -> torch.Tensor:
        return self.gradbuf[: self.layout.total]

    @property
    def loss(self) -> torch.Tensor:
        return self.gradbuf[self.layout.total: self.layout.total + 1]

    def _local(self, x, y, date_ptr, *, unit_base=0, train=True, eps=None, keep_mask=None, step_dev=None):
        """Forward + backward over the given dates into self.gradbuf (gradient of the MEAN over these dates, loss in the
        tail); no collective.  step_dev: a device word that holds the Philox step counter (graph-captured steps)."""
        step = self.step_index if step_dev is None else step_dev
        noise = dict(eps=eps, keep_mask=keep_mask) if eps is not None else dict(philox=(self.seed, step, unit_base))
        out, st = engine.elbo_forward(self.layout, self.flat, x, y, date_ptr, train=train, precision=self.precision,
                                      workspace=self.workspace, loss_out=self.loss, out_cache=self._outs, **noise)
        self.workspace = st.workspace
        engine.elbo_backward(self.layout, st, grad=self.grad)
        return out, st

    def _reduce(self, buf: torch.Tensor, local_weight: float) -> None:
        """The single gradient exchange of the step: buf <- sum over ranks of local_weight * buf (loss in the tail)."""
        if self.world > 1 and self.p2p is not None:
            self.p2p.all_reduce(buf, local_weight)
        elif self.world > 1:
            if abs(local_weight * self.world - 1.0) < 1e-12 and dist.get_backend(self.group) == "nccl":
                dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)
            else:
                if local_weight != 1.0:
                    buf.mul_(local_weight)
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        elif local_weight != 1.0:
            buf.mul_(local_weight)

    def step(self, x: torch.Tensor, y: torch.Tensor, date_ptr: torch.Tensor, *, global_dates: Optional[int] = None,
             unit_base: int = 0, train: bool = True, eps: Optional[torch.Tensor] = None,
             keep_mask: Optional[torch.Tensor] = None):
        """Forward + backward over this rank's dates, then the single gradient all-reduce.

        x (S_local, T, C), y (S_local,), date_ptr (B_local+1,) local CSR.  global_dates = B of the whole
        batch (default: B_local * world).  Returns (outputs, state); self.grad / self.loss hold the
        GLOBAL-batch gradient / loss afterwards."""
        self.step_index += 1
        B_local = date_ptr.numel() - 1
        B_global = global_dates if global_dates is not None else B_local * self.world
        out, st = self._local(x, y, date_ptr, unit_base=unit_base, train=train, eps=eps, keep_mask=keep_mask)
        self._reduce(self.gradbuf, float(B_local) / float(B_global))
        return out, st

    def step_accumulate(self, micro_batches, *, global_dates: int, train: bool = True):
        """One step over this rank's dates processed as several micro-batches (a per-GPU share too large for one workspace:
        BASELINE.json configs[3..4] on few GPUs).  micro_batches: iterable of (x, y, date_ptr, unit_base).  Gradients are
        accumulated locally with weight B_micro / B_global, then exchanged ONCE.  The result equals step() over the
        concatenated dates (noise is keyed by the global unit id): to fp32 round-off in fp32 mode (measured 1e-7), to the bf16
        noise of the mode in bf16 mode (measured 1.4e-4 rel-L2: the 1 / B of the call sits inside the bf16 gradient tiles the
        backward kernels hand to each other) -- tests/test_shard_invariance_gpu.py."""
        self.step_index += 1
        acc = getattr(self, "_acc", None)
        if acc is None:
            acc = self._acc = torch.zeros_like(self.gradbuf)
        first = True
        for x, y, date_ptr, unit_base in micro_batches:
            self._local(x, y, date_ptr, unit_base=unit_base, train=train)
            w = float(date_ptr.numel() - 1) / float(global_dates)
            if first:
                torch.mul(self.gradbuf, w, out=acc)
                first = False
            else:
                acc.add_(self.gradbuf, alpha=w)
        self._reduce(acc, 1.0)
        self.gradbuf.copy_(acc)

    def capture(self, x, y: torch.Tensor, date_ptr: torch.Tensor, *, unit_base: int = 0, train: bool = True) -> "GraphedStep":
        """Capture forward + backward of this batch SHAPE in a CUDA graph (single GPU): see GraphedStep."""
        return GraphedStep(self, x, y, date_ptr, unit_base=unit_base, train=train)

    def step_from_host(self, x_host: torch.Tensor, y_host: torch.Tensor, date_ptr_host: torch.Tensor, **kw):
        """End-to-end entry: pinned HOST buffers in, loss (a Python float) out -- the H2D copy of the
        batch and the D2H read of the loss are part of the call."""
        dev = self.flat.device
        for name, src in (("x", x_host), ("y", y_host), ("date_ptr", date_ptr_host)):
            buf = self._dev.get(name)
            if buf is None or buf.shape != src.shape or buf.dtype != src.dtype:
                buf = torch.empty(src.shape, dtype=src.dtype, device=dev)
                self._dev[name] = buf
            buf.copy_(src, non_blocking=True)
        out, st = self.step(self._dev["x"], self._dev["y"], self._dev["date_ptr"], **kw)
        return float(self.loss.item()), out, st

    def run_from_host(self, batches, **kw):
        """Pipelined end-to-end loop: yields the loss (Python float) of every batch of `batches`, an iterable of
        pinned HOST tuples (x, y, date_ptr).  The H2D copy of batch i+1 runs on a copy stream while batch i computes,
        so a step costs max(copy, compute) instead of their sum; every batch is still copied and its loss read back."""
        dev = self.flat.device
        compute = torch.cuda.current_stream(dev)
        copy_stream = getattr(self, "_copy_stream", None)
        if copy_stream is None:
            copy_stream = self._copy_stream = torch.cuda.Stream(dev)
        bufs = [dict(), dict()]
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]

        def enqueue_copy(slot, batch):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])           # the step that last read this slot has finished
                for name, src in zip(("x", "y", "date_ptr"), batch):
                    buf = bufs[slot].get(name)
                    if buf is None or buf.shape != src.shape or buf.dtype != src.dtype:
                        buf = torch.empty(src.shape, dtype=src.dtype, device=dev)
                        bufs[slot][name] = buf
                    buf.copy_(src, non_blocking=True)
                copied[slot].record(copy_stream)

        it = iter(batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        consumed[0].record(compute)
        consumed[1].record(compute)
        enqueue_copy(0, nxt)
        slot = 0
        while nxt is not None:
            try:
                following = next(it)
            except StopIteration:
                following = None
            if following is not None:
                enqueue_copy(slot ^ 1, following)
            compute.wait_event(copied[slot])
            b = bufs[slot]
            self.step(b["x"], b["y"], b["date_ptr"], **kw)
            consumed[slot].record(compute)
            yield float(self.loss.item())                       # D2H read of this step's result
            nxt, slot = following, slot ^ 1

    def assign_grads(self, model) -> None:
        """Expose the flat gradient as `.grad` of the model's parameters (views, no copy)."""
        for name, p in model.named_parameters():
            p.grad = self.layout.view(self.grad, name)


class GraphedStep:
    """One ELBO step (forward + backward into stepper.grad / stepper.loss) of a fixed batch shape, captured once in a CUDA graph
    and replayed with one launch.  For launch-bound batches -- the reference's own per-date step (BASELINE.json configs[0]:
    64 stocks, ~15 kernels of a few microseconds each) -- the host-side launch path, not the GPU, sets the step time.

    `x`, `y`, `date_ptr` are the STATIC inputs: refill them in place (`x.copy_(...)`) between replays.  Kernel arguments are frozen
    at capture, so the Philox step counter lives in device memory (fvae_noise.step_dev) and the graph itself increments it:
    replay k draws exactly what `DateShardedStep.step` draws at step_index = first_step + k.  Parameters are read through
    stepper.flat at replay time, so an optimizer stepping them in place between replays is seen.  Single GPU only (the NVLink
    all-reduce takes its epoch as a kernel argument)."""

    def __init__(self, stepper: DateShardedStep, x, y: torch.Tensor, date_ptr: torch.Tensor, *, unit_base: int = 0,
                 train: bool = True, warmup: int = 2):
        if stepper.world != 1:
            raise NotImplementedError("graph capture of the step covers the single-GPU path")
        self.stepper, self.x, self.y, self.date_ptr = stepper, x, y, date_ptr
        dev = stepper.flat.device
        self.step_dev = torch.full((1,), int(stepper.step_index), dtype=torch.int64, device=dev)
        self.out = None

        def body():
            self.step_dev.add_(1)
            self.out, _ = stepper._local(x, y, date_ptr, unit_base=unit_base, train=train, step_dev=self.step_dev)

        side = torch.cuda.Stream(dev)                     # warm-up off the capture: kernel attributes, workspace, output tensors
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            body()
        # the warm-up and the capture pass advanced the counter's initial value by warmup (capture itself does not execute)
        stepper.step_index = int(self.step_dev.item())

    def replay(self):
        """Run the captured step; returns the output dict (static tensors, overwritten by the next replay)."""
        self.graph.replay()
        self.stepper.step_index += 1
        return self.out
