"""Batch-sharded data parallelism for the heatmap networks (the reference has no
distributed code at all -- run1.job asks for one GPU).

One process per GPU (torchrun / torch.distributed, NCCL over NVLink).  All trainable parameters
live in ONE flat fp32 buffer and all gradients in ONE flat fp32 buffer, both laid out in
REVERSE execution order (decoder first, encoder conv1 last) and cut into a few contiguous
buckets.  The weight-gradient kernels write straight into the bucket memory (``param.grad`` is a
view), so there is no copy-in / copy-out around the collective; as soon as the last gradient of a
bucket has been enqueued an event is recorded and ``all_reduce(SUM)`` of that bucket is launched
on a communication stream, overlapping the rest of the backward pass.  The 1/world scaling is
folded into the fused Adam kernel, which updates the whole flat buffer in one launch.

Inert parameters (BatchNorm gamma/beta, ViT cls_token -- never touched by the forward, see
pytorch/CNNs.py:56-71) get no gradient in the reference either and are left out of the buckets.

Inference is frame-sharded with no collective: ``shard_range`` gives each rank its frames.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import nn


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous [begin, end) of `n_items` owned by `rank` (remainder spread over the low ranks)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class FlatBuckets:
    """Flat parameter / gradient storage with bucketed, overlapped all-reduce.

    `ordered` is the list of (name, parameter) in the order their gradients become ready
    (reverse execution order).  `bucket_bytes` is the target bucket size.
    """

    def __init__(self, ordered: Sequence[Tuple[str, nn.Parameter]], bucket_bytes: int = 4 << 20,
                 process_group=None, align: int = 64):
        self.names = [n for n, _ in ordered]
        self.params = [p for _, p in ordered]
        self.group = process_group
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + align - 1) // align * align  # keep every view 256-byte aligned
        self.offsets, self.total = offs, total
        self.flat_param = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, offs):
            self.flat_param[o:o + p.numel()].copy_(p.detach().reshape(-1))
        if self.world() > 1:
            # every rank starts from rank 0's parameters, whatever seed each process happened to construct with
            dist.broadcast(self.flat_param, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                           group=self.group)
        for p, o in zip(self.params, offs):
            p.data = self.flat_param[o:o + p.numel()].view(p.shape)
            p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)
        # buckets: contiguous runs of parameters, closed once they reach bucket_bytes
        self.buckets: List[Tuple[int, int, int]] = []  # (first param idx, last param idx, end offset)
        start_idx, start_off = 0, 0
        for i, p in enumerate(self.params):
            end = offs[i] + (p.numel() + align - 1) // align * align
            if (end - start_off) * 4 >= bucket_bytes or i == len(self.params) - 1:
                self.buckets.append((start_idx, i, end))
                start_idx, start_off = i + 1, end
        self._bucket_of = {}
        for b, (lo, hi, _) in enumerate(self.buckets):
            for i in range(lo, hi + 1):
                self._bucket_of[self.names[i]] = b
        self._pending: Dict[int, int] = {}
        self._works: list = []
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.reset()

    # ---- geometry helpers --------------------------------------------------------------------
    def bucket_slice(self, b: int) -> slice:
        lo = self.buckets[b][0]
        return slice(self.offsets[lo], self.buckets[b][2])

    def bucket_sizes_bytes(self) -> List[int]:
        return [(s.stop - s.start) * 4 for s in (self.bucket_slice(b) for b in range(len(self.buckets)))]

    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    # ---- per-step protocol -------------------------------------------------------------------
    def reset(self) -> None:
        """call before each backward: every bucket waits for all of its parameters."""
        self._pending = {b: hi - lo + 1 for b, (lo, hi, _) in enumerate(self.buckets)}
        self._works = []

    def grad_ready(self, name: str) -> None:
        """the kernel producing `name`'s gradient has been enqueued on the current stream."""
        b = self._bucket_of.get(name)
        if b is None:
            return
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b: int) -> None:
        if self.world() == 1:
            return
        view = self.flat_grad[self.bucket_slice(b)]
        if self.comm_stream is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                work = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            work = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._works.append((b, work))

    def flush(self) -> None:
        """launch any bucket whose gradients were not all reported (e.g. frozen layers)."""
        for b, left in list(self._pending.items()):
            if left > 0:
                self._pending[b] = 0
                self._launch(b)

    def wait(self) -> None:
        """make the compute stream wait for every outstanding bucket reduction."""
        for _ in self.wait_each():
            pass

    def wait_each(self):
        """yields bucket indices in launch order, each once the compute stream has been made to wait for THAT
        bucket's reduction only -- the caller can consume bucket b (fused Adam on its slice) while later, smaller
        buckets are still on the wire."""
        works, self._works = self._works, []
        if not works:                       # single process: nothing was launched
            yield from range(len(self.buckets))
            return
        seen = set()
        for b, w in works:
            w.wait()                        # stream-level wait: the current stream waits for this collective
            seen.add(b)
            yield b
        for b in range(len(self.buckets)):
            if b not in seen:
                yield b

    def params_checksum(self) -> torch.Tensor:
        """two int64 words over the flat parameter buffer's BITS (sum and index-weighted sum of the fp32 words):
        equal on every rank iff the ranks hold bit-identical parameters (bench.py's dp_check)."""
        bits = self.flat_param.view(torch.int32).to(torch.int64)
        idx = torch.arange(1, bits.numel() + 1, device=bits.device, dtype=torch.int64)
        return torch.stack((bits.sum(), (bits * (idx % 65521)).sum()))


def reverse_execution_order(model: nn.Module) -> List[Tuple[str, nn.Parameter]]:
    """live (gradient-receiving) parameters of a heatmap network, last layer first."""
    live: List[Tuple[str, nn.Parameter]] = []
    for name, p in model.named_parameters():
        if ".bn" in name or name.endswith("cls_token") or not p.requires_grad:
            continue
        live.append((name, p))
    return list(reversed(live))


class FusedAdam:
    """torch.optim.Adam(lr=1e-3) semantics (pytorch/train_pytorch.py:111) as one kernel launch over
    the flat parameter buffer."""

    def __init__(self, buckets: FlatBuckets, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, on_update: Optional[Callable[[], None]] = None):
        """on_update: called after the parameters changed (DataParallelStep refreshes every packed operand with one
        launch there).  Without it nothing is lost: ops.adam_step bumps the package's weight generation, which every
        cached packed operand is tagged with, so the next forward re-packs lazily."""
        self.b, self._lr, self.betas, self.eps, self.wd = buckets, float(lr), betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(buckets.flat_param)
        self.exp_avg_sq = torch.zeros_like(buckets.flat_param)
        self.step_count = 0
        self.on_update = on_update
        # CUDA-graph-safe stepping (DataParallelStep.enable_graph): the step number and the learning rate live on the
        # device, so a captured Adam launch reads the CURRENT values at every replay
        self.step_dev: Optional[torch.Tensor] = None
        self.lr_dev: Optional[torch.Tensor] = None

    def use_device_scalars(self) -> None:
        dev = self.b.flat_param.device
        if self.step_dev is None:
            self.step_dev = torch.full((1,), self.step_count, device=dev, dtype=torch.int32)
            self.lr_dev = torch.full((1,), self._lr, device=dev, dtype=torch.float32)

    @property
    def lr(self) -> float:
        return self._lr

    @lr.setter
    def lr(self, value: float) -> None:
        self._lr = float(value)
        if getattr(self, "lr_dev", None) is not None:
            self.lr_dev.fill_(self._lr)

    def sync_step_count(self, step_count: int) -> None:
        """after a state restore: host and device step numbers agree again."""
        self.step_count = int(step_count)
        if self.step_dev is not None:
            self.step_dev.fill_(self.step_count)

    def step(self, grad_scale: float = 1.0, per_bucket: bool = False) -> None:
        """per_bucket: one launch per gradient bucket, each as soon as that bucket's all-reduce has landed, so the
        update of the large early buckets overlaps the reduction of the last one (the only collective that cannot
        hide behind the backward pass)."""
        from . import ops
        self.step_count += 1
        kw = dict(lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.wd, grad_scale=grad_scale,
                  step_dev=self.step_dev, lr_dev=self.lr_dev)
        if per_bucket:
            n_b = len(self.b.buckets)
            for k, b in enumerate(self.b.wait_each()):
                sl = self.b.bucket_slice(b)
                ops.adam_step(self.b.flat_param[sl], self.b.flat_grad[sl], self.exp_avg[sl], self.exp_avg_sq[sl],
                              self.step_count, inc_step=(k == n_b - 1), **kw)
        else:
            self.b.wait()
            ops.adam_step(self.b.flat_param, self.b.flat_grad, self.exp_avg, self.exp_avg_sq, self.step_count,
                          inc_step=True, **kw)
        if self.on_update is not None:
            self.on_update()

    def zero_grad(self) -> None:
        self.b.flat_grad.zero_()

    def state_dict(self) -> dict:
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "lr": self.lr}

    # ---- torch.optim.Adam-compatible checkpoint schema (pytorch/train_pytorch.py:253-260 saves
    #      optimizer.state_dict() of Adam(model.parameters())) ------------------------------------------
    def torch_state_dict(self, model: nn.Module) -> dict:
        """the dict torch.optim.Adam(model.parameters(), lr).state_dict() would hold after the same steps:
        parameters indexed in model.parameters() order, state only for parameters that received gradients."""
        index = {id(p): i for i, p in enumerate(model.parameters())}
        state = {}
        if self.step_count > 0:
            for p, o in zip(self.b.params, self.b.offsets):
                n = p.numel()
                state[index[id(p)]] = {"step": torch.tensor(float(self.step_count)),
                                       "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                                       "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.wd,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                 "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "params": list(range(len(index)))}
        return {"state": state, "param_groups": [group]}

    def load_torch_state_dict(self, model: nn.Module, sd: dict) -> None:
        index = {id(p): i for i, p in enumerate(model.parameters())}
        steps = set()
        for p, o in zip(self.b.params, self.b.offsets):
            st = sd["state"].get(index[id(p)])
            if st is None:
                continue
            n = p.numel()
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("FusedAdam keeps one step counter; the checkpoint has per-parameter steps %s" % sorted(steps))
        self.sync_step_count(steps.pop() if steps else 0)
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.wd = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]


class DataParallelStep:
    """One data-parallel optimisation step of a BasicNet-like module exposing ``train_step``:
    local forward/backward on this rank's shard, bucketed all-reduce overlapped with backward,
    fused Adam with the 1/world factor."""

    def __init__(self, model: nn.Module, lr: float = 1e-3, bucket_bytes: Optional[int] = None, process_group=None):
        import os
        if bucket_bytes is None:
            # measured on 2 and 8 B200s (profiles/r2k_dp_bucket_probe.txt): the heatmap CNN's 10 MB of gradients reduce
            # fastest as ONE collective behind the backward pass.  Smaller buckets do overlap it, but every
            # contraction kernel is a persistent one-CTA-per-SM grid that owns the whole register file, so the SMs a
            # concurrent NCCL kernel occupies start their share of the next contraction late and the collective's
            # duration lands on the critical path anyway (6.43 ms with 4 MB buckets vs 6.35 ms per step at 8 GPUs).
            # Models with more parameters than this (the ViT: 126 MB) still split into overlapped buckets.
            bucket_bytes = int(os.environ.get("POSEB200_BUCKET_BYTES", 64 << 20))
        self.model = model
        self.buckets = FlatBuckets(reverse_execution_order(model), bucket_bytes, process_group)
        self.opt = FusedAdam(self.buckets, lr=lr, on_update=self._weights_changed)
        self.world = self.buckets.world()
        if hasattr(model, "set_grad_ready_hook"):
            model.set_grad_ready_hook(self.buckets.grad_ready)
        self._graph_on = False
        self._graphs: Dict[tuple, tuple] = {}
        self._eager_seen: Dict[tuple, int] = {}

    # ---- CUDA-graph replay of the whole step ---------------------------------------------------------------------
    def enable_graph(self, on: bool = True) -> None:
        """Replay the optimisation step -- forward, loss, backward, bucket all-reduces on the communication stream,
        fused Adam, weight re-pack: ~60 launches -- from ONE CUDA graph per input shape instead of launching it from
        Python.  The first two calls of a shape run eagerly (they perform every lazy initialisation: kernel
        attributes, workspaces, packed-weight caches); the third is captured and replayed, later ones only replay.
        Every call is exactly one training step either way.  Applies to the plain call `step(x, points=...)` or
        `step(x, target)`; accumulation / extra model inputs / per-launch profiling take the eager path."""
        self._graph_on = bool(on)
        if on:
            self.opt.use_device_scalars()

    def _graph_step(self, x: torch.Tensor, target: Optional[torch.Tensor], points: Optional[torch.Tensor]) -> torch.Tensor:
        from . import ops
        key = (tuple(x.shape), x.dtype, None if target is None else tuple(target.shape),
               None if points is None else tuple(points.shape))
        entry = self._graphs.get(key)
        if entry is None:
            seen = self._eager_seen.get(key, 0)
            if seen < 2:
                self._eager_seen[key] = seen + 1
                return self._eager_step(x, target, points)
            sx = torch.empty_like(x)
            st = None if target is None else torch.empty_like(target)
            sp = None if points is None else torch.empty_like(points)
            sx.copy_(x)
            if st is not None:
                st.copy_(target)
            if sp is not None:
                sp.copy_(points)
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g):
                loss = self._eager_step(sx, st, sp, count=False)
            entry = (g, sx, st, sp, loss, ops.launch_count() - n0)
            self._graphs[key] = entry
        else:
            g, sx, st, sp, loss, _ = entry
            sx.copy_(x)
            if st is not None:
                st.copy_(target)
            if sp is not None:
                sp.copy_(points)
        entry[0].replay()
        self.opt.step_count += 1          # the captured Adam launches advanced the device-side step number
        ops.note_launches(entry[5])
        # (the replay re-packed the weight operands itself -- the captured step ends with the re-pack launch -- so the
        #  host-side cache tags, which were valid when the graph was captured, still describe them)
        return entry[4]

    def _eager_step(self, x, target, points, count: bool = True) -> torch.Tensor:
        self.buckets.reset()
        loss = self.model.train_step(x, target, points=points)
        self.buckets.flush()
        if count:
            self.opt.step(grad_scale=1.0 / self.world, per_bucket=self.world > 1)
        else:
            # under capture: same launches, but the host-side step counter moves at replay time
            self.opt.step(grad_scale=1.0 / self.world, per_bucket=self.world > 1)
            self.opt.step_count -= 1
        return loss

    def _weights_changed(self) -> None:
        if hasattr(self.model, "repack_weights"):
            self.model.repack_weights()
        elif hasattr(self.model, "invalidate_packed_weights"):
            self.model.invalidate_packed_weights()

    def step(self, x: torch.Tensor, target: Optional[torch.Tensor] = None, *, points: Optional[torch.Tensor] = None,
             accumulation_steps: int = 1, micro_index: int = 0, accumulate: Optional[bool] = None,
             do_step: Optional[bool] = None, **model_kwargs) -> torch.Tensor:
        """returns the local mean loss (device tensor).  With accumulation the collective and the
        optimiser run only on the last micro-batch (pytorch/train_pytorch.py:139-142).  `accumulate` /
        `do_step` override the micro_index arithmetic (the Trainer uses them to keep the reference's
        behaviour of gradients that are never stepped leaking into the next epoch)."""
        if (self._graph_on and accumulation_steps == 1 and micro_index == 0 and not accumulate and do_step in (None, True)
                and not model_kwargs and x.is_cuda):
            from . import ops
            if not ops.profiling_active():
                return self._graph_step(x, target, points)
        last = (micro_index + 1) % accumulation_steps == 0 if do_step is None else bool(do_step)
        acc = (micro_index % accumulation_steps) != 0 if accumulate is None else bool(accumulate)
        self.buckets.reset()
        hooked = hasattr(self.model, "set_grad_ready_hook")
        if not last and hooked:
            self.model.set_grad_ready_hook(None)
        try:
            # model_kwargs: further inputs of the model's step (the camera matrices of FourCamerasDisentanglement)
            loss = self.model.train_step(x, target, points=points, accumulation_steps=accumulation_steps,
                                         accumulate=acc, **model_kwargs)
        finally:
            if hooked:
                self.model.set_grad_ready_hook(self.buckets.grad_ready)
        if last:
            self.buckets.flush()
            self.opt.step(grad_scale=1.0 / self.world, per_bucket=self.world > 1)
        return loss
