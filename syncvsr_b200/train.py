"""Data-parallel training step for the native LRW module: one process per GPU (torchrun), per-rank local BatchNorm
statistics (the reference never wires sync_batchnorm, LRS/video/main.py:33-49), ONE NCCL all-reduce of the flat
gradient arena per step (replaces DDP's bucketed reducer of Trainer(strategy="ddp"), LRW/video/src/train.py:28),
then the fused clip + AdamW kernel on every rank (lightning.py:216-223; gradient_clip_val train.py:32)."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch
import torch.distributed as dist

from ._lib import check, lib
from .lightning import TransformerLightningModule, _cfg_get


def cosine_with_warmup(step: int, base_lr: float, warmup: int, total: int) -> float:
    """transformers.get_scheduler("cosine") as used by the reference (lightning.py:222): linear warm-up then
    0.5*(1+cos(pi*progress))."""
    if step < warmup:
        return base_lr * step / max(1, warmup)
    progress = (step - warmup) / max(1, total - warmup)
    return base_lr * max(0.0, 0.5 * (1.0 + math.cos(math.pi * min(1.0, progress))))


def lrw_stage_of(key: str) -> int:
    """Backward stage (svsr_lrw_backward_stage) after which the gradient of LRW parameter `key` is final."""
    if key.startswith(("resnet.layer3.", "resnet.layer4.")):
        return 1
    if key.startswith(("stem3d.", "resnet.")):
        return 2
    return 0  # cls_token, encoder.*, audio_projection.*, category_classifier.*


class FusedAdamW:
    """AdamW over the module's flat arenas (decayed region first), with global-norm clipping fused in."""

    def __init__(self, module: TransformerLightningModule, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01,
                 max_grad_norm: float = 1.0):
        self.module = module
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        p = module.flat_params
        self.exp_avg = torch.zeros_like(p)
        self.exp_avg_sq = torch.zeros_like(p)
        self.scratch = torch.zeros(8, dtype=torch.float64, device=p.device)
        if hasattr(module, "n_decay"):  # LRS E2E mirror (e2e.py) records it when it builds its arenas
            self.n_decay = int(module.n_decay)
        else:
            L = lib()
            L.svsr_lrw_decay_count.restype = C.c_int64
            self.n_decay = int(L.svsr_lrw_decay_count(module._h))
        self.n_total = p.numel()
        self.t = 0
        self._build_segments()

    def _build_segments(self) -> None:
        """Ranges of the arena that share an Adam step count. Group 0 = everything that gets a gradient every step;
        group 1+i = x-transformers sublayer i (`encoder.layers.<i>.*`), which layer_dropout may skip on a step -- the
        reference's torch.optim.AdamW then leaves those parameters, their moments and their per-parameter step count
        untouched (`p.grad is None`, lightning.py:216-221)."""
        import re

        offs = getattr(self.module, "_offsets", None) or {}
        items = []
        for key, (off, n, _shp, _decay) in offs.items():
            mt = re.match(r"encoder\.layers\.(\d+)\.", key)
            items.append((off, (n + 3) // 4 * 4, 1 + int(mt.group(1)) if mt else 0))
        items.sort()
        ranges = []  # [begin, group]
        for off, _n, grp in items:
            if not ranges or ranges[-1][1] != grp:
                ranges.append([off, grp])
        if not ranges or ranges[0][0] != 0:
            ranges.insert(0, [0, 0])
        self._seg_group = [g for _, g in ranges]
        self._seg_begin = (C.c_int64 * (len(ranges) + 1))(*[b for b, _ in ranges], self.n_total)
        self._group_steps: Dict[int, int] = {g: 0 for g in set(self._seg_group)}

    @classmethod
    def from_config(cls, module: TransformerLightningModule) -> "FusedAdamW":
        o = dict(_cfg_get(module.config, "optim.optimizer", {}) or {})
        return cls(module, lr=float(o.get("lr", 1e-4)), betas=tuple(o.get("betas", (0.9, 0.999))),
                   eps=float(o.get("eps", 1e-6)), weight_decay=float(o.get("weight_decay", 0.01)),
                   max_grad_norm=float(_cfg_get(module.config, "train.gradient_clip_val", 1.0)))

    def zero_grad(self) -> None:
        self.module.flat_grads.zero_()

    def step(self, lr: Optional[float] = None, grad_div: float = 1.0, skip_mask: Optional[int] = None) -> None:
        """skip_mask: bit i set = x-transformers sublayer i was dropped in the forward this gradient came from (default:
        the module's last forward). Dropped sublayers are left untouched, like `p.grad is None` under torch.optim.AdamW."""
        self.t += 1
        m = self.module
        if skip_mask is None:
            skip_mask = int(getattr(m, "_last_skip", 0))
        for g in self._group_steps:
            if g == 0 or not (skip_mask >> (g - 1)) & 1:
                self._group_steps[g] += 1
        nseg = len(self._seg_group)
        steps = (C.c_int32 * nseg)(*[
            0 if (g > 0 and (skip_mask >> (g - 1)) & 1) else self._group_steps[g] for g in self._seg_group])
        check(lib().svsr_adamw_step_segmented(
            C.c_void_p(m.flat_params.data_ptr()), C.c_void_p(m.flat_grads.data_ptr()),
            C.c_void_p(self.exp_avg.data_ptr()), C.c_void_p(self.exp_avg_sq.data_ptr()), C.c_int64(self.n_decay),
            C.c_int64(self.n_total), C.c_float(self.lr if lr is None else lr), C.c_float(self.betas[0]),
            C.c_float(self.betas[1]), C.c_float(self.eps), C.c_float(self.weight_decay), self._seg_begin, steps,
            C.c_int(nseg), C.c_float(self.max_grad_norm), C.c_float(grad_div), C.c_void_p(self.scratch.data_ptr()),
            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "svsr_adamw_step_segmented")
        m.mark_weights_updated()

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm seen by the last step (device scalar, no sync)."""
        return self.scratch.view(torch.float32)[3]


class DataParallelStep:
    """zero_grad -> forward -> backward -> all-reduce(SUM) of the flat gradient arena -> fused clip+AdamW.

    The native forward/backward are called directly (no autograd graph); the reference-facing
    `module(...)`/`loss.backward()` API is equivalent and is what the parity tests use.

    graph=True: zero_grad + weight repack + forward + backward (~470 launches on two streams) are captured ONCE per
    set of input buffers into a CUDA graph and replayed; on N > 1 ranks they are three graphs (heads + encoder | layer4-3 | layer2-1 +
    stem) so that each group's all-reduce overlaps the next group's backward. Only a step whose launch sequence is the
    same every time can be replayed: with layer_dropout / dropout probabilities > 0 the per-step host RNG (dropped
    sublayers, mask seeds) is handed over in device memory and every sublayer is launched predicated
    (svsr_lrw_step_control), so the shipped training config replays from the same graph(s) too.
    high_priority=True: the step's main stream is a high-priority stream, so that whenever an SM frees up the
    critical chain (forward, input gradients, BatchNorm backward) is scheduled before the weight-gradient kernels
    the engine runs beside it on its own (default-priority) stream."""

    MAX_GRAPH_SETS = 4

    def __init__(self, module: TransformerLightningModule, optimizer: FusedAdamW, group=None, graph: bool = False,
                 high_priority: bool = False, staged: Optional[bool] = None):
        self.module, self.opt, self.group = module, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        sch = dict(_cfg_get(module.config, "optim.scheduler", {}) or {})
        self.warmup = int(sch.get("num_warmup_steps", 0))
        self.total = int(sch.get("num_training_steps", 1))
        self.global_step = 0
        # the gradient arena as contiguous slices per backward stage (heads + encoder | layer4 + layer3 | layer2-1 + stem)
        self._ranges = stage_ranges(module._offsets, lrw_stage_of)
        self.graph = bool(graph)
        self.staged = (self.world > 1) if staged is None else bool(staged)
        self._hi = torch.cuda.Stream(device=module.flat_params.device, priority=-1) if high_priority else None
        self._graphs: Dict[tuple, dict] = {}
        self._graph_gen = None      # engine generation the cached graphs were captured against
        self._static_inputs = None  # used once more than MAX_GRAPH_SETS distinct input buffer sets have been seen
        self._static: Dict[int, tuple] = {}  # ... per engine
        self.graph_launches = 0     # kernels launched by graph replays (the library's counter only sees captures)
        self.graph_replays = 0
        lib().svsr_launch_count.restype = C.c_longlong

    # ---- the three pieces of a step; `stage` as in svsr_lrw_backward_stage (-1 = whole backward) ----
    def _forward(self, batch):
        self.opt.zero_grad()
        with torch.no_grad():
            out = self.module(*batch)
        return out

    def _backward(self, stage: int) -> None:
        m = self.module
        if stage < 0:
            check(lib().svsr_lrw_backward(m._h, C.c_void_p(0), m._stream()), "svsr_lrw_backward")
        else:
            check(lib().svsr_lrw_backward_stage(m._h, C.c_void_p(0), C.c_int(stage), m._stream()),
                  f"backward stage {stage}")

    def _reduce(self, stage: int):
        """All-reduce (async, NCCL's stream) of the slices whose gradients backward stage `stage` has just completed."""
        return allreduce_ranges(self.module.flat_grads, self._ranges[stage], self.group) if self.world > 1 else []

    def _stochastic(self) -> bool:
        m = self.module
        return m.layer_dropout > 0.0 or m.ff_dropout > 0.0 or m.emb_dropout > 0.0 or m.attn_dropout > 0.0

    def _replayable(self) -> bool:
        m = self.module
        return self.graph and m.training and not m.hf and m._shape_key is not None

    def _capture(self, batch, ctl=None) -> dict:
        m = self.module
        cap_stream = self._hi if self._hi is not None else torch.cuda.Stream(device=m.flat_params.device)
        cap_stream.wait_stream(torch.cuda.current_stream())
        n0 = lib().svsr_launch_count()
        graphs, pool = [], None
        m._weights_dirty = True  # the repack belongs to every replay: the optimizer has always just run
        if self._stochastic():
            # layer_dropout / dropout seeds: device-resident control words instead of kernel arguments, every sublayer
            # launched predicated -- the captured launch sequence is the same for every step (svsr_lrw_step_control)
            m.device_control = True
            m._apply_step_control(*(ctl if ctl is not None else m._draw_step_control()))
            m._ctl_preset = (m._last_skip, 0)
        try:
            for part in ((0, 1, 2) if self.staged else (-1,)):
                g = torch.cuda.CUDAGraph()
                # thread_local: NCCL's watchdog thread polls events while we capture
                with torch.cuda.graph(g, pool=pool, stream=cap_stream, capture_error_mode="thread_local"):
                    if part <= 0:
                        metrics = self._forward(batch)
                    self._backward(part)
                pool = g.pool()
                graphs.append(g)
        finally:
            m._ctl_preset = None
        return {"graphs": graphs, "metrics": metrics, "batch": batch,
                "launches": int(lib().svsr_launch_count() - n0)}

    def _graph_entry(self, batch, ctl=None):
        gen = getattr(self.module, "_engine_gen", 0)
        if gen != self._graph_gen:  # another clip geometry = another engine: its graphs are kept while it is alive
            alive = self.module._engines.alive
            for k in [k for k in self._graphs if not alive(k[0])]:
                del self._graphs[k]
            for g in [g for g in self._static if not alive(g)]:
                del self._static[g]
            self._graph_gen = gen
        self._static_inputs = self._static.get(gen)
        key = (gen,) + tuple((t.data_ptr(), tuple(t.shape), t.dtype) for t in batch if isinstance(t, torch.Tensor))
        ent = self._graphs.get(key)
        if ent is None and self._static_inputs is None and sum(k[0] == gen for k in self._graphs) >= self.MAX_GRAPH_SETS:
            # the caller hands over fresh tensors every step: copy them into one static set from now on
            self._static_inputs = self._static[gen] = tuple(
                t.clone() if isinstance(t, torch.Tensor) else t for t in batch)
        if self._static_inputs is not None and ent is None:
            for d, src in zip(self._static_inputs, batch):
                if isinstance(d, torch.Tensor):
                    d.copy_(src, non_blocking=True)
            batch = self._static_inputs
            key = (gen, "static")
            ent = self._graphs.get(key)
        if ent is None:
            ent = self._graphs[key] = self._capture(batch, ctl)
        return ent

    def __call__(self, videos, audio_tokens, labels, word_mask=None) -> Dict[str, torch.Tensor]:
        m = self.module
        batch = (videos, audio_tokens, labels, word_mask)
        cur = torch.cuda.current_stream()
        if self._replayable() and m._shape_key == (videos.shape[0], videos.shape[2], videos.shape[3], videos.shape[4]):
            # this step's layer_dropout mask + dropout seed: host RNG like the reference, handed over in device memory
            ctl = m._draw_step_control() if self._stochastic() else None
            ent = self._graph_entry(batch, ctl)
            graphs = ent["graphs"]
            if ctl is not None:
                m._apply_step_control(*ctl)
            graphs[0].replay()
            if self.staged:
                hs = self._reduce(0)
                graphs[1].replay()
                hs += self._reduce(1)
                graphs[2].replay()
                hs += self._reduce(2)
                for h in hs:
                    h.wait()
            m._weights_dirty = False  # (num_batches_tracked += 1 is a device op: it is part of the graph)
            self.graph_replays += 1
            self.graph_launches += ent["launches"]
            metrics = ent["metrics"]
        else:
            # kernel-by-kernel launches (also the first step of a module: the engine is built and bound here)
            if self._hi is not None:
                self._hi.wait_stream(cur)
                torch.cuda.set_stream(self._hi)
            try:
                metrics = self._forward(batch)
                if not self.staged:
                    self._backward(-1)
                else:
                    # overlap: the encoder/head gradients (~160 MB, finished first) are all-reduced on NCCL's stream
                    # while the trunk backward runs, layer4 + layer3 (42 MB) while layer2-1 + stem run; only the last
                    # 3 MB cannot hide behind compute.
                    hs = []
                    for stage in range(3):
                        self._backward(stage)
                        hs += self._reduce(stage)
                    for h in hs:
                        h.wait()
            finally:
                if self._hi is not None:
                    torch.cuda.set_stream(cur)
                    cur.wait_stream(self._hi)
        # the reference's scheduler is stepped AFTER optimizer.step(): step k (0-based) runs at lr(k), lr(0) = 0
        lr = cosine_with_warmup(self.global_step, self.opt.lr, self.warmup, self.total) if self.total > 1 else self.opt.lr
        self.global_step += 1
        self.opt.step(lr=lr, grad_div=float(self.world))
        return metrics


def lrs_stage_of(key: str) -> int:
    """Backward stage (svsr_lrs_backward_stage) after which the gradient of parameter `key` is final."""
    if key.startswith(("decoder.", "ctc.", "audio_classifier.")):
        return 0
    if key.startswith(("encoder.encoders.", "encoder.after_norm.")):
        return 1
    return 2  # encoder.frontend.*, encoder.embed.*


def stage_ranges(offsets: Dict[str, tuple], stage_of=lrs_stage_of, n_stages: int = 3):
    """Contiguous element ranges of the flat arena per backward stage: {name: (offset, numel, ...)} -> [[(a, b), ...]].
    Tensors are 4-element aligned in the arena; adjacent tensors of one stage merge into one range (one all-reduce)."""
    per = [[] for _ in range(n_stages)]
    for key, (off, n, *_rest) in offsets.items():
        per[stage_of(key)].append((off, off + (n + 3) // 4 * 4))
    out = []
    for rs in per:
        rs.sort()
        merged = []
        for a, b in rs:
            if merged and merged[-1][1] == a:
                merged[-1] = (merged[-1][0], b)
            else:
                merged.append((a, b))
        out.append(merged)
    return out


def allreduce_ranges(flat: torch.Tensor, ranges, group=None):
    """Async SUM all-reduce of each (a, b) slice of the flat gradient arena; returns the work handles."""
    return [dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM, group=group, async_op=True) for a, b in ranges]


class SentenceDataParallelStep:
    """The same data-parallel step for the LRS sentence-level module (syncvsr_b200.e2e.E2E): zero_grad -> forward ->
    backward -> all-reduce(SUM) of the flat gradient arena -> fused clip + AdamW (LRS/video/lightning.py:89-96,
    gradient_clip_val 5.0, main.py:33-49). BatchNorm statistics stay per rank, as in the reference.

    The 1.0 GB gradient arena is reduced in three pieces as the backward retires them (svsr_lrs_backward_stage): the loss
    heads + decoder (65 M parameters) while the Conformer blocks compute, the blocks (171 M) while the frontend
    computes, and only the frontend + embed (12 M) after the last kernel -- what DDP's bucketed reducer does for the
    reference, with three buckets cut at the engine's stage boundaries."""

    MAX_GRAPH_SETS = 4

    def __init__(self, module, optimizer: FusedAdamW, group=None, warmup: int = 0, total: int = 1,
                 staged: Optional[bool] = None, graph: bool = False):
        """graph=True: zero_grad + weight repack + forward + backward (~1 250 launches on two streams) are captured once per
        set of input buffers and clip geometry into a CUDA graph (three under torchrun, cut at the backward stages so that
        each all-reduce still overlaps the next stage) and replayed. With dropout (the shipped lrs2/lrs3.yaml: 0.1 / 0.1)
        the step seed is drawn on the host as before but handed over in device memory (svsr_lrs_step_control): every
        dropout site adds it to its own constant on the device, so the same graph serves every step."""
        self.module, self.opt, self.group = module, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.warmup, self.total, self.global_step = warmup, total, 0
        self.staged = (self.world > 1) if staged is None else bool(staged)
        self._ranges = stage_ranges(module._offsets)
        self.graph = bool(graph)
        self._graphs: Dict[tuple, dict] = {}
        self._static: Dict[int, tuple] = {}
        self.graph_replays = 0
        self.graph_launches = 0

    def _backward(self, stage: int) -> None:
        m = self.module
        if stage < 0:
            check(lib().svsr_lrs_backward(m._h, C.c_void_p(0), m._stream()), "svsr_lrs_backward")
        else:
            check(lib().svsr_lrs_backward_stage(m._h, C.c_void_p(0), C.c_int(stage), m._stream()),
                  f"svsr_lrs_backward_stage {stage}")

    def _reduce(self, stage: int):
        return allreduce_ranges(self.module.flat_grads, self._ranges[stage], self.group) if self.world > 1 else []

    def _replayable(self, x, label) -> bool:
        m = self.module
        return (self.graph and m.training
                and m._shape_key == (x.shape[0], x.shape[1], x.shape[3], x.shape[4]) and label.shape[1] + 1 <= m._lmax)

    def _stochastic(self) -> bool:
        m = self.module
        return m.dropout_rate > 0.0 or m.attn_dropout_rate > 0.0

    def _capture(self, batch) -> dict:
        m = self.module
        cap = torch.cuda.Stream(device=m.flat_params.device)
        cap.wait_stream(torch.cuda.current_stream())
        n0 = lib().svsr_launch_count()
        graphs, pool, out = [], None, None
        if self._stochastic():  # the engine switches to the device-resident step seed (kept for this engine's lifetime)
            m._apply_step_seed(0)  # (the value is set per replay; nothing is drawn from the host RNG here)
        for part in ((0, 1, 2) if self.staged else (-1,)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool, stream=cap, capture_error_mode="thread_local"):
                if part <= 0:
                    m._weights_dirty = True  # the repack belongs to every replay: the optimizer has always just run
                    self.opt.zero_grad()
                    with torch.no_grad():
                        out = m(*batch)
                self._backward(part)
            pool = g.pool()
            graphs.append(g)
        return {"graphs": graphs, "out": out, "batch": batch, "launches": int(lib().svsr_launch_count() - n0)}

    def _graph_entry(self, batch) -> dict:
        gen = getattr(self.module, "_engine_gen", 0)
        alive = self.module._engines.alive
        for k in [k for k in self._graphs if not alive(k[0])]:  # graphs of a destroyed engine point into freed memory
            del self._graphs[k]
        for g in [g for g in self._static if not alive(g)]:
            del self._static[g]
        key = (gen,) + tuple((t.data_ptr(), tuple(t.shape), t.dtype) if isinstance(t, torch.Tensor) else None for t in batch)
        ent = self._graphs.get(key)
        static = self._static.get(gen)
        if ent is None and static is None and sum(k[0] == gen for k in self._graphs) >= self.MAX_GRAPH_SETS:
            static = self._static[gen] = tuple(t.clone() if isinstance(t, torch.Tensor) else t for t in batch)
        if ent is None and static is not None:
            if any(isinstance(d, torch.Tensor) and d.shape != t.shape for d, t in zip(static, batch)):
                return None  # another label / token length than the static set was made for: launch kernel by kernel
            for d, t in zip(static, batch):
                if isinstance(d, torch.Tensor):
                    d.copy_(t, non_blocking=True)
            batch, key = static, (gen, "static")
            ent = self._graphs.get(key)
        if ent is None:
            ent = self._graphs[key] = self._capture(batch)
        return ent

    def __call__(self, x, lengths, audios, label):
        m = self.module
        ent = None
        if self._replayable(x, label) and (audios is None or audios.dtype == torch.long) and x.dtype == torch.float32 \
                and lengths.dtype == torch.long and label.dtype == torch.long and x.is_contiguous():
            ent = self._graph_entry((x, lengths, audios, label))
        if ent is not None:
            graphs = ent["graphs"]
            if self._stochastic():
                m._apply_step_seed(m._step_seed())  # this step's masks: host RNG like the reference, read on the device
            graphs[0].replay()
            if self.staged:
                hs = self._reduce(0)
                graphs[1].replay()
                hs += self._reduce(1)
                graphs[2].replay()
                hs += self._reduce(2)
                for h in hs:
                    h.wait()
            m._weights_dirty = False
            self.graph_replays += 1
            self.graph_launches += ent["launches"]
            out = ent["out"]
        else:
            self.opt.zero_grad()
            with torch.no_grad():
                out = m(x, lengths, audios, label)
            if not self.staged:
                self._backward(-1)
                if self.world > 1:
                    dist.all_reduce(m.flat_grads, op=dist.ReduceOp.SUM, group=self.group)
            else:
                hs = []
                for stage in range(3):
                    self._backward(stage)
                    hs += self._reduce(stage)
                for h in hs:
                    h.wait()
        lr = cosine_with_warmup(self.global_step, self.opt.lr, self.warmup, self.total) if self.total > 1 else self.opt.lr
        self.global_step += 1
        self.opt.step(lr=lr, grad_div=float(self.world))
        return out


class PrefetchedStep:
    """Feeds a training step from pinned HOST batches: the host->device copy of batch i+1 runs on its own stream while
    batch i computes (two device buffer sets, event hand-off), and the step's loss is read back to pinned host memory.
    `step` is a DataParallelStep / SentenceDataParallelStep; batches are tuples of pinned CPU tensors of fixed shapes."""

    def __init__(self, step, example_batch):
        self.step = step
        dev = step.module.flat_params.device
        self.bufs = [tuple(torch.empty_like(t, device=dev) for t in example_batch) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.copied = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self.loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        self._i = 0
        self._primed = False

    def _enqueue_copy(self, slot: int, host_batch) -> None:
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[slot])  # the step that last read this slot has finished
            for d, h in zip(self.bufs[slot], host_batch):
                d.copy_(h, non_blocking=True)
            self.copied[slot].record(self.copy_stream)

    def __call__(self, host_batch, next_host_batch=None):
        """Runs one step on `host_batch`; if `next_host_batch` is given its copy overlaps this step's compute."""
        slot = self._i & 1
        cur = torch.cuda.current_stream()
        if not self._primed:
            for ev in self.consumed:
                ev.record(cur)
            self._enqueue_copy(slot, host_batch)
            self._primed = True
        if next_host_batch is not None:
            self._enqueue_copy(slot ^ 1, next_host_batch)
        cur.wait_event(self.copied[slot])
        out = self.step(*self.bufs[slot])
        self.consumed[slot].record(cur)
        loss = out["loss_total"] if isinstance(out, dict) else out[0]
        self.loss_host.copy_(loss.reshape(1), non_blocking=True)
        self._i += 1
        if next_host_batch is None:
            self._primed = False
        return out
