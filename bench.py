#!/usr/bin/env python
"""Benchmark of the SyncVSR LRW hot path (BASELINE.json: clips/sec fwd+bwd on LRW-shape [B,1,29,88,88]).

    python bench.py --gpus N --steps K --warmup W            # native sm_100a arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm (oracle port on the host cores)

A step = zero_grad + forward + backward (+ gradient all-reduce for N > 1) + fused clip/AdamW + bf16 weight repack over
one batch of B=64 clips per GPU (BASELINE.json configs[1]); `value` has the inputs resident in HBM, `e2e` copies them
from pinned host memory every step and reads the loss back. Prints ONE JSON line on rank 0."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "clips/sec (fwd+bwd) LRW-shape [B,1,29,88,88]"
FLOPS_PER_CLIP = 62.61e9  # fwd 21.46 GF, fwd+bwd 62.61 GF (2*MAC, A*G*V = 2560)
B_PER_GPU = 64
T, S = 29, 88
WORKLOAD = ("LRW word-level ResNet18+Transformer-12L (x-transformers), fwd+bwd+allreduce+AdamW, "
            f"[{B_PER_GPU},1,29,88,88] per GPU (BASELINE configs[1]); dropouts 0 = every sublayer runs every step "
            "(the shipped layer_dropout .2 / ff_dropout .3 step is timed beside it: `shipped_dropouts`)")


class _Attrs(dict):
    """dict with attribute access (stands in for the reference's OmegaConf DictConfig)."""

    __getattr__ = dict.get


def _attrs(d):
    return _Attrs({k: _attrs(v) if isinstance(v, dict) else v for k, v in d.items()})


def lrw_config(depth=12, layer_dropout=0.0, ff_dropout=0.0):
    return _attrs({
        "data": {"use_word_boundary": False, "input_size": S},
        "model": {"resnet": "resnet18", "wav2vec": {"path": "./vq-wav2vec_kmeans.pt"},
                  "bert": {"type": "x-transformers", "num_tokens": 1, "dim": 512, "depth": depth, "heads": 8,
                           "emb_dropout": 0.0, "attn_dropout": 0.0, "layer_dropout": layer_dropout,
                           "ff_dropout": ff_dropout,
                           "use_rmsnorm": True, "ff_glu": True, "rotary_pos_emb": True, "num_labels": 500}},
        "optim": {"optimizer": {"lr": 1e-4, "betas": [0.9, 0.999], "eps": 1e-6, "weight_decay": 0.01},
                  "scheduler": {"name": "cosine", "num_warmup_steps": 15000, "num_training_steps": 270000},
                  "lambda_audio": 10.0},
        "train": {"label_smoothing": 0.0, "use_cutmix": False, "precision": "bf16", "gradient_clip_val": 1.0},
    })


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.samples, self._stop_evt = gpu_index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("bf16_tflops_sustained", 1364.0), d.get("hbm_gbs", 6556.5), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# -------------------------------------------------------------------------------------------------------------------
# reference / CPU-baseline arm: the oracle port of the reference forward+backward on the host cores
# -------------------------------------------------------------------------------------------------------------------
def cpu_reference_clips_per_s(steps: int, warmup: int, batch: int = 2):
    import torch

    from oracle import lrw_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = {k: v.clone().requires_grad_("running" not in k) for k, v in O.make_params(0).items()}
    videos, tokens, labels, wm = O.make_inputs(1234, batch)

    def one():
        for v in P.values():
            v.grad = None
        out = O.lrw_forward(P, videos, tokens, labels, wm)
        out["loss_total"].backward()
        return float(out["loss_total"].detach())

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, cores, torch.get_num_threads()


def cpu_reference_lrs_clips_per_s(steps: int, warmup: int, T_: int, batch: int = 1):
    """The oracle port of E2E.forward (oracle/lrs_oracle.py, pinned to the reference's own module) + backward on the host
    cores: lrs2.yaml widths, `batch` clips of T_ frames per step."""
    import torch

    from oracle import lrs_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = {k: v.clone().requires_grad_("running" not in k and v.dtype.is_floating_point) for k, v in O.make_params(0).items()}
    x, lengths, tokens, label = O.make_inputs(1234, batch, T_)

    def one():
        for v in P.values():
            v.grad = None
        out = O.lrs_forward(P, x, lengths, tokens, label, elayers=12, dlayers=6, heads=12, odim=5049, audio_alignment=2,
                            audio_vocab_size=640)
        loss = out["loss"] if isinstance(out, dict) else out[0]
        loss.backward()
        return float(loss.detach())

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, cores, torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config in ("c3", "c4"):
        T_, B = (150, 16) if args.config == "c3" else (250, 8)
        steps, warmup = min(args.steps, 6), min(args.warmup, 1)
        cps, ms, cores, threads = cpu_reference_lrs_clips_per_s(steps, warmup, T_)
        sample = (f"oracle port of e2e_asr_transformer.py:186-227 fwd+bwd, fp32, B=1 x T={T_} x {steps} steps (+{warmup} "
                  f"warm-up), {threads} torch threads on {cores} host cores")
        print(json.dumps({
            "impl": "reference", "metric": f"clips/sec (fwd+bwd) LRS-shape [B,{T_},1,88,88]", "config_id": args.config,
            "value": cps, "unit": "clips/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": lrs_workload(T_, B), "global_batch": args.gpus * B, "parallelism": f"dp{args.gpus}",
                       "sample": f"each CPU step is a bounded sample of that workload: [1,{T_},1,88,88], fwd+bwd, fp32"},
            "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return
    steps, warmup = min(args.steps, 8), min(args.warmup, 2)
    cps, ms, cores, threads = cpu_reference_clips_per_s(steps, warmup)
    sample = (f"oracle port of lightning.py:133-191 fwd+bwd, fp32, B=2 x {steps} steps (+{warmup} warm-up), "
              f"{threads} torch threads on {cores} host cores")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": cps, "unit": "clips/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": args.gpus * B_PER_GPU, "parallelism": f"dp{args.gpus}",
                   "sample": "each CPU step is a bounded sample of that workload: [2,1,29,88,88] clips, fwd+bwd, fp32"},
        "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# -------------------------------------------------------------------------------------------------------------------
# torch.cuda eager comparator (BASELINE configs[1] "vs reference torch.cuda"): the reference's module
# graph from stock torch.nn modules (oracle/eager_module.py), autocast(bf16), channels_last trunk, clip + fused AdamW --
# every FLOP through ATen -> cuDNN / cuBLAS like the reference; none of this repo's kernels.
# -------------------------------------------------------------------------------------------------------------------
def eager_clips_per_s(steps: int, warmup: int, batch: int, local: int = 0, world: int = 1):
    import torch
    import torch.distributed as dist

    from oracle.eager_module import EagerLRW, EagerStep

    dev = torch.device("cuda", local)
    torch.manual_seed(1234 + local)
    model = EagerLRW(depth=12).to(dev).train()
    step = EagerStep(model)
    if world > 1:  # the reference's Trainer(strategy="ddp") with PL-1.x's find_unused_parameters=True (resnet.conv1/bn1/fc)
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True)
        step.model = ddp
    g = torch.Generator(device=dev).manual_seed(1234 + local)
    batches = [(torch.randn(batch, 1, T, S, S, device=dev, generator=g),
                torch.randint(0, 320, (batch, T * 4, 2), device=dev, generator=g),
                torch.randint(0, 500, (batch,), device=dev, generator=g), None) for _ in range(2)]
    for i in range(warmup):
        step(*batches[i % 2])
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out = step(*batches[i % 2])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    loss = float(out["loss_total"])
    del step, model, batches
    torch.cuda.empty_cache()
    return world * batch * 1e3 / ms, ms, loss


def run_eager_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cps, ms, loss = eager_clips_per_s(args.steps, max(args.warmup, 5), B_PER_GPU, local, world)
    if rank == 0:
        print(json.dumps({
            "impl": "eager", "metric": METRIC, "value": cps, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 5), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B_PER_GPU, "parallelism": f"dp{world}",
                       "how": "reference module graph from stock torch.nn (oracle/eager_module.py), autocast(bf16), "
                              "channels_last trunk, clip_grad_norm_ + fused torch.optim.AdamW"
                              + (", torch DDP(find_unused_parameters=True)" if world > 1 else ""),
                       "loss_total": loss},
            "gpu_launches": 0}))
    if world > 1:
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[4] (C5): audio-head stress -- seq 400, A=4, G=8, V=1024 (32 768 logits per frame), H=768, B=16/GPU.
# A step = fused head forward (projection + reshape + log-softmax + NLL, svsr_audio_head_fwd) + backward (recompute ->
# d logits bf16, input-gradient GEMM, weight-gradient GEMM). No exchange step exists on this path: ranks are replicas.
# -------------------------------------------------------------------------------------------------------------------
def head_stress(steps: int, warmup: int, B: int = 16, H: int = 768, seed: int = 0):
    import torch

    from syncvsr_b200 import ops

    T_, A, G, V = 400, 4, 8, 1024
    N, M = A * G * V, B * T_
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(M, H, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, H, device="cuda", generator=g) * 0.02).bfloat16()
    wt = w.t().contiguous()
    bias = torch.zeros(N, device="cuda")
    tokens = torch.randint(0, V, (B, T_ * A, G), device="cuda", generator=g)
    head = ops.AudioHead(B, T_, A, G, V, H)
    dx = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(N, H, device="cuda")

    def step():
        loss = head.forward(x, w, bias, tokens)
        head.backward(x, w, wt, bias, tokens, dx=dx, dw=dw, want_db=False)
        return loss

    for _ in range(warmup):
        loss = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    fused_bytes = M * H * 2 + N * H * 2 + tokens.numel() * 8 + M * H * 2 + N * H * 4  # X + W + tokens + dX + dW
    moved_bytes = fused_bytes + 3 * M * N * 2  # + d logits written once, read by the two gradient GEMMs
    flops = 3 * 2.0 * M * N * H                # algorithmic: projection + dX + dW (the recompute is not counted)
    return {"workload": f"audio-head stress: seq 400, A=4, G=8, V=1024, H={H}, B={B} (BASELINE configs[4])",
            "ms_per_step": ms, "frames_per_s": M * 1e3 / ms, "loss": float(loss), "algorithmic_tflops": flops / ms / 1e9,
            "algorithmic_MB": fused_bytes / 1e6, "algorithmic_GBps": fused_bytes / ms / 1e6,
            "designed_traffic_MB": moved_bytes / 1e6,
            "note": "fp32 logits never written; d logits cross HBM once in bf16 (TMEM cannot hold a [128 x H] dX accumulator "
                    "next to the logits tile, DESIGN.md section 3)"}


def run_head_arm(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    r = head_stress(args.steps, max(args.warmup, 3), seed=rank)
    ms = r["ms_per_step"]
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        peak_tf, peak_hbm, peak_src = measured_peaks()
        tf = r["algorithmic_tflops"] * r["ms_per_step"] / ms
        print(json.dumps({
            "metric": "frames/sec through the fused audio head (fwd+bwd), seq 400 x 32768 logits/frame", "config_id": "c5",
            "value": world * 6400 * 1e3 / ms, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": r["workload"], "parallelism": f"replicas x{world} (no exchange step on this path)",
                       "l2": "per-step working set 0.6 GB >> 126 MB L2", "loss": r["loss"]},
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                         "peak_source": peak_src, "traffic": None,
                         "algorithmic_MB": r["algorithmic_MB"], "designed_traffic_MB": r["designed_traffic_MB"]},
            "clocks": clocks, "gpu_launches": 5 * args.steps}))
    if world > 1:
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2] / [3] (C3 = LRS2 T=150 B=16/GPU, C4 = LRS3 T=250 B=8/GPU): the LRS sentence-level step
# (Conformer-12L adim 768 + CTC + 6-layer decoder + audio CE), zero_grad + forward + backward + all-reduce + clip/AdamW.
# -------------------------------------------------------------------------------------------------------------------
LRS_FWD_GF = {150: 94.8 + 54.9 + 0.59 + 6.5, 250: 158.1 + 93.4 + 0.98 + 8.0}  # SURVEY.md 8(d), per clip, forward


def lrs_args(lmax, dropout=0.0):
    from types import SimpleNamespace

    return SimpleNamespace(adim=768, aheads=12, eunits=3072, elayers=12, ddim=768, dheads=12, dunits=3072, dlayers=6,
                           mtlalpha=0.1, lsm_weight=0.1, dropout_rate=dropout, transformer_attn_dropout_rate=dropout,
                           transformer_input_layer="conv3d", transformer_encoder_attn_layer_type="rel_mha",
                           macaron_style=True, use_cnn_module=True, cnn_module_kernel=31, zero_triu=False,
                           a_upsample_ratio=1, relu_type="swish", transformer_length_normalized_loss=False,
                           ctc_type="builtin", rel_pos_type="latest", codec="wav2vec2", audio_weight=10.0,
                           max_label_len=lmax)


def lrs_workload(T_: int, B: int) -> str:
    return (f"LRS E2E step: Conformer-12L adim 768 + CTC + decoder-6L + audio CE, x[{B},{T_},1,88,88] per GPU, "
            f"fwd+bwd+allreduce+AdamW (BASELINE configs[{2 if T_ == 150 else 3}])")


def lrs_step_ms(steps: int, warmup: int, T_: int, B: int, rank: int = 0, world: int = 1, graph: bool = True,
                dropout: float = 0.0, e2e: bool = True):
    import torch
    import torch.distributed as dist

    from syncvsr_b200.e2e import E2E
    from syncvsr_b200.train import FusedAdamW, SentenceDataParallelStep

    Lmax = 40
    torch.manual_seed(1234)
    m = E2E(5049, lrs_args(Lmax, dropout)).train()
    opt = FusedAdamW(m, lr=1e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.03, max_grad_norm=5.0)
    dp = SentenceDataParallelStep(m, opt, graph=graph)
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    x = torch.randn(B, T_, 1, S, S, device="cuda", generator=g)
    lengths = torch.randint(T_ // 2, T_ + 1, (B,), device="cuda", generator=g)
    lengths[0] = T_
    for b in range(B):
        x[b, int(lengths[b]):] = 0
    tokens = torch.randint(0, 640, (B, 2 * T_, 2), device="cuda", generator=g)
    label = torch.full((B, Lmax), -1, dtype=torch.long, device="cuda")
    for b in range(B):
        n = int(torch.randint(10, Lmax + 1, (1,), generator=g, device="cuda"))
        label[b, :n] = torch.randint(1, 5048, (n,), device="cuda", generator=g)
    label[0, :] = torch.randint(1, 5048, (Lmax,), device="cuda", generator=g)

    def step():
        out = dp(x, lengths, tokens, label)
        if not dp.graph_replays:
            m._ensure(x, Lmax)  # the bf16 weight repack belongs to the step (graph mode: it is the graph's first node)
        return out

    for _ in range(warmup):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from syncvsr_b200._lib import lib

    lib().svsr_launch_count.restype = C.c_longlong
    n0, g0 = int(lib().svsr_launch_count()), dp.graph_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = int(lib().svsr_launch_count()) - n0 + dp.graph_launches - g0  # replays do not pass the library's counter
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    e2e_res, pipe = None, None
    if e2e:
        # ---- end to end through the public API: pinned HOST inputs copied every step (the copy of step i+1 overlaps step i
        # on a second stream), the step's loss read back to pinned host memory ----
        from syncvsr_b200.train import PrefetchedStep

        host = tuple(t.cpu().pin_memory() for t in (x, lengths, tokens, label))
        h2d = sum(t.numel() * t.element_size() for t in host)
        pipe = PrefetchedStep(dp, host)

        def e2e_run(n):
            for i in range(n):
                pipe(host, host if i + 1 < n else None)
                if not dp.graph_replays:
                    m._ensure(x, Lmax)

        e2e_run(2)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        e2e_run(steps)
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e_res = {"value": world * B * 1e3 / e2e_ms, "unit": "clips/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": 4}
    gf = LRS_FWD_GF.get(T_, LRS_FWD_GF[150] * T_ / 150) * 3 - 1.76 * T_ / 29
    res = {"e2e": e2e_res, "ms_per_step": ms, "clips_per_s": world * B * 1e3 / ms, "frames_per_s": world * B * T_ * 1e3 / ms,
           "algorithmic_tflops_per_gpu": B * gf / ms, "loss": [float(v) for v in out[:4]], "acc": float(out[4]),
           "workspace_gb": m._ws.numel() / 2 ** 30, "params_M": m.flat_params.numel() / 1e6,
           "graph_replays": dp.graph_replays, "launches": launches}
    del pipe, dp, opt, m
    torch.cuda.empty_cache()
    return res


def run_lrs_arm(args, T_: int, B: int, cid: str):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    r = lrs_step_ms(args.steps, max(args.warmup, 3), T_, B, rank, world, graph=bool(args.graph))
    clocks = sampler.stop() if sampler else None
    # the reference's yaml (lrs2.yaml / lrs3.yaml: dropout_rate 0.1, transformer_attn_dropout_rate 0.1): the LRS kernels take
    # their dropout seeds as launch arguments, so this step launches kernel by kernel
    rd = lrs_step_ms(max(args.steps // 2, 3), 3, T_, B, rank, world, graph=bool(args.graph), dropout=0.1, e2e=False)
    shipped = {"value": rd["clips_per_s"], "unit": "clips/s", "ms_per_step": rd["ms_per_step"],
               "launch_mode": "cuda graph replay" if rd["graph_replays"] else "kernel by kernel",
               "config": "dropout_rate 0.1, transformer_attn_dropout_rate 0.1 (the reference's yaml)"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cps, cms, cores, threads = cpu_reference_lrs_clips_per_s(6, 1, T_)
        cpu = {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
               "sample": f"oracle port of E2E.forward + backward (oracle/lrs_oracle.py), fp32, B=1 x T={T_} x 6 steps "
                         f"(+1 warm-up), {cms:.0f} ms/step, {threads} torch threads on {cores} host cores"}
    if rank == 0:
        peak_tf, _, peak_src = measured_peaks()
        print(json.dumps({
            "metric": f"clips/sec (fwd+bwd) LRS-shape [B,{T_},1,88,88]", "config_id": cid, "value": r["clips_per_s"],
            "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": lrs_workload(T_, B),
                       "global_batch": world * B, "parallelism": f"dp{world}", "loss": r["loss"], "acc": r["acc"],
                       "frames_per_s": r["frames_per_s"], "workspace_gb": r["workspace_gb"],
                       "launch_mode": "cuda graph replay" if r["graph_replays"] else "kernel by kernel"},
            "gpu_launches": r["launches"], "e2e": r["e2e"], "cpu_baseline": cpu, "shipped_dropouts": shipped,
            "roofline": {"bound": "tensor", "kernel": "whole step", "achieved": r["algorithmic_tflops_per_gpu"],
                         "peak": peak_tf, "unit": "TFLOP/s", "frac": r["algorithmic_tflops_per_gpu"] / peak_tf,
                         "peak_source": peak_src, "traffic": None},
            "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------------------------
# native arm
# -------------------------------------------------------------------------------------------------------------------
def run_native_arm(args):
    import torch
    import torch.distributed as dist

    from syncvsr_b200._lib import check, lib
    from syncvsr_b200.lightning import TransformerLightningModule
    from syncvsr_b200.train import DataParallelStep, FusedAdamW, PrefetchedStep

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = lib()
    L.svsr_launch_count.restype = C.c_longlong

    torch.manual_seed(1234 + rank)
    model = TransformerLightningModule(lrw_config()).train()
    opt = FusedAdamW.from_config(model)
    step = DataParallelStep(model, opt, graph=bool(args.graph), high_priority=bool(args.priority))

    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    B = B_PER_GPU
    n_batches = 2
    dev_batches = [(torch.randn(B, 1, T, S, S, device="cuda", generator=g),
                    torch.randint(0, 320, (B, T * 4, 2), device="cuda", generator=g),
                    torch.randint(0, 500, (B,), device="cuda", generator=g),
                    torch.zeros(B, 1, device="cuda")) for _ in range(n_batches)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    try:
        for i in range(args.warmup):
            step(*dev_batches[i % n_batches])
        torch.cuda.synchronize()
    except Exception as ex:  # a refused capture: the same kernels, launched one by one
        if not args.graph:
            raise
        print(f"bench: CUDA-graph capture failed ({ex!r}); launching kernel by kernel", file=sys.stderr)
        args.graph = 0
        step.graph = False
        step._graphs.clear()
        for i in range(args.warmup):
            step(*dev_batches[i % n_batches])
    barrier()

    # ---- timed region 1: device-resident inputs, CUDA events on the launching stream ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = L.svsr_launch_count() + step.graph_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        metrics = step(*dev_batches[i % n_batches])
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = L.svsr_launch_count() + step.graph_launches - launches0  # graph replays are counted by the step
    # ---- timed region 1b: the same K steps again with a CUDA-event pair around every tensor-core launch (on the
    # launching stream) -> per-family kernel time for the roofline; kept out of region 1 so that the ~300 extra event
    # records per step do not tax `value`
    check(L.svsr_prof_enable(1), "prof_enable")
    step.graph = False  # event pairs around single launches need kernel-by-kernel launches
    barrier()
    e0.record()
    for i in range(args.steps):
        step(*dev_batches[i % n_batches])
    e1.record()
    barrier()
    ms_prof_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    prof = {}
    for kind, name in ((0, "igemm_kernel"), (1, "wgrad_kernel")):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int()
        check(L.svsr_prof_read(kind, C.byref(ms), C.byref(fl), C.byref(n)), "prof_read")
        prof[name] = {"ms": ms.value, "flops": fl.value, "launches": n.value}
    check(L.svsr_prof_enable(0), "prof_enable")
    step.graph = bool(args.graph)
    loss = float(metrics["loss_total"])
    ms_per_step = ms_total / args.steps
    value = world * B * 1e3 / ms_per_step

    # ---- timed region 2: end to end through the public API with pinned HOST inputs ----
    host_batches = [tuple(t.cpu().pin_memory() for t in b) for b in dev_batches]
    h2d = sum(t.numel() * t.element_size() for t in host_batches[0])
    # public API: PrefetchedStep copies every step's inputs from pinned host memory (the copy of step i+1 overlaps the
    # compute of step i on a second stream) and reads every step's loss back to pinned host memory
    pipe = PrefetchedStep(step, host_batches[0])

    def e2e_run(n):
        for i in range(n):
            nxt = host_batches[(i + 1) % n_batches] if i + 1 < n else None
            pipe(host_batches[i % n_batches], nxt)

    e2e_run(2)
    barrier()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e_value = world * B * 1e3 / e2e_ms

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak_tf, peak_hbm, peak_src = measured_peaks()
    dom = max(prof, key=lambda k: prof[k]["ms"])
    d = prof[dom]
    achieved_tf = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
    share = {k: round(v["ms"] / ms_prof_total, 4) for k, v in prof.items()}
    traffic, traffic_src = None, None
    tp = ROOT / "profiles" / "r2_traffic.json"  # dram__bytes_read+write per launch from one ncu pass (tools/ncu_traffic.py)
    if tp.exists():
        fam = json.loads(tp.read_text()).get("families", {}).get(dom)
        if fam:
            traffic, traffic_src = fam["dram_bytes_per_launch"], "profiles/r2_traffic.json (ncu dram__bytes, avg per launch)"
    roofline = {
        "bound": "tensor", "kernel": dom, "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "launches_per_step": d["launches"] / args.steps, "avg_launch_us": d["ms"] * 1e3 / max(1, d["launches"]),
        "share_of_step": share, "instrumented_ms_per_step": ms_prof_total / args.steps,
        "whole_step_frac": value / world * FLOPS_PER_CLIP / (peak_tf * 1e12),
    }
    line = {
        "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "global_batch": world * B, "parallelism": f"dp{world}",
                   "l2": "per-step working set 4.5 GB >> 126 MB L2; two alternating input batches",
                   "launch_mode": ("cuda-graph replay of zero_grad+repack+fwd+bwd" if args.graph else "kernel by kernel")
                                  + (", high-priority main stream" if args.priority else ""),
                   "loss_total": loss},
        "e2e": {"value": e2e_value, "unit": "clips/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
    }
    # parity status of what was timed (VERDICT r1 #7e): everything on the path is pinned to the reference's own code
    # except the x-transformers encoder, whose package (x-transformers==1.9.2, LRW/video/setup.sh:30) is not obtainable
    # here -- its oracle is a restatement (oracle/xt_encoder.py)
    line["parity"] = "unpinned(a5: x_transformers.Encoder 1.9.2 restated); all other rows pinned to the reference"
    if world == 1 and not args.no_gpu_baseline:
        del pipe, step, opt, model
        torch.cuda.empty_cache()
        # the reference's shipped training config (LRW/video/config/bert-12l-512d_...yaml:25-28): layer_dropout .2 drops
        # ~20 % of the 24 sublayers per step (host RNG), ff_dropout .3 -- replayed from ONE CUDA graph through the
        # device-resident step control (svsr_lrw_step_control)
        try:
            torch.manual_seed(1234)
            m2 = TransformerLightningModule(lrw_config(layer_dropout=0.2, ff_dropout=0.3)).train()
            st2 = DataParallelStep(m2, FusedAdamW.from_config(m2), graph=bool(args.graph), high_priority=bool(args.priority))
            for i in range(max(args.warmup, 3)):
                st2(*dev_batches[i % n_batches])
            torch.cuda.synchronize()
            e0.record()
            for i in range(2 * args.steps):
                st2(*dev_batches[i % n_batches])
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / (2 * args.steps)
            line["shipped_dropouts"] = {"value": B * 1e3 / ms2, "unit": "clips/s", "ms_per_step": ms2,
                                        "steps": 2 * args.steps, "graph_replays": st2.graph_replays,
                                        "config": "layer_dropout 0.2, ff_dropout 0.3 (the reference's yaml); the mask and "
                                                  "the dropout seed change every step, the captured graph does not"}
            del st2, m2
            torch.cuda.empty_cache()
        except Exception as ex:
            line["shipped_dropouts"] = {"error": repr(ex)}
        # the other BASELINE configs, measured in the same run on the same GPU (their own arms: --config c3 | c4 | c5)
        also = {}
        try:
            also["c5_audio_head_stress"] = head_stress(steps=10, warmup=3)
            peak_tf_, _, _ = measured_peaks()
            also["c5_audio_head_stress"]["frac_of_tensor_peak"] = also["c5_audio_head_stress"]["algorithmic_tflops"] / peak_tf_
            also["c3_lrs2_T150_B16"] = lrs_step_ms(steps=5, warmup=3, T_=150, B=16, e2e=False)
        except Exception as ex:  # never lose the headline line to a side measurement
            also["error"] = repr(ex)
        line["also"] = also
        gcps, gms, gloss = eager_clips_per_s(steps=min(args.steps, 10), warmup=5, batch=B)
        line["gpu_baseline"] = {"value": gcps, "unit": "clips/s", "ms_per_step": gms, "kind": "torch.cuda eager",
                                "native_over_eager": value / gcps, "loss_total": gloss,
                                "sample": "reference module graph from stock torch.nn (oracle/eager_module.py), "
                                          f"autocast(bf16), channels_last trunk, clip + fused AdamW, B={B}, "
                                          f"{min(args.steps, 10)} steps (+5 warm-up), same GPU, after the native arm"}
    if world == 1 and not args.no_cpu_baseline:
        cps, cms, cores, threads = cpu_reference_clips_per_s(steps=4, warmup=1)
        line["cpu_baseline"] = {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
                                "sample": f"oracle port fwd+bwd fp32, B=2 x 4 steps (+1 warm-up), {cms:.0f} ms/step, "
                                          f"{threads} threads on {cores} cores"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "eager"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="BASELINE.json config: c2 = LRW B=64/GPU (the headline, default), c3 / c4 = LRS2 / LRS3 step, "
                         "c5 = audio-head stress")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="replay the step's launches from a CUDA graph (train.py)")
    ap.add_argument("--priority", type=int, default=1, help="run the step's main stream at high priority")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.impl == "eager":
        run_eager_arm(args)
    elif args.config == "c5":
        run_head_arm(args)
    elif args.config in ("c3", "c4"):
        run_lrs_arm(args, 150 if args.config == "c3" else 250, 16 if args.config == "c3" else 8, args.config)
    else:
        run_native_arm(args)


if __name__ == "__main__":
    main()
