"""LRS sentence-level step benchmark (BASELINE.json configs[2]/[3]: C3 = LRS2 T=150 B=16/GPU, C4 = LRS3 T=250 B=8/GPU,
lrs2.yaml widths: adim 768, 12 Conformer blocks, 6 decoder blocks, odim 5049, wav2vec2 codec 2x2x640).
A step = zero_grad + E2E.forward + backward + fused clip/AdamW + bf16 weight repack on one GPU, synthetic inputs
resident in HBM, CUDA-event timing. Prints one JSON line (a developer tool: the driver's contract is bench.py)."""
import argparse
import ctypes as C
import json
import os
import sys
from pathlib import Path
from types import SimpleNamespace

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from syncvsr_b200._lib import check, lib  # noqa: E402
from syncvsr_b200.e2e import E2E  # noqa: E402
from syncvsr_b200.train import FusedAdamW, SentenceDataParallelStep  # noqa: E402

# algorithmic forward GFLOP per clip (2*MAC; SURVEY.md section 8d probe, 88x88, adim 768): frontend + conformer + audio
# head + decoder (L=40); fwd+bwd = 3x forward minus the stem's input gradient (1.76 GF/29 frames per frame)
FWD_GF = {150: 94.8 + 54.9 + 0.59 + 6.5, 250: 158.1 + 93.4 + 0.98 + 8.0}


def args_ns(lmax):
    return SimpleNamespace(adim=768, aheads=12, eunits=3072, elayers=12, ddim=768, dheads=12, dunits=3072, dlayers=6,
                           mtlalpha=0.1, lsm_weight=0.1, dropout_rate=0.0, transformer_attn_dropout_rate=0.0,
                           transformer_input_layer="conv3d", transformer_encoder_attn_layer_type="rel_mha",
                           macaron_style=True, use_cnn_module=True, cnn_module_kernel=31, zero_triu=False,
                           a_upsample_ratio=1, relu_type="swish", transformer_length_normalized_loss=False,
                           ctc_type="builtin", rel_pos_type="latest", codec="wav2vec2", audio_weight=10.0,
                           max_label_len=lmax)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=150)
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--phases", action="store_true")
    a = ap.parse_args()
    B, T, Lmax = a.B, a.T, 40
    # data parallel under torchrun (one process per GPU, NCCL): SentenceDataParallelStep all-reduces the flat gradient arena
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(1234)
    m = E2E(5049, args_ns(Lmax)).train()
    opt = FusedAdamW(m, lr=1e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.03, max_grad_norm=5.0)
    dp = SentenceDataParallelStep(m, opt) if world > 1 else None
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    x = torch.randn(B, T, 1, 88, 88, device="cuda", generator=g)
    lengths = torch.randint(T // 2, T + 1, (B,), device="cuda", generator=g)
    lengths[0] = T
    for b in range(B):
        x[b, int(lengths[b]):] = 0
    tokens = torch.randint(0, 640, (B, 2 * T, 2), device="cuda", generator=g)
    label = torch.full((B, Lmax), -1, dtype=torch.long, device="cuda")
    for b in range(B):
        n = int(torch.randint(10, Lmax + 1, (1,), generator=g, device="cuda"))
        label[b, :n] = torch.randint(1, 5048, (n,), device="cuda", generator=g)
    label[0, :] = torch.randint(1, 5048, (Lmax,), device="cuda", generator=g)
    L = lib()
    L.svsr_launch_count.restype = C.c_longlong

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step(rec=None):
        if dp is not None:
            out = dp(x, lengths, tokens, label)
            m._ensure(x, Lmax)
            return out
        opt.zero_grad()
        if rec:
            rec[0].record()
        with torch.no_grad():
            out = m(x, lengths, tokens, label)
        if rec:
            rec[1].record()
        check(L.svsr_lrs_backward(m._h, C.c_void_p(0), m._stream()), "svsr_lrs_backward")
        if rec:
            rec[2].record()
        opt.step()
        m._ensure(x, Lmax)  # repack inside the step
        if rec:
            rec[3].record()
        return out

    for _ in range(a.warmup):
        out = step()
    torch.cuda.synchronize()
    n0 = L.svsr_launch_count()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(a.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # replicas must stay bit-identical: same all-reduced gradients, same AdamW
        chk = torch.stack([m.flat_params.double().sum(), m.flat_params.double().abs().max()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN), dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "parameter replicas diverged"
        if rank != 0:
            dist.destroy_process_group()
            return
    launches = (L.svsr_launch_count() - n0) // a.steps
    gf = FWD_GF.get(T, FWD_GF[150] * T / 150) * 3 - 1.76 * T / 29
    line = {"workload": f"LRS E2E step (Conformer-12L adim768 + CTC + decoder-6L + audio CE), x[{B},{T},1,88,88], bf16",
            "n_gpus": world, "ms_per_step": ms, "clips_per_s": world * B * 1e3 / ms, "frames_per_s": world * B * T * 1e3 / ms,
            "algorithmic_tflops_per_gpu": B * gf / ms, "launches_per_step": int(launches),
            "loss": [float(v) for v in out[:4]], "acc": float(out[4]), "workspace_gb": m._ws.numel() / 2**30,
            "params_M": m.flat_params.numel() / 1e6}
    if a.phases and world == 1:
        acc = [0.0, 0.0, 0.0]
        for _ in range(5):
            rec = [ev() for _ in range(4)]
            step(rec)
            torch.cuda.synchronize()
            for i in range(3):
                acc[i] += rec[i].elapsed_time(rec[i + 1]) / 5
        line["phase_ms"] = {"fwd": acc[0], "bwd": acc[1], "opt+repack": acc[2]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
