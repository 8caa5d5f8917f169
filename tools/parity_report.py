"""Prints the achieved parity numbers (native bf16 path and parity mode vs the committed fp32 reference golden vectors)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle.lrw_oracle as O  # noqa: E402
from test_lrw_gpu import _native, make_cfg, rel  # noqa: E402
from syncvsr_b200.lightning import TransformerLightningModule  # noqa: E402

for name in ("lrw_c1_vq", "lrw_c1_a2", "lrw_96_d2"):
    fx = torch.load(ROOT / "tests" / "golden" / f"{name}.pt")
    meta = fx["meta"]
    for mode in ("bf16", "parity"):
        m, P, (videos, tokens, labels, wm) = _native(TransformerLightningModule, meta)
        args = (videos.cuda(), tokens.cuda(), labels.cuda(), wm.cuda())
        with torch.no_grad():
            out = m(*args) if mode == "bf16" else m.forward_precise(*args)
        last = m.last_hidden_state().cpu()
        la = m.logits_audio().cpu().reshape(meta["B"], 29, -1)
        print(f"{name:10s} {mode:6s} loss_total rel {abs(float(out['loss_total'])/fx['metrics']['loss_total']-1):.2e} "
              f"last_hidden[cls] {rel(last[:,0,:], fx['last_hidden_state_cls']):.2e} last_hidden[t7] "
              f"{rel(last[:,7,:], fx['last_hidden_state_t7']):.2e} logits_audio[t3] {rel(la[:,3,:], fx['logits_audio_t3']):.2e} "
              f"logits_category {rel(m.logits_category().cpu(), fx['logits_category']):.2e}")


def _variant(name):
    """The dim-513 word-boundary and HuggingFace-BERT configurations (tests/test_lrw_gpu.py builds them the same way)."""
    fx = torch.load(ROOT / "tests" / "golden" / f"{name}.pt")
    meta = fx["meta"]
    if meta.get("wb"):
        m = TransformerLightningModule(make_cfg(depth=meta["depth"], use_wb=True)).train()
        P = O.make_params(meta["seed_p"], depth=meta["depth"], dim=513)
    else:
        cfg = make_cfg(depth=meta["depth"])
        cfg["model"]["bert"]["type"] = "huggingface"
        for k, v in meta["hf"].items():
            cfg["model"]["bert"][k] = v
        m = TransformerLightningModule(cfg).train()
        P = O.make_hf_params(O.make_params(meta["seed_p"], depth=meta["depth"]), meta["hf"], seed=meta["seed_p"] + 100)
    m.load_state_dict(P, strict=False)
    videos, tokens, labels, wm = O.make_inputs(meta["seed_x"], meta["B"])
    return fx, m, (videos.cuda(), tokens.cuda(), labels.cuda(), fx["word_mask"].cuda() if meta.get("wb") else wm.cuda())


for name in ("lrw_wb_d2", "lrw_hf_d2"):
    for mode in ("bf16", "parity"):
        fx, m, args = _variant(name)
        with torch.no_grad():
            out = m(*args) if mode == "bf16" else m.forward_precise(*args)
        last = m.last_hidden_state().cpu()
        la = m.logits_audio().cpu().reshape(fx["meta"]["B"], 29, -1)
        print(f"{name:10s} {mode:6s} loss_total rel {abs(float(out['loss_total'])/fx['metrics']['loss_total']-1):.2e} "
              f"last_hidden[cls] {rel(last[:,0,:], fx['last_hidden_state_cls']):.2e} last_hidden[t7] "
              f"{rel(last[:,7,:], fx['last_hidden_state_t7']):.2e} logits_audio[t3] {rel(la[:,3,:], fx['logits_audio_t3']):.2e} "
              f"logits_category {rel(m.logits_category().cpu(), fx['logits_category']):.2e}")
