"""Prints the achieved parity numbers (native bf16 path and parity mode vs the committed fp32 reference golden vectors)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from test_lrw_gpu import _native, rel  # noqa: E402
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
