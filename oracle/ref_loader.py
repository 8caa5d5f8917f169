"""TEST INFRASTRUCTURE ONLY (oracle). Imports the UNMODIFIED reference LRW module from /root/reference
through `sys.modules` stubs for the packages that are not installed here.

Only usable in the build container (the reference tree does not exist on the GPU box); it is used by
tests/golden/make_golden.py to generate fixtures and by the CPU tests that pin oracle/lrw_oracle.py against
the reference's own forward(). Nothing on the product path imports this file."""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path

REF_ROOT = Path(os.environ.get("SVSR_REFERENCE", "/root/reference"))
REF_LRW_SRC = REF_ROOT / "LRW" / "video" / "src"


def reference_available() -> bool:
    return (REF_LRW_SRC / "lightning.py").exists()


class AttrDict(dict):
    """OmegaConf.DictConfig stand-in: attribute + item access, nested, `**cfg` works."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(d):
        if isinstance(d, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in d.items()})
        return d


_ref_module = None


def load_reference_lrw():
    """Returns the reference's `lightning` module (LRW/video/src/lightning.py), imported unmodified."""
    global _ref_module
    if _ref_module is not None:
        return _ref_module
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch.nn as nn
    import torchvision
    import transformers  # noqa: F401  -- must be imported BEFORE `timm` is stubbed (its lazy loader probes timm)
    from transformers import BertConfig, BertModel, Wav2Vec2ForPreTraining, get_scheduler  # noqa: F401

    from . import xt_encoder

    timm = types.ModuleType("timm")
    timm.create_model = lambda name, **kw: getattr(torchvision.models, name)(**kw)
    timm_optim = types.ModuleType("timm.optim")
    timm_optim.create_optimizer_v2 = None
    timm.optim = timm_optim

    pl = types.ModuleType("pytorch_lightning")

    class _LightningModule(nn.Module):
        def log_dict(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    pl.LightningModule = _LightningModule
    oc = types.ModuleType("omegaconf")
    oc.DictConfig = dict
    xt = types.ModuleType("x_transformers")
    xt.Encoder = xt_encoder.Encoder  # restatement: the real package is not installable here (parity unpinned)

    saved = {k: sys.modules.get(k) for k in ("timm", "timm.optim", "pytorch_lightning", "omegaconf", "x_transformers",
                                             "lightning", "utils", "augment", "tcn", "tcn.model")}
    sys.modules.update({"timm": timm, "timm.optim": timm_optim, "pytorch_lightning": pl, "omegaconf": oc,
                        "x_transformers": xt})
    sys.path.insert(0, str(REF_LRW_SRC))
    try:
        for k in ("lightning", "utils", "augment"):
            sys.modules.pop(k, None)
        import lightning as ref  # the reference's LRW/video/src/lightning.py

        _ref_module = ref
    finally:
        sys.path.remove(str(REF_LRW_SRC))
        # leave the stubs registered under their names only if nothing real was there before
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    return _ref_module


def reference_config(encoder_type: str = "x-transformers", depth: int = 12, use_wb: bool = False,
                     codec_path: str = "./vq-wav2vec_kmeans.pt", dropout: bool = False) -> AttrDict:
    """The shipped yaml (LRW/video/config/bert-12l-512d_LRW_96_bf16_rrc_noWB.yaml) as a dict, with cutmix off and
    (by default) every dropout zeroed so that forward() is deterministic."""
    import yaml

    name = "bert-12l-512d_LRW_96_bf16_rrc_WB.yaml" if use_wb else "bert-12l-512d_LRW_96_bf16_rrc_noWB.yaml"
    cfg = yaml.safe_load((REF_ROOT / "LRW" / "video" / "config" / name).read_text())
    cfg["train"]["use_cutmix"] = False
    cfg["model"]["wav2vec"]["path"] = codec_path
    cfg["model"]["bert"]["type"] = encoder_type
    cfg["model"]["bert"]["depth"] = depth
    cfg["data"]["use_word_boundary"] = use_wb
    if not dropout:
        cfg["model"]["bert"]["layer_dropout"] = 0.0
        cfg["model"]["bert"]["ff_dropout"] = 0.0
        cfg["model"]["bert"]["attn_dropout"] = 0.0
        cfg["model"]["bert"]["emb_dropout"] = 0.0
    return AttrDict.wrap(cfg)


# ------------------------------------------------------------------------------------------------------------------
# LRS sentence-level reference (vendored espnet under LRS/video): only `timm` needs a stub.
# ------------------------------------------------------------------------------------------------------------------
REF_LRS_ROOT = REF_ROOT / "LRS" / "video"
_ref_lrs = None


def reference_lrs_available() -> bool:
    return (REF_LRS_ROOT / "espnet" / "nets" / "pytorch_backend" / "e2e_asr_transformer.py").exists()


def load_reference_lrs():
    """Returns the reference's `E2E` class (LRS/video/espnet/nets/pytorch_backend/e2e_asr_transformer.py), unmodified."""
    global _ref_lrs
    if _ref_lrs is not None:
        return _ref_lrs
    if not reference_lrs_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torchvision
    import transformers  # noqa: F401  -- before stubbing timm

    had = sys.modules.get("timm")
    if had is None:
        timm = types.ModuleType("timm")
        timm.create_model = lambda name, **kw: getattr(torchvision.models, name)(**kw)
        sys.modules["timm"] = timm
    saved_utils = sys.modules.pop("utils", None)  # LRS/video/utils.py vs LRW/video/src/utils.py
    sys.path.insert(0, str(REF_LRS_ROOT))
    try:
        from espnet.nets.pytorch_backend.e2e_asr_transformer import E2E

        _ref_lrs = E2E
    finally:
        sys.path.remove(str(REF_LRS_ROOT))
        if saved_utils is not None:
            sys.modules["utils"] = saved_utils
    return _ref_lrs


def reference_lrs_args(**overrides):
    """model.visual_backbone of LRS/video/config/lrs2.yaml as an argparse.Namespace; every dropout zeroed by default."""
    import argparse

    import yaml

    cfg = yaml.safe_load((REF_LRS_ROOT / "config" / "lrs2.yaml").read_text())["model"]["visual_backbone"]
    cfg["dropout_rate"] = 0.0
    cfg["transformer_attn_dropout_rate"] = 0.0
    cfg.update(overrides)
    return argparse.Namespace(**cfg)


def build_reference_lrs(P, *, odim: int, audio_alignment: int, audio_vocab_size: int, n_audio: int, tokens, **overrides):
    """Reference E2E with the audio head enabled without the network download of e2e_asr_transformer.py:148
    (SURVEY.md section 8c): codec attributes set by hand, `forward_audios` returns the pre-made tokens."""
    import torch.nn as nn

    E2E = load_reference_lrs()
    m = E2E(odim, reference_lrs_args(**overrides))
    adim = overrides.get("adim", 768)
    m.codec = "wav2vec2"
    m.audio_alignment = audio_alignment
    m.audio_vocab_size = audio_vocab_size
    m.audio_weight = 10.0
    m.audio_classifier = nn.Linear(adim, n_audio)
    m.forward_audios = lambda audios: tokens
    missing, unexpected = m.load_state_dict(P, strict=False)
    assert not unexpected, unexpected
    assert all("num_batches_tracked" in k for k in missing), missing
    return m
