"""Host-side mirror of the reference's LRS sentence-level module
(/root/reference/LRS/video/espnet/nets/pytorch_backend/e2e_asr_transformer.py:29-227).

`E2E(odim, args, ignore_id=-1)` keeps the reference's constructor arguments, attribute names (`encoder` with
`.frontend.frontend3D / .frontend.trunk / .embed / .encoders / .after_norm`, `ctc`, `decoder`, `criterion`,
`audio_classifier`, `audio_weight`, `audio_alignment`, `audio_vocab_size`, `codec`, `odim`, `sos`, `eos`, `adim`,
`mtlalpha`), `forward(x, lengths, audios, label) -> (loss, loss_ctc, loss_att, loss_audio, acc)`,
`encoder(xs, masks) -> (xs, masks)` and the state-dict keys, so the reference's LRS training loop
(LRS/video/lightning.py:89-130) can construct it and load the reference's checkpoints. All arithmetic runs in the
sm_100a kernels of libsvsr.so through the native step executor (csrc/engine_lrs.cu); PyTorch only owns the memory.
There is no CPU / eager fallback: without the shared library or a CUDA device construction fails.
"""
from __future__ import annotations

import ctypes as C
import math
import random
import weakref
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from ._lib import SvsrError, check, lib
from .lightning import allreduce_mean_


class LrsConfig(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("T", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("adim", C.c_int), ("aheads", C.c_int), ("eunits", C.c_int), ("elayers", C.c_int),
        ("dlayers", C.c_int), ("dunits", C.c_int), ("odim", C.c_int), ("cnn_kernel", C.c_int),
        ("audio_alignment", C.c_int), ("vq_groups", C.c_int), ("audio_vocab", C.c_int), ("Lmax", C.c_int),
        ("mtlalpha", C.c_float), ("lsm_weight", C.c_float), ("audio_weight", C.c_float),
        ("bn_eps", C.c_float), ("bn_momentum", C.c_float),
        ("dropout_rate", C.c_float), ("attn_dropout_rate", C.c_float),
    ]


def _arg(args: Any, key: str, default: Any = None) -> Any:
    if isinstance(args, dict):
        v = args.get(key, None)
    else:
        v = getattr(args, key, None)
    return default if v is None else v


class _Node(nn.Module):
    """Anonymous container reproducing the reference's module tree (and therefore its state-dict keys)."""


class _EncoderNode(_Node):
    """`model.encoder(xs, masks)` (transformer/encoder.py:257-289) on the native engine."""

    def __init__(self, owner: "E2E"):
        super().__init__()
        object.__setattr__(self, "_owner", weakref.ref(owner))

    def forward(self, xs: torch.Tensor, masks: Optional[torch.Tensor], extract_resnet_feats: bool = False):
        return self._owner()._encode(xs, masks, extract_resnet_feats)


class _StepFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor: torch.Tensor, module: "E2E", metrics: torch.Tensor):
        ctx.module = module
        return metrics.clone()

    @staticmethod
    def backward(ctx, grad_metrics: torch.Tensor):
        ctx.module._native_backward(grad_metrics)
        return torch.zeros((), device=grad_metrics.device), None, None


class E2E(nn.Module):
    def __init__(self, odim: int, args: Any, ignore_id: int = -1, device: Optional[torch.device | str] = None):
        super().__init__()
        if not torch.cuda.is_available():
            raise SvsrError("syncvsr_b200 needs a CUDA (sm_100a) device: the hot path has no CPU fallback")
        lib()
        self.device_ = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.args = args
        self.odim = int(odim)
        self.ignore_id = int(ignore_id)
        if self.ignore_id != -1:
            raise SvsrError("only ignore_id == -1 is supported (the reference's only call site uses it)")
        self.blank, self.sos, self.eos = 0, self.odim - 1, self.odim - 1
        self.adim = int(_arg(args, "adim", 768))
        self.aheads = int(_arg(args, "aheads", 12))
        self.eunits = int(_arg(args, "eunits", 3072))
        self.elayers = int(_arg(args, "elayers", 12))
        self.ddim = int(_arg(args, "ddim", self.adim))
        self.dheads = int(_arg(args, "dheads", self.aheads))
        self.dunits = int(_arg(args, "dunits", self.eunits))
        self.dlayers = int(_arg(args, "dlayers", 6))
        self.mtlalpha = float(_arg(args, "mtlalpha", 0.1))
        self.lsm_weight = float(_arg(args, "lsm_weight", 0.1))
        self.cnn_kernel = int(_arg(args, "cnn_module_kernel", 31))
        self.transformer_input_layer = _arg(args, "transformer_input_layer", "conv3d")
        self.a_upsample_ratio = int(_arg(args, "a_upsample_ratio", 1))
        # configurations of E2E.__init__ the native path does not implement are refused, never silently changed
        if self.ddim != self.adim or self.dheads != self.aheads:
            raise SvsrError("proj_decoder (adim != ddim) / dheads != aheads are not built in the sm_100a path")
        if self.transformer_input_layer != "conv3d":
            raise SvsrError(f"transformer_input_layer={self.transformer_input_layer!r}: only the conv3d visual frontend is native")
        if _arg(args, "transformer_encoder_attn_layer_type", "rel_mha") != "rel_mha" or not _arg(args, "macaron_style", True) \
                or not _arg(args, "use_cnn_module", True) or _arg(args, "relu_type", "swish") != "swish":
            raise SvsrError("only the rel_mha / macaron / cnn-module / swish Conformer of lrs2.yaml is native")
        if _arg(args, "zero_triu", False) or _arg(args, "transformer_length_normalized_loss", False):
            raise SvsrError("zero_triu / length-normalised loss are not supported")
        if not (0.0 < self.mtlalpha < 1.0):
            raise SvsrError("the native step needs both CTC and attention losses (0 < mtlalpha < 1)")
        # dropout_rate: every sub-block output, FFN hidden units, positional encodings, CTC input;
        # transformer_attn_dropout_rate: attention probabilities (encoder and decoder). Training mode only.
        self.dropout_rate = float(_arg(args, "dropout_rate", 0.0))
        self.attn_dropout_rate = float(_arg(args, "transformer_attn_dropout_rate", 0.0))
        self.dropout_seed: Optional[int] = None  # fixed seed for reproducible masks (tests); None = fresh per step
        # Cross-modal sync head (e2e_asr_transformer.py:126-160). Explicit keys win; else derived from the codec name.
        codec = _arg(args, "codec", None)
        self.codec = None
        a = g = v = 0
        if codec is not None and "vq" in str(codec).lower():
            self.codec, a, g, v = "vq", 4, 2, 320
        elif codec is not None and "wav2vec2" in str(codec).lower():
            self.codec, a, g, v = "wav2vec2", 2, 2, 640
        elif codec is not None:
            raise SvsrError(f"unknown codec {codec!r}")
        if self.codec is not None:
            self.audio_alignment = int(_arg(args, "audio_alignment", a))
            self.vq_groups = int(_arg(args, "vq_groups", g))
            self.audio_vocab_size = int(_arg(args, "audio_vocab_size", v))
            self.audio_weight = float(_arg(args, "audio_weight", 10.0))
        else:
            self.audio_alignment = self.vq_groups = self.audio_vocab_size = 0
            self.audio_weight = 0.0
        self._codec_fn = None
        self._lmax = int(_arg(args, "max_label_len", 64)) + 1

        from ._engines import EngineCache

        self._engines = EngineCache("svsr_lrs", self.device_)
        self._ent = None
        self._native_updates = 0
        self.sync_grads = True  # see lightning.TransformerLightningModule.sync_grads
        self._h = C.c_void_p()
        self._shape_key = None
        self._flat_p = self._flat_g = self._flat_b = self._ws = None
        self._metrics = torch.zeros(8, device=self.device_, dtype=torch.float32)
        self._anchor = torch.zeros((), device=self.device_, requires_grad=True)
        self._param_views: Dict[str, nn.Parameter] = {}
        self._offsets: Dict[str, tuple] = {}
        self.encoder = _EncoderNode(self)
        self._build_engine(B=1, T=8, H=88, W=88, first=True)
        self.criterion = _Node()  # LabelSmoothingLoss has no parameters; the loss itself is fused into the step

    # ------------------------------------------------------------------------------------------------------------
    def _engine_cfg(self, B, T, H, W) -> LrsConfig:
        return LrsConfig(B, T, H, W, self.adim, self.aheads, self.eunits, self.elayers, self.dlayers, self.dunits,
                         self.odim, self.cnn_kernel, self.audio_alignment, self.vq_groups, self.audio_vocab_size,
                         self._lmax, self.mtlalpha, self.lsm_weight, self.audio_weight, 1e-5, 0.1,
                         self.dropout_rate, self.attn_dropout_rate)

    def _build_engine(self, B, T, H, W, first=False):
        """Selects (building it on first use) the engine of this clip geometry; see _engines.EngineCache. The LRS
        datamodule pads every batch to its own longest clip (datamodule/data_module.py:12-43), so T changes almost every
        step: engines already built are kept (LRU, SVSR_ENGINE_CACHE_GB) instead of rebuilt."""
        L = lib()
        for f in ("param_count", "buffer_count", "workspace_bytes", "decay_count"):
            getattr(L, f"svsr_lrs_{f}").restype = C.c_int64
        key = (B, T, H, W, self._lmax)
        ent = self._engines.get(key)
        if ent is None:
            def bind(h, ws_ptr, ws_bytes):
                check(L.svsr_lrs_bind(h, C.c_void_p(self._flat_p.data_ptr()), C.c_void_p(self._flat_g.data_ptr()),
                                      C.c_void_p(self._flat_b.data_ptr()), C.c_void_p(ws_ptr), C.c_int64(ws_bytes)),
                      "svsr_lrs_bind")

            def arenas(h):
                self._h = h
                self._create_arenas()

            ent = self._engines.create(key, self._engine_cfg(B, T, H, W), bind, arenas if first else None)
        self._ent, self._h, self._ws = ent, ent.h, ent.ws
        self._engine_gen = ent.id
        self._shape_key = (B, T, H, W)

    def _create_arenas(self):
        L = lib()
        h = self._h
        n_p, n_b = L.svsr_lrs_param_count(h), L.svsr_lrs_buffer_count(h)
        self._flat_p = torch.zeros(n_p, device=self.device_)
        self._flat_g = torch.zeros(n_p, device=self.device_)
        self._flat_b = torch.zeros(n_b, device=self.device_)
        self.n_decay = int(L.svsr_lrs_decay_count(h))
        name, ndim, off, decay = C.c_char_p(), C.c_int(), C.c_int64(), C.c_int()
        shape = (C.c_int64 * 5)()
        gen = torch.Generator(device="cpu").manual_seed(torch.initial_seed() & 0x7FFFFFFF)
        for i in range(L.svsr_lrs_num_params(h)):
            check(L.svsr_lrs_param_info(h, i, C.byref(name), C.byref(ndim), shape, C.byref(off), C.byref(decay)), "info")
            key, shp = name.value.decode(), tuple(shape[k] for k in range(ndim.value))
            n = math.prod(shp)
            view = self._flat_p[off.value: off.value + n].view(shp)
            wkey = key.rsplit(".", 1)[0] + ".weight"
            fan_in = math.prod(self._offsets[wkey][2][1:]) if (key.endswith(".bias") and wkey in self._offsets) else 0
            self._init_param(key, view, gen, fan_in)
            p = nn.Parameter(view)
            self._register(key, p, is_buffer=False)
            self._param_views[key] = p
            self._offsets[key] = (off.value, n, shp, bool(decay.value))
        for i in range(L.svsr_lrs_num_buffers(h)):
            check(L.svsr_lrs_buffer_info(h, i, C.byref(name), C.byref(ndim), shape, C.byref(off)), "info")
            key, shp = name.value.decode(), tuple(shape[k] for k in range(ndim.value))
            view = self._flat_b[off.value: off.value + shp[0]].view(shp)
            if key.endswith("running_var"):
                view.fill_(1.0)
            self._register(key, view, is_buffer=True)
        bn_keys = [k for k in dict(self.named_buffers()) if k.endswith("running_var")]
        self._nbt = torch.zeros(len(bn_keys), dtype=torch.long, device=self.device_)
        for i, key in enumerate(bn_keys):
            self._register(key.replace("running_var", "num_batches_tracked"), self._nbt[i], is_buffer=True)
        self._attach_grads()

    @staticmethod
    def _init_param(key: str, view: torch.Tensor, gen: torch.Generator, fan_in_of_weight: int = 0) -> None:
        """PyTorch-default initialisation, as in the reference (E2E calls no custom init): Conv/Linear
        kaiming-uniform(a=sqrt 5) = U(+-1/sqrt(fan_in)) for weight and bias, BatchNorm/LayerNorm 1/0, Embedding N(0,1),
        pos_bias_u/v xavier-uniform (transformer/attention.py:213-214)."""
        shp = tuple(view.shape)
        leaf = key.rsplit(".", 1)[-1]
        is_norm = any(t in key for t in (".bn1.", ".bn2.", "downsample.1.", "frontend3D.1.", ".norm", "after_norm"))
        if is_norm:
            view.fill_(1.0) if leaf == "weight" else view.zero_()
        elif "pos_bias_" in key:
            bound = math.sqrt(6.0 / (shp[0] + shp[1]))
            view.copy_((torch.rand(shp, generator=gen) * 2 - 1) * bound)
        elif key == "decoder.embed.0.weight":
            view.copy_(torch.randn(shp, generator=gen))
        elif leaf == "weight":
            fan_in = math.prod(shp[1:])
            view.copy_((torch.rand(shp, generator=gen) * 2 - 1) / math.sqrt(fan_in))
        else:  # bias of a Linear / Conv1d: U(+-1/sqrt(fan_in of its weight))
            view.copy_((torch.rand(shp, generator=gen) * 2 - 1) / math.sqrt(fan_in_of_weight or 1))

    def _register(self, key: str, value, is_buffer: bool) -> None:
        parts = key.split(".")
        node: nn.Module = self
        for part in parts[:-1]:
            if part not in node._modules:
                node.add_module(part, _Node())
            node = node._modules[part]
        if is_buffer:
            node.register_buffer(parts[-1], value)
        else:
            node.register_parameter(parts[-1], value)

    def _attach_grads(self) -> None:
        for key, p in self._param_views.items():
            off, n, shp, _ = self._offsets[key]
            p.grad = self._flat_g[off: off + n].view(shp)

    @property
    def flat_params(self) -> torch.Tensor:
        return self._flat_p

    @property
    def flat_grads(self) -> torch.Tensor:
        return self._flat_g

    def mark_weights_updated(self) -> None:
        """See lightning.TransformerLightningModule.mark_weights_updated (torch-side updates are detected)."""
        self._native_updates += 1

    @property
    def _weights_dirty(self) -> bool:
        e = self._ent
        return e is None or e.packed_native != self._native_updates or e.packed_version != self._flat_p._version

    @_weights_dirty.setter
    def _weights_dirty(self, dirty: bool) -> None:
        if dirty:
            self._native_updates += 1
        elif self._ent is not None:
            self._ent.packed_native, self._ent.packed_version = self._native_updates, self._flat_p._version

    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _stream() -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _ensure(self, x: torch.Tensor, label_len: int = 0) -> None:
        if x.dim() != 5 or x.shape[2] != 1:
            raise ValueError(f"x must be [B,T,1,H,W], got {tuple(x.shape)}")
        B, T, _, H, W = x.shape
        if label_len + 1 > self._lmax:
            self._lmax = label_len + 1
            self._shape_key = None
        if self._shape_key != (B, T, H, W):
            self._build_engine(B, T, H, W)
        if self._weights_dirty:
            check(lib().svsr_lrs_pack_weights(self._h, self._stream()), "svsr_lrs_pack_weights")
            self._weights_dirty = False

    def _named_tensor(self, name: str, shape=None) -> torch.Tensor:
        ptr, numel, dt = C.c_void_p(), C.c_int64(), C.c_int()
        check(lib().svsr_lrs_tensor(self._h, name.encode(), C.byref(ptr), C.byref(numel), C.byref(dt)), "svsr_lrs_tensor")
        tdt = {0: torch.float32, 1: torch.bfloat16, 2: torch.uint8, 3: torch.int32, 4: torch.int64}[dt.value]
        esz = {0: 4, 1: 2, 2: 1, 3: 4, 4: 8}[dt.value]
        off = ptr.value - self._ws.data_ptr()
        flat = self._ws[off: off + numel.value * esz].view(tdt)
        return flat.view(shape) if shape is not None else flat

    def _step_seed(self) -> int:
        if not self.training or (self.dropout_rate == 0.0 and self.attn_dropout_rate == 0.0):
            return 0
        return self.dropout_seed if self.dropout_seed is not None else random.getrandbits(63)

    def _apply_step_seed(self, seed: Optional[int]) -> None:
        """Device-resident step seed (svsr_lrs_step_control): seed = this step's dropout seed written to the engine's
        control word on the current stream -- every dropout site then reads it from device memory, which is what lets
        train.SentenceDataParallelStep replay ONE CUDA graph under the shipped dropout config. None: host-valued again."""
        check(lib().svsr_lrs_step_control(self._h, C.c_int(0 if seed is None else 1), C.c_uint64(seed or 0), self._stream()),
              "svsr_lrs_step_control")
        if self._ent is not None:
            self._ent.dev_seed = seed is not None

    def _seed_arg(self) -> int:
        """The step's dropout seed. An engine switched to the device-resident seed ignores the argument: a step launched
        kernel by kernel on such an engine refreshes the device word first (never inside a graph capture, where the
        caller sets it per replay)."""
        dev = self._ent is not None and self._ent.dev_seed
        if dev and torch.cuda.is_current_stream_capturing():
            return 0  # ignored by the engine; nothing is drawn, so the host RNG stream stays one draw per step
        seed = self._step_seed()
        if dev:
            self._apply_step_seed(seed)
        return seed

    def attach_codec(self, fn) -> None:
        """`fn(audios [B, samples]) -> int64 tokens [B, Ta, G]`: the frozen neural audio quantiser of
        e2e_asr_transformer.py:167-180 (wav2vec 2.0 / vq-wav2vec). It is off the gradient path and needs pretrained
        weights, so it stays the caller's PyTorch module; pre-tokenised `audios` bypass it."""
        self._codec_fn = fn

    def forward_audios(self, audios: torch.Tensor) -> torch.Tensor:
        if self._codec_fn is None:
            raise SvsrError("raw waveforms were passed but no quantiser is attached: call attach_codec(fn) or pass "
                            "int64 audio tokens [B, >=T*alignment, groups] as `audios`")
        with torch.no_grad():
            return self._codec_fn(audios)

    def _encode(self, xs: torch.Tensor, masks: Optional[torch.Tensor], extract_resnet_feats: bool = False):
        xs = xs.to(self.device_, torch.float32).contiguous()
        self._ensure(xs)
        B, T = xs.shape[:2]
        lengths = None
        if masks is not None:
            lengths = masks.to(self.device_).reshape(B, -1).sum(-1).long().contiguous()
        check(lib().svsr_lrs_encode(self._h, C.c_void_p(xs.data_ptr()),
                                    C.c_void_p(lengths.data_ptr() if lengths is not None else 0),
                                    C.c_int(int(self.training)), C.c_uint64(self._seed_arg()), self._stream()),
              "svsr_lrs_encode")
        if self.training:
            self._nbt += 1
        if extract_resnet_feats:
            return self._named_tensor("frontend", (B, T, 512)).float()
        return self._named_tensor("encoder_out", (B, T, self.adim)).clone(), masks

    def forward(self, x: torch.Tensor, lengths: torch.Tensor, audios: Optional[torch.Tensor], label: torch.Tensor):
        x = x.to(self.device_, torch.float32).contiguous()
        lengths = lengths.to(self.device_).long().contiguous()
        label = label.to(self.device_).long().contiguous()
        if label.dim() != 2 or label.shape[0] != x.shape[0]:
            raise ValueError(f"label must be [B, Lmax] padded with -1, got {tuple(label.shape)}")
        self._ensure(x, int(label.shape[1]))
        B, T = x.shape[:2]
        tokens = None
        if self.codec is not None:
            if audios is None:
                raise ValueError("codec is set: `audios` (waveforms or pre-quantised int64 tokens) is required")
            if audios.dtype in (torch.long, torch.int32) and audios.dim() == 3:
                tokens = audios.to(self.device_).long().contiguous()
            else:  # the reference's call convention: audios [1?, B, samples] -> permute/squeeze (e2e...py:197)
                tokens = self.forward_audios(audios.permute(1, 0, 2).squeeze(0)).to(self.device_).long().contiguous()
            if tokens.shape[2] != self.vq_groups or tokens.shape[1] < T * self.audio_alignment:
                raise ValueError(f"audio tokens must be [B, >={T * self.audio_alignment}, {self.vq_groups}]")
        check(lib().svsr_lrs_forward(
            self._h, C.c_void_p(x.data_ptr()), C.c_void_p(lengths.data_ptr()),
            C.c_void_p(tokens.data_ptr() if tokens is not None else 0),
            C.c_int64(tokens.stride(0) if tokens is not None else 0), C.c_void_p(label.data_ptr()),
            C.c_int(int(label.shape[1])), C.c_int(int(self.training)), C.c_uint64(self._seed_arg()),
            C.c_void_p(self._metrics.data_ptr()), self._stream()), "svsr_lrs_forward")
        self._last_BL = (B, int(label.shape[1]) + 1)
        if self.training:
            self._nbt += 1
        if torch.is_grad_enabled() and self.training:
            m = _StepFunction.apply(self._anchor, self, self._metrics)
        else:
            m = self._metrics.clone()  # the metrics buffer is overwritten by the next step
        loss_audio = m[3] if self.codec is not None else None
        # (loss, loss_ctc, loss_att, loss_audio, acc) -- acc stays a device scalar (the reference returns a Python
        # float, which costs a host sync per step; float(acc) gives the same number)
        return m[0], m[1], m[2], loss_audio, m[4]

    def _native_backward(self, grad_metrics: torch.Tensor) -> None:
        g = grad_metrics.contiguous()
        need_attach = any(p.grad is None for p in self._param_views.values())
        if need_attach:
            self._flat_g.zero_()
        check(lib().svsr_lrs_backward(self._h, C.c_void_p(g.data_ptr()), self._stream()), "svsr_lrs_backward")
        allreduce_mean_(self._flat_g, self.sync_grads)
        if need_attach:
            self._attach_grads()

    def scorers(self):
        raise SvsrError("beam-search scorers (inference) are outside the native training hot path")

    # named intermediate tensors (parity tests)
    def encoder_out(self) -> torch.Tensor:
        B, T, _, _ = self._shape_key
        return self._named_tensor("encoder_out", (B, T, self.adim)).clone()

    def logits_audio(self) -> torch.Tensor:
        B, T, _, _ = self._shape_key
        check(lib().svsr_lrs_logits_audio(self._h, self._stream()), "svsr_lrs_logits_audio")  # not written by the step
        return self._named_tensor("logits_audio", (B, T, -1)).clone()

    def ctc_logits(self) -> torch.Tensor:
        B, T, _, _ = self._shape_key
        return self._named_tensor("ctc_logits", (B, T, -1))[:, :, : self.odim].clone()

    def pred(self) -> torch.Tensor:
        B, L = self._last_BL
        return self._named_tensor("pred", (B, L, -1))[:, :, : self.odim].clone()

    def __del__(self):
        try:
            self._engines.destroy()
        except Exception:
            pass
