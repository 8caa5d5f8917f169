"""Host-side mirror of the reference's LRW module (/root/reference/LRW/video/src/lightning.py:36-223).

`TransformerLightningModule(config)` keeps the reference's constructor argument, attribute names
(`stem3d`, `resnet.layer1..4`, `encoder`, `audio_projection`, `category_classifier`, `cls_token`, `lambda_audio`,
`audio_alignment`, `vq_groups`, `audio_vocab_size`, `codec`), `forward_videos`, `forward(videos, audio_tokens,
labels, word_mask) -> dict`, `training_step/validation_step/test_step`, `configure_optimizers` and state-dict keys,
so the reference's training loop can construct it and load the reference's checkpoints. All arithmetic runs in the
sm_100a kernels of libsvsr.so through the native step executor (csrc/engine.cu); PyTorch only owns the memory
(three flat fp32 arenas + one workspace) and the autograd hook that makes `loss_total.backward()` work.
There is no CPU / eager fallback: without the shared library or a CUDA device construction fails.
"""
from __future__ import annotations

import ctypes as C
import random
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from ._lib import SvsrError, check, lib


class LrwConfig(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("T", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("dim", C.c_int), ("depth", C.c_int), ("heads", C.c_int),
        ("audio_alignment", C.c_int), ("vq_groups", C.c_int), ("audio_vocab", C.c_int),
        ("num_labels", C.c_int), ("rotary_v", C.c_int),
        ("lambda_audio", C.c_float), ("label_smoothing", C.c_float),
        ("bn_eps", C.c_float), ("bn_momentum", C.c_float), ("ff_dropout", C.c_float),
        ("enc_type", C.c_int), ("bert_intermediate", C.c_int), ("bert_max_pos", C.c_int),
        ("bert_ln_eps", C.c_float), ("bert_hidden_dropout", C.c_float), ("bert_attn_dropout", C.c_float),
        ("emb_dropout", C.c_float), ("attn_dropout", C.c_float),
    ]


def _cfg_get(cfg: Any, path: str, default: Any = None) -> Any:
    cur = cfg
    for key in path.split("."):
        if cur is None:
            return default
        if isinstance(cur, dict):
            cur = cur.get(key, None)
        else:
            cur = getattr(cur, key, None)
    return default if cur is None else cur


def allreduce_mean_(flat: torch.Tensor, enabled: bool = True, group=None) -> torch.Tensor:
    """What DDP's reducer does for the reference (`Trainer(strategy="ddp")`, LRW/video/src/train.py:28): average the
    gradients over the ranks -- here ONE all-reduce over the flat arena. No-op without an initialised process group."""
    import torch.distributed as dist

    if enabled and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(dist.get_world_size(group))
    return flat


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's module tree (and therefore its state-dict keys)."""


class _StepFunction(torch.autograd.Function):
    """Makes `metrics['loss_total'].backward()` run the native backward. Gradients are accumulated straight into the
    flat gradient arena that every parameter's `.grad` is a view of, so nothing is returned through autograd."""

    @staticmethod
    def forward(ctx, anchor: torch.Tensor, module: "TransformerLightningModule", metrics: torch.Tensor):
        ctx.module = module
        return metrics.clone()

    @staticmethod
    def backward(ctx, grad_metrics: torch.Tensor):
        ctx.module._native_backward(grad_metrics)
        return torch.zeros((), device=grad_metrics.device), None, None


class TransformerLightningModule(nn.Module):
    def __init__(self, config: Any, device: Optional[torch.device | str] = None):
        super().__init__()
        if not torch.cuda.is_available():
            raise SvsrError("syncvsr_b200 needs a CUDA (sm_100a) device: the hot path has no CPU fallback")
        lib()  # fail loudly now if libsvsr.so is missing
        self.config = config
        self.device_ = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.is_train = False
        self.word_labels = int(_cfg_get(config, "model.bert.num_labels", 500))
        self.lambda_audio = float(_cfg_get(config, "optim.lambda_audio", 10.0))
        self.label_smoothing = float(_cfg_get(config, "train.label_smoothing", 0.0))
        self.use_wb = bool(_cfg_get(config, "data.use_word_boundary", False))
        self.encoder_type = str(_cfg_get(config, "model.bert.type", "x-transformers"))
        if self.encoder_type not in ("x-transformers", "huggingface"):
            raise SvsrError(f"model.bert.type={self.encoder_type!r}: only 'x-transformers' and 'huggingface' exist")

        # codec constants: explicit keys win, else derived from the codec path exactly like lightning.py:58-67
        path = str(_cfg_get(config, "model.wav2vec.path", "vq"))
        if "vq" in path:
            self.codec, a, g, v = "vq", 4, 2, 320
        elif "wav2vec2" in path:
            self.codec, a, g, v = "wav2vec2", 2, 2, 640
        else:
            raise SvsrError(f"cannot derive the audio codec from model.wav2vec.path={path!r}")
        self.audio_alignment = int(_cfg_get(config, "model.audio_alignment", a))
        self.vq_groups = int(_cfg_get(config, "model.vq_groups", g))
        self.audio_vocab_size = int(_cfg_get(config, "model.audio_vocab_size", v))
        # lightning.py:46-47: the word-boundary indicator becomes one extra hidden channel (dim 513)
        self.dim = int(_cfg_get(config, "model.bert.dim", 512)) + (1 if self.use_wb else 0)
        self.hf: Dict[str, Any] = {}
        if self.encoder_type == "huggingface":
            # lightning.py:90-92: BertModel(BertConfig(**config.model.bert)); BertConfig defaults apply to absent keys
            g = lambda k, d: _cfg_get(config, f"model.bert.{k}", d)  # noqa: E731
            self.hf = dict(hidden_size=int(g("hidden_size", 768)), num_hidden_layers=int(g("num_hidden_layers", 12)),
                           num_attention_heads=int(g("num_attention_heads", 12)),
                           intermediate_size=int(g("intermediate_size", 3072)), vocab_size=int(g("vocab_size", 30522)),
                           max_position_embeddings=int(g("max_position_embeddings", 512)),
                           layer_norm_eps=float(g("layer_norm_eps", 1e-12)),
                           hidden_dropout_prob=float(g("hidden_dropout_prob", 0.1)),
                           attention_probs_dropout_prob=float(g("attention_probs_dropout_prob", 0.1)),
                           hidden_act=str(g("hidden_act", "gelu")), type_vocab_size=int(g("type_vocab_size", 2)))
            if self.use_wb or self.hf["hidden_size"] != 512 or self.hf["hidden_size"] != 64 * self.hf["num_attention_heads"]:
                raise SvsrError("huggingface encoder: hidden_size must be 512 (the trunk's width, no word boundary) with "
                                "heads of 64; other BertConfig geometries cannot consume forward_videos() either")
            if self.hf["hidden_act"] != "gelu" or self.hf["type_vocab_size"] != 2:
                raise SvsrError("huggingface encoder: only hidden_act='gelu' and type_vocab_size=2 are native")
            self.dim = 512
        self._dim_pitch = (self.dim + 63) // 64 * 64  # row pitch of the engine's dim-wide tensors
        self.depth = int(_cfg_get(config, "model.bert.depth", 12))
        self.heads = int(_cfg_get(config, "model.bert.heads", 8))
        if self.hf:
            self.depth, self.heads = self.hf["num_hidden_layers"], self.hf["num_attention_heads"]
        self.layer_dropout = float(_cfg_get(config, "model.bert.layer_dropout", 0.0))
        self.ff_dropout = float(_cfg_get(config, "model.bert.ff_dropout", 0.0))
        # Dropout(emb_dropout) on cat(cls, inputs_embeds) (lightning.py:106,150) and x-transformers' attn_dropout on the
        # attention probabilities: both 0.0 in the shipped yamls, honoured in training mode when set
        self.emb_dropout = float(_cfg_get(config, "model.bert.emb_dropout", 0.0))
        self.attn_dropout = 0.0 if self.hf else float(_cfg_get(config, "model.bert.attn_dropout", 0.0))
        self.rotary_v = bool(_cfg_get(config, "model.bert.rotary_v", True))

        from ._engines import EngineCache

        self._engines = EngineCache("svsr_lrw", self.device_)
        self._ent = None
        self._native_updates = 0  # bumped by mark_weights_updated(): parameter writes torch cannot see (raw pointers)
        self._h = C.c_void_p()
        self._shape_key = None
        self._flat_p = self._flat_g = self._flat_b = self._ws = None
        self._metrics = torch.zeros(8, device=self.device_, dtype=torch.float32)
        self._anchor = torch.zeros((), device=self.device_, requires_grad=True)
        self._last_skip = 0  # layer_dropout mask of the last forward (bit i = sublayer i dropped); read by FusedAdamW
        # device_control: the step's layer_dropout mask / dropout seed live in device memory and every sublayer is
        # launched predicated on them (svsr_lrw_step_control) -- what lets train.DataParallelStep replay ONE CUDA graph
        # under the shipped layer_dropout / ff_dropout config. Off: they are kernel arguments (host-side skipping).
        self.device_control = False
        self._ctl_preset = None
        # Gradient synchronisation: gradients are written straight into the flat arena (no autograd graph through the
        # parameters), so torch DDP's reducer never sees them. Under torchrun (process group initialised) the native
        # backward all-reduces the arena itself (SUM / world = DDP's mean); DataParallelStep does its own staged
        # all-reduce and calls the backward entry points directly. Do NOT wrap the module in DistributedDataParallel.
        self.sync_grads = True
        self._param_views: Dict[str, nn.Parameter] = {}
        self._offsets: Dict[str, tuple] = {}
        # geometry-independent part: parameter arenas. Built with a nominal clip geometry (the parameter layout
        # does not depend on B/H/W); the workspace is (re)built lazily for the geometry actually fed to forward().
        self._build_engine(B=1, T=29, H=88, W=88, first=True)

        from .augment import CutMix  # on-device mirror of augment.py (pre-quantised audio tokens: wav2vec = None)

        self.cutmix = CutMix(self.word_labels, None).eval()

        # parameters the reference's state dict carries but never uses on this path (lightning.py:55 creates the full
        # timm resnet18; only .layer1-4 run): kept for checkpoint compatibility, never receive gradients.
        if self.hf:  # BertModel members that forward(inputs_embeds=...).last_hidden_state never touches
            emb = self.encoder._modules["embeddings"]
            emb.word_embeddings = nn.Embedding(self.hf["vocab_size"], 512, padding_idx=0).to(self.device_)
            self.encoder.pooler = _Node()
            self.encoder.pooler.dense = nn.Linear(512, 512).to(self.device_)
        rn = self.resnet
        rn.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        rn.bn1 = nn.BatchNorm2d(64)
        rn.fc = nn.Linear(512, 1000)
        for m in (rn.conv1, rn.bn1, rn.fc):
            m.to(self.device_)

    # ------------------------------------------------------------------------------------------------------------
    # engine / arena management
    # ------------------------------------------------------------------------------------------------------------
    def _engine_cfg(self, B, T, H, W) -> LrwConfig:
        hf = self.hf
        return LrwConfig(B, T, H, W, self.dim, self.depth, self.heads, self.audio_alignment, self.vq_groups,
                         self.audio_vocab_size, self.word_labels, int(self.rotary_v), self.lambda_audio,
                         self.label_smoothing, 1e-5, 0.1, self.ff_dropout, 1 if hf else 0,
                         hf.get("intermediate_size", 0), hf.get("max_position_embeddings", 0),
                         hf.get("layer_norm_eps", 1e-12), hf.get("hidden_dropout_prob", 0.0),
                         hf.get("attention_probs_dropout_prob", 0.0), self.emb_dropout, self.attn_dropout)

    def _build_engine(self, B, T, H, W, first=False):
        """Selects (building it on first use) the engine of this clip geometry; see _engines.EngineCache."""
        L = lib()
        L.svsr_lrw_param_count.restype = C.c_int64
        L.svsr_lrw_buffer_count.restype = C.c_int64
        key = (B, T, H, W)
        ent = self._engines.get(key)
        if ent is None:
            def bind(h, ws_ptr, ws_bytes):
                check(L.svsr_lrw_bind(h, C.c_void_p(self._flat_p.data_ptr()), C.c_void_p(self._flat_g.data_ptr()),
                                      C.c_void_p(self._flat_b.data_ptr()), C.c_void_p(ws_ptr), C.c_int64(ws_bytes)),
                      "svsr_lrw_bind")

            def arenas(h):
                self._h = h
                self._create_arenas()

            ent = self._engines.create(key, self._engine_cfg(B, T, H, W), bind, arenas if first else None)
        self._ent, self._h, self._ws = ent, ent.h, ent.ws
        self._engine_gen = ent.id  # CUDA graphs are captured against one engine (train.DataParallelStep keys them by it)
        self._shape_key = key

    def _create_arenas(self):
        L = lib()
        h = self._h
        n_p, n_b = L.svsr_lrw_param_count(h), L.svsr_lrw_buffer_count(h)
        self._flat_p = torch.zeros(n_p, device=self.device_)
        self._flat_g = torch.zeros(n_p, device=self.device_)
        self._flat_b = torch.zeros(n_b, device=self.device_)
        name, ndim, off, decay = C.c_char_p(), C.c_int(), C.c_int64(), C.c_int()
        shape = (C.c_int64 * 5)()
        gen = torch.Generator(device="cpu").manual_seed(torch.initial_seed() & 0x7FFFFFFF)
        self._decay_mask_segments = []
        for i in range(L.svsr_lrw_num_params(h)):
            check(L.svsr_lrw_param_info(h, i, C.byref(name), C.byref(ndim), shape, C.byref(off), C.byref(decay)), "info")
            key, shp = name.value.decode(), tuple(shape[k] for k in range(ndim.value))
            n = 1
            for s in shp:
                n *= s
            view = self._flat_p[off.value: off.value + n].view(shp)
            self._init_param(key, view, gen)
            if key == "cls_token" and self.use_wb:
                view[0, 0, -1] = 0.0  # [CLS] is not part of word_mask (lightning.py:109-110)
            p = nn.Parameter(view)
            self._register(key, p, is_buffer=False)
            self._param_views[key] = p
            self._offsets[key] = (off.value, n, shp, bool(decay.value))
        for i in range(L.svsr_lrw_num_buffers(h)):
            check(L.svsr_lrw_buffer_info(h, i, C.byref(name), C.byref(ndim), shape, C.byref(off)), "info")
            key, shp = name.value.decode(), tuple(shape[k] for k in range(ndim.value))
            view = self._flat_b[off.value: off.value + shp[0]].view(shp)
            if key.endswith("running_var"):
                view.fill_(1.0)
            self._register(key, view, is_buffer=True)
        # reference BN modules also carry num_batches_tracked: 0-dim views of one counter vector (one add per step)
        bn_keys = [k for k in dict(self.named_buffers()) if k.endswith("running_var")]
        self._nbt = torch.zeros(len(bn_keys), dtype=torch.long, device=self.device_)
        for i, key in enumerate(bn_keys):
            self._register(key.replace("running_var", "num_batches_tracked"), self._nbt[i], is_buffer=True)
        self._attach_grads()

    @staticmethod
    def _init_param(key: str, view: torch.Tensor, gen: torch.Generator) -> None:
        """Reference-equivalent default initialisation (torchvision/timm kaiming-normal fan_out convs, BN 1/0,
        PyTorch-default Linear, randn CLS, RMSNorm g = 1)."""
        import math

        shp = view.shape
        if key.startswith(("encoder.embeddings.", "encoder.encoder.layer.")):  # BertPreTrainedModel._init_weights
            if "LayerNorm" in key:
                view.fill_(1.0) if key.endswith("weight") else view.zero_()
            elif key.endswith("bias"):
                view.zero_()
            else:
                view.copy_(torch.randn(shp, generator=gen) * 0.02)
            return
        is_bn = ".bn" in key or "stem3d.1" in key or "downsample.1" in key
        if key.endswith(".g") or (is_bn and key.endswith("weight")):
            view.fill_(1.0)
        elif is_bn and key.endswith("bias"):
            view.zero_()
        elif key == "cls_token":
            view.copy_(torch.randn(shp, generator=gen))
        elif view.dim() >= 4:  # conv: kaiming normal, mode=fan_out, relu
            fan_out = shp[0]
            for s in shp[2:]:
                fan_out *= s
            view.copy_(torch.randn(shp, generator=gen) * math.sqrt(2.0 / fan_out))
        elif view.dim() == 2:  # nn.Linear default: U(-1/sqrt(in), 1/sqrt(in))
            bound = 1.0 / math.sqrt(shp[1])
            view.copy_((torch.rand(shp, generator=gen) * 2 - 1) * bound)
        elif view.dim() == 1:  # Linear bias
            fan_in = 4 * 512 if "ff.3" in key else 512
            bound = 1.0 / math.sqrt(fan_in)
            view.copy_((torch.rand(shp, generator=gen) * 2 - 1) * bound)

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

    # flat views used by the data-parallel step and the fused optimizer
    @property
    def flat_params(self) -> torch.Tensor:
        return self._flat_p

    @property
    def flat_grads(self) -> torch.Tensor:
        return self._flat_g

    def mark_weights_updated(self) -> None:
        """Parameters were written through raw pointers (the native optimizer): every engine's bf16 operand copies are
        stale and are repacked lazily. Torch-side writes (torch.optim.*, load_state_dict, p.data.add_) need no call:
        they bump the arena's autograd version counter, which _ensure() compares."""
        self._native_updates += 1

    @property
    def _weights_dirty(self) -> bool:
        e = self._ent
        return e is None or e.packed_native != self._native_updates or e.packed_version != self._flat_p._version

    @_weights_dirty.setter
    def _weights_dirty(self, dirty: bool) -> None:
        if dirty:
            self._native_updates += 1
        elif self._ent is not None:  # (a replayed CUDA graph contains the repack of the current engine)
            self._ent.packed_native, self._ent.packed_version = self._native_updates, self._flat_p._version

    # ------------------------------------------------------------------------------------------------------------
    # reference API
    # ------------------------------------------------------------------------------------------------------------
    def _ensure(self, videos: torch.Tensor) -> None:
        if videos.dim() != 5 or videos.shape[1] != 1:
            raise ValueError(f"videos must be [B,1,T,H,W], got {tuple(videos.shape)}")
        B, _, T, H, W = videos.shape
        if self._shape_key != (B, T, H, W):
            self._build_engine(B, T, H, W)
        # The bf16 operand copies go stale whenever the fp32 arena changes. The native optimizer (raw pointers) says so
        # through mark_weights_updated(); any torch-side in-place update of a parameter view (torch.optim.*, the
        # optimizers configure_optimizers() returns, manual `p.data.add_`) bumps the arena's autograd version counter,
        # which every view shares.
        if self._weights_dirty:
            check(lib().svsr_lrw_pack_weights(self._h, self._stream()), "svsr_lrw_pack_weights")
            self._weights_dirty = False

    @staticmethod
    def _stream() -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _named_tensor(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        ptr, numel, dt = C.c_void_p(), C.c_int64(), C.c_int()
        check(lib().svsr_lrw_tensor(self._h, name.encode(), C.byref(ptr), C.byref(numel), C.byref(dt)), "svsr_lrw_tensor")
        tdt = {0: torch.float32, 1: torch.bfloat16, 2: torch.uint8, 3: torch.int32}[dt.value]
        esz = {0: 4, 1: 2, 2: 1, 3: 4}[dt.value]
        off = ptr.value - self._ws.data_ptr()
        flat = self._ws[off: off + numel.value * esz].view(tdt)
        return flat.view(shape) if shape is not None else flat

    def forward_videos(self, videos: torch.Tensor) -> torch.Tensor:
        """lightning.py:112-119 -> [B, T, 512] (fp32 copy of the pooled trunk features)."""
        videos = videos.to(self.device_, torch.float32).contiguous()
        self._ensure(videos)
        B, _, T, _, _ = videos.shape
        check(lib().svsr_lrw_forward_videos(self._h, C.c_void_p(videos.data_ptr()), C.c_int(int(self.training)),
                                            self._stream()), "svsr_lrw_forward_videos")
        return self._named_tensor("inputs_embeds", (B, T + 1, self._dim_pitch))[:, 1:, :512].clone()

    def forward(self, videos: torch.Tensor, audio_tokens: torch.Tensor, labels: torch.Tensor,
                word_mask: torch.Tensor) -> Dict[str, torch.Tensor]:
        videos = videos.to(self.device_, torch.float32).contiguous()
        audio_tokens = audio_tokens.to(self.device_, torch.long).contiguous()
        labels = labels.to(self.device_)
        self._ensure(videos)
        B, _, T, _, _ = videos.shape
        if audio_tokens.dim() != 3 or audio_tokens.shape[2] != self.vq_groups:
            raise ValueError(f"audio_tokens must be [B, >=T*{self.audio_alignment}, {self.vq_groups}]")
        if audio_tokens.shape[1] < T * self.audio_alignment:
            raise ValueError("audio_tokens has fewer than seq_len * audio_alignment rows")
        wm = None
        if self.use_wb:  # [B, T] 0/1 indicator of the frames inside the target word (data.py:58-64)
            wm = word_mask.to(self.device_, torch.float32).contiguous()
            if wm.shape != (B, T):
                raise ValueError(f"word_mask must be [B, T] = {(B, T)} with data.use_word_boundary, got {tuple(wm.shape)}")
        hard = soft = None
        if labels.dtype in (torch.long, torch.int32, torch.int64):
            hard = labels.long().contiguous()
        else:  # CutMix soft labels [B, num_labels] (augment.py; lightning.py:163-165,177-179)
            soft = labels.float().contiguous()
        if self._ctl_preset is not None:  # train.DataParallelStep already wrote this step's control words (graph mode)
            skip, seed = self._ctl_preset
        else:
            skip, seed = self._draw_step_control()
            if self.device_control:
                self._apply_step_control(skip, seed)
        self._last_skip = skip
        check(lib().svsr_lrw_forward(
            self._h, C.c_void_p(videos.data_ptr()), C.c_void_p(audio_tokens.data_ptr()),
            C.c_int64(audio_tokens.stride(0)), C.c_void_p(hard.data_ptr() if hard is not None else 0),
            C.c_void_p(soft.data_ptr() if soft is not None else 0), C.c_void_p(wm.data_ptr() if wm is not None else 0),
            C.c_int(int(self.training)), C.c_uint32(skip),
            C.c_uint64(seed),
            C.c_void_p(self._metrics.data_ptr()), self._stream()), "svsr_lrw_forward")
        self._precise_logits = False
        if self.training:
            self._nbt += 1
        if torch.is_grad_enabled():
            m = _StepFunction.apply(self._anchor, self, self._metrics)
        else:  # the persistent buffer is overwritten by the next forward: hand out a copy (epoch averages collect these)
            m = self._metrics.clone()
        return {"loss_total": m[0], "loss_category": m[1], "loss_audio": m[2], "accuracy_top1": m[3],
                "accuracy_top5": m[4]}

    def _draw_step_control(self):
        """This step's layer_dropout mask (bit i = sublayer i dropped; host RNG per sublayer, like x-transformers'
        `random() < layer_dropout`) and dropout seed. Eval mode: (0, 0)."""
        skip = 0
        if self.training and self.layer_dropout > 0.0:
            for i in range(2 * self.depth):  # (the engine refuses depth > 16, so 32 bits hold every sublayer)
                if random.random() < self.layer_dropout:
                    skip |= 1 << i
        return skip, self._step_seed()

    def _apply_step_control(self, skip: int, seed: int) -> None:
        """Device-resident control (svsr_lrw_step_control): the mask and the seed go to two control words in the
        workspace on the current stream; the engine then launches every sublayer predicated on them."""
        check(lib().svsr_lrw_step_control(self._h, C.c_int(1), C.c_uint32(skip), C.c_uint64(seed), self._stream()),
              "svsr_lrw_step_control")
        self._last_skip = skip

    def _step_seed(self) -> int:
        """Seed of this step's dropout masks (training mode only); `self.dropout_seed` pins it for reproducible tests."""
        if not self.training or not (self.ff_dropout > 0 or self.hf or self.emb_dropout > 0 or self.attn_dropout > 0):
            return 0
        fixed = getattr(self, "dropout_seed", None)
        return fixed if fixed is not None else random.getrandbits(63)

    @torch.no_grad()
    def forward_precise(self, videos: torch.Tensor, audio_tokens: torch.Tensor, labels: torch.Tensor,
                        word_mask: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Parity-mode forward (fp32 activations, split-bf16 tensor-core operands; csrc/precise.cuh): same outputs as
        forward() at fp32-class accuracy, for every encoder configuration (x-transformers dim 512, the dim-513 word-boundary
        variant, HuggingFace BERT). Forward only and dropout-free; `last_hidden_state()` / `logits_audio()` read its
        results."""
        videos = videos.to(self.device_, torch.float32).contiguous()
        wm = None
        if self.use_wb:  # word_mask is hidden channel 512 (lightning.py:145-150)
            B, T = videos.shape[0], videos.shape[2]
            wm = word_mask.to(self.device_, torch.float32).contiguous()
            if tuple(wm.shape) != (B, T):
                raise ValueError(f"word_mask must be [B, T] = {(B, T)} with data.use_word_boundary, got {tuple(wm.shape)}")
        audio_tokens = audio_tokens.to(self.device_, torch.long).contiguous()
        labels = labels.to(self.device_)
        self._ensure(videos)
        L = lib()
        L.svsr_lrw_precise_workspace_bytes.restype = C.c_int64
        need = L.svsr_lrw_precise_workspace_bytes(self._h)
        if getattr(self, "_pws", None) is None or self._pws.numel() < need + 1024:
            self._pws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device_)
        pptr = (self._pws.data_ptr() + 1023) & ~1023
        hard = labels.long().contiguous() if labels.dtype in (torch.long, torch.int32) else None
        soft = labels.float().contiguous() if hard is None else None
        metrics = torch.zeros(8, device=self.device_)
        check(L.svsr_lrw_forward_precise(
            self._h, C.c_void_p(pptr), C.c_int64(need), C.c_void_p(videos.data_ptr()),
            C.c_void_p(audio_tokens.data_ptr()), C.c_int64(audio_tokens.stride(0)),
            C.c_void_p(hard.data_ptr() if hard is not None else 0), C.c_void_p(soft.data_ptr() if soft is not None else 0),
            C.c_void_p(wm.data_ptr() if wm is not None else 0), C.c_int(int(self.training)), C.c_uint32(0),
            C.c_void_p(metrics.data_ptr()), self._stream()),
            "svsr_lrw_forward_precise")
        self._precise_logits = True
        return {"loss_total": metrics[0], "loss_category": metrics[1], "loss_audio": metrics[2],
                "accuracy_top1": metrics[3], "accuracy_top5": metrics[4]}

    def _native_backward(self, grad_metrics: torch.Tensor) -> None:
        """d(loss_total) only: the other returned entries are metrics (the reference logs them, never differentiates
        them separately). Accumulates into the flat gradient arena."""
        g = grad_metrics.contiguous()  # element 0 = d(loss_total), read on the device (no host sync)
        need_attach = any(p.grad is None for p in self._param_views.values())
        if need_attach:  # zero_grad(set_to_none=True) dropped the views: start from a clean arena
            self._flat_g.zero_()
        check(lib().svsr_lrw_backward(self._h, C.c_void_p(g.data_ptr()), self._stream()), "svsr_lrw_backward")
        allreduce_mean_(self._flat_g, self.sync_grads)
        if need_attach:
            self._attach_grads()

    # named intermediate tensors (parity tests)
    def last_hidden_state(self) -> torch.Tensor:
        B, T, _, _ = self._shape_key
        return self._named_tensor("last_hidden_state", (B, T + 1, self._dim_pitch))[:, :, : self.dim].clone()

    def logits_audio(self) -> torch.Tensor:
        """lightning.py:168-169 `logits_audio`: NOT produced by the step (fused head) -- materialised here on request
        (after forward_precise the buffer already holds the parity-mode logits)."""
        B, T, _, _ = self._shape_key
        if not getattr(self, "_precise_logits", False):
            check(lib().svsr_lrw_logits_audio(self._h, self._stream()), "svsr_lrw_logits_audio")
        return self._named_tensor("logits_audio", (B, T, self.audio_alignment * self.vq_groups,
                                                   self.audio_vocab_size)).clone()

    def logits_category(self) -> torch.Tensor:
        B = self._shape_key[0]
        return self._named_tensor("logits_category", (B, -1))[:, : self.word_labels].clone()

    # ---- the reference's Lightning hooks (lightning.py:194-223) ----
    def log_dict(self, *a, **k):
        pass

    def training_step(self, batch, idx: int) -> torch.Tensor:
        self.is_train = True
        if _cfg_get(self.config, "train.use_cutmix", False):  # lightning.py:196-197
            batch = self.cutmix(*batch)
        metrics = self(*batch)
        self.log_dict({f"train/{k}": v for k, v in metrics.items()})
        return metrics["loss_total"]

    def validation_step(self, batch, idx: int):
        self.is_train = False
        metrics = self(*batch)
        self.log_dict({f"val/{k}": v for k, v in metrics.items()}, sync_dist=True)

    def test_step(self, batch, idx: int):
        self.is_train = False
        metrics = self(*batch)
        self.log_dict({f"test/{k}": v for k, v in metrics.items()}, sync_dist=True)

    def configure_optimizers(self):
        """AdamW with decay on ndim >= 2 parameters only (lightning.py:216-223); the cosine schedule comes from
        transformers.get_scheduler exactly as in the reference when that package is importable."""
        do_decay = [p for p in self.parameters() if p.requires_grad and p.ndim >= 2]
        no_decay = [p for p in self.parameters() if p.requires_grad and p.ndim < 2]
        groups = [{"params": do_decay}, {"params": no_decay, "weight_decay": 0.0}]
        okw = dict(_cfg_get(self.config, "optim.optimizer", {}) or {})
        optimizer = torch.optim.AdamW(groups, **okw)
        skw = dict(_cfg_get(self.config, "optim.scheduler", {}) or {})
        try:
            from transformers import get_scheduler

            scheduler = get_scheduler(optimizer=optimizer, **skw)
            return [optimizer], [{"scheduler": scheduler, "interval": "step"}]
        except Exception:
            return [optimizer], []

    def __del__(self):
        try:
            self._engines.destroy()
        except Exception:
            pass


# README pseudo-API name (README.md:26-56)
Model = TransformerLightningModule
