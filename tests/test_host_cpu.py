"""CPU tests of the host-side logic: arena layout defined by the native engine (host code, no GPU needed), the LR
schedule, and the data-parallel gradient reduction on a 2-rank gloo group."""
import ctypes as C
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lrw_oracle as O
from syncvsr_b200 import _lib
from syncvsr_b200.lightning import LrwConfig


def _engine(B=2, depth=12, A=4, V=320):
    L = _lib.lib()
    for f in ("svsr_lrw_param_count", "svsr_lrw_buffer_count", "svsr_lrw_workspace_bytes", "svsr_lrw_decay_count"):
        getattr(L, f).restype = C.c_int64
    h = C.c_void_p()
    cfg = LrwConfig(B, 29, 88, 88, 512, depth, 8, A, 2, V, 500, 1, 10.0, 0.0, 1e-5, 0.1, 0.0)
    _lib.check(L.svsr_lrw_create(C.byref(cfg), C.byref(h)), "create")
    return L, h


def _table(L, h):
    name, ndim, off, decay = C.c_char_p(), C.c_int(), C.c_int64(), C.c_int()
    shape = (C.c_int64 * 5)()
    rows = []
    for i in range(L.svsr_lrw_num_params(h)):
        _lib.check(L.svsr_lrw_param_info(h, i, C.byref(name), C.byref(ndim), shape, C.byref(off), C.byref(decay)), "info")
        rows.append((name.value.decode(), tuple(shape[k] for k in range(ndim.value)), off.value, bool(decay.value)))
    return rows


def test_arena_layout_matches_reference_state_dict_and_decay_rule():
    L, h = _engine()
    rows = _table(L, h)
    ref = O.make_params(0)  # reference-named tensors (validated against the reference module in test_oracle_cpu)
    names = {r[0]: r[1] for r in rows}
    trainable = {k: tuple(v.shape) for k, v in ref.items() if "running" not in k}
    assert names == trainable
    n_decay, n_total = L.svsr_lrw_decay_count(h), L.svsr_lrw_param_count(h)
    spans = sorted((off, off + math.prod(shape)) for _, shape, off, _ in rows)
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0, "overlapping parameters"
    assert spans[-1][1] <= n_total
    for name, shape, off, decay in rows:
        assert off % 4 == 0
        assert decay == (len(shape) >= 2)  # lightning.py:217-219
        assert (off < n_decay) == decay, name
    # q, k, v of a layer are adjacent (one [1536, 512] operand)
    offs = {r[0]: r[2] for r in rows}
    assert offs["encoder.layers.0.1.to_k.weight"] - offs["encoder.layers.0.1.to_q.weight"] == 512 * 512
    assert offs["encoder.layers.0.1.to_v.weight"] - offs["encoder.layers.0.1.to_k.weight"] == 512 * 512
    assert sum(math.prod(s) for _, s, _, _ in rows) == 63_152_308  # = reference trainable params minus unused resnet stem/fc
    L.svsr_lrw_destroy(h)


def test_engine_rejects_unsupported_configs():
    L = _lib.lib()
    h = C.c_void_p()
    bad = LrwConfig(2, 29, 88, 88, 520, 12, 8, 4, 2, 320, 500, 1, 10.0, 0.0, 1e-5, 0.1, 0.0)
    assert L.svsr_lrw_create(C.byref(bad), C.byref(h)) == -1
    assert b"dim" in L.svsr_last_error()
    too_long = LrwConfig(2, 80, 88, 88, 512, 12, 8, 4, 2, 320, 500, 1, 10.0, 0.0, 1e-5, 0.1, 0.0)
    assert L.svsr_lrw_create(C.byref(too_long), C.byref(h)) == -1


def test_workspace_scales_with_batch():
    L, h2 = _engine(B=2)
    _, h64 = _engine(B=64)
    w2, w64 = L.svsr_lrw_workspace_bytes(h2), L.svsr_lrw_workspace_bytes(h64)
    assert w2 < w64 < 8 * 2**30
    assert L.svsr_lrw_param_count(h2) == L.svsr_lrw_param_count(h64)


def test_cosine_schedule_matches_transformers():
    from transformers import get_scheduler

    from syncvsr_b200.train import cosine_with_warmup

    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=1e-4)
    sch = get_scheduler("cosine", optimizer=opt, num_warmup_steps=10, num_training_steps=100)
    for step in range(1, 100):
        opt.step()
        sch.step()
        assert sch.get_last_lr()[0] == pytest.approx(cosine_with_warmup(step, 1e-4, 10, 100), rel=1e-6, abs=1e-12)


def _dp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank holds the gradient arena of its own shard (synthetic): the DP step SUM-reduces the flat arena once
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(1000, generator=g)
    mine = flat.clone()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ok = torch.allclose(flat, sum(gathered)) and torch.allclose(flat / world, torch.stack(gathered).mean(0), atol=1e-6)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ---------------------------------------------------------------------------------------------------------------
# LRS sentence-level engine: arena layout (host code only)
# ---------------------------------------------------------------------------------------------------------------
def _lrs_engine(adim=256, heads=4, eunits=512, elayers=2, dlayers=1, odim=300, A=2, G=2, V=64, B=3, T=12, Lmax=17):
    from syncvsr_b200.e2e import LrsConfig

    L = _lib.lib()
    for f in ("param_count", "buffer_count", "workspace_bytes", "decay_count"):
        getattr(L, f"svsr_lrs_{f}").restype = C.c_int64
    h = C.c_void_p()
    cfg = LrsConfig(B, T, 88, 88, adim, heads, eunits, elayers, dlayers, eunits, odim, 31, A, G, V, Lmax, 0.1, 0.1, 10.0,
                    1e-5, 0.1)
    _lib.check(L.svsr_lrs_create(C.byref(cfg), C.byref(h)), "create")
    return L, h


def test_lrs_arena_layout_matches_reference_state_dict():
    from oracle import lrs_oracle as OL

    L, h = _lrs_engine()
    name, ndim, off, decay = C.c_char_p(), C.c_int(), C.c_int64(), C.c_int()
    shape = (C.c_int64 * 5)()
    rows = []
    for i in range(L.svsr_lrs_num_params(h)):
        _lib.check(L.svsr_lrs_param_info(h, i, C.byref(name), C.byref(ndim), shape, C.byref(off), C.byref(decay)), "info")
        rows.append((name.value.decode(), tuple(shape[k] for k in range(ndim.value)), off.value, bool(decay.value)))
    # oracle.make_params is loaded strictly by the reference E2E module (tests/test_lrs_oracle_cpu.py): same key set
    ref = OL.make_params(0, adim=256, heads=4, eunits=512, elayers=2, dlayers=1, odim=300, n_audio=256)
    assert {r[0]: r[1] for r in rows} == {k: tuple(v.shape) for k, v in ref.items() if "running" not in k}
    n_decay, n_total = L.svsr_lrs_decay_count(h), L.svsr_lrs_param_count(h)
    spans = sorted((o, o + math.prod(s)) for _, s, o, _ in rows)
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0, "overlapping parameters"
    assert spans[-1][1] <= n_total
    for nm, shp, o, dec in rows:
        assert o % 4 == 0 and dec == (len(shp) >= 2) and (o < n_decay) == dec, nm  # LRS/video/lightning.py:89-92
    offs = {r[0]: r[2] for r in rows}
    for pre in ("encoder.encoders.1.self_attn", "decoder.decoders.0.self_attn", "decoder.decoders.0.src_attn"):
        for kind, step in (("weight", 256 * 256), ("bias", 256)):  # q | k | v adjacent: one fused GEMM operand
            assert offs[f"{pre}.linear_k.{kind}"] - offs[f"{pre}.linear_q.{kind}"] == step
            assert offs[f"{pre}.linear_v.{kind}"] - offs[f"{pre}.linear_k.{kind}"] == step
    bufs = []
    for i in range(L.svsr_lrs_num_buffers(h)):
        _lib.check(L.svsr_lrs_buffer_info(h, i, C.byref(name), C.byref(ndim), shape, C.byref(off)), "info")
        bufs.append(name.value.decode())
    assert sorted(bufs) == sorted(k for k in ref if "running" in k)
    assert L.svsr_lrs_workspace_bytes(h) > 0
    L.svsr_lrs_destroy(h)


def test_lrs_engine_rejects_unsupported_geometry():
    from syncvsr_b200.e2e import LrsConfig

    L = _lib.lib()
    h = C.c_void_p()
    bad = LrsConfig(2, 12, 88, 88, 200, 4, 512, 1, 1, 512, 300, 31, 2, 2, 64, 17, 0.1, 0.1, 10.0, 1e-5, 0.1)  # adim 200
    assert L.svsr_lrs_create(C.byref(bad), C.byref(h)) != 0
    assert b"adim" in L.svsr_last_error()
    bad = LrsConfig(2, 12, 88, 88, 256, 4, 512, 1, 1, 512, 300, 33, 2, 2, 64, 17, 0.1, 0.1, 10.0, 1e-5, 0.1)  # kernel 33
    assert L.svsr_lrs_create(C.byref(bad), C.byref(h)) != 0


# ---------------------------------------------------------------------------------------------------------------
# staged gradient all-reduce of the LRS step (train.SentenceDataParallelStep): ranges + a 2-rank gloo run
# ---------------------------------------------------------------------------------------------------------------
def _lrs_offsets():
    L, h = _lrs_engine()
    name, ndim, off, decay = C.c_char_p(), C.c_int(), C.c_int64(), C.c_int()
    shape = (C.c_int64 * 5)()
    offs = {}
    for i in range(L.svsr_lrs_num_params(h)):
        _lib.check(L.svsr_lrs_param_info(h, i, C.byref(name), C.byref(ndim), shape, C.byref(off), C.byref(decay)), "info")
        shp = tuple(shape[k] for k in range(ndim.value))
        offs[name.value.decode()] = (off.value, math.prod(shp), shp, bool(decay.value))
    n_total = L.svsr_lrs_param_count(h)
    L.svsr_lrs_destroy(h)
    return offs, n_total


def test_lrw_stage_ranges_partition_the_gradient_arena():
    """LRW: heads + encoder | resnet.layer4-3 | layer2-1 + stem3d -- six slices (decayed + non-decayed per stage) tile the
    arena; the last stage (the only all-reduce that cannot overlap compute) is < 2 % of the bytes."""
    from syncvsr_b200.train import lrw_stage_of, stage_ranges

    L, h = _engine()
    rows = _table(L, h)
    n_total = L.svsr_lrw_param_count(h)
    L.svsr_lrw_destroy(h)
    offs = {nm: (off, math.prod(shp), shp, dec) for nm, shp, off, dec in rows}
    ranges = stage_ranges(offs, lrw_stage_of)
    flat = sorted(r for rs in ranges for r in rs)
    assert flat[0][0] == 0 and flat[-1][1] == n_total and len(flat) == 6
    for (a0, a1), (b0, b1) in zip(flat, flat[1:]):
        assert a1 == b0
    size = [sum(b - a for a, b in rs) for rs in ranges]
    assert size[2] < 0.02 * n_total and size[0] > 0.8 * n_total and size[1] > 10_000_000


def test_lrs_stage_ranges_partition_the_gradient_arena():
    """The three backward stages' ranges (heads + decoder | Conformer blocks | frontend + embed) cover every parameter
    exactly once with a handful of contiguous slices: one NCCL call per slice, nothing reduced twice, nothing missed."""
    from syncvsr_b200.train import lrs_stage_of, stage_ranges

    offs, n_total = _lrs_offsets()
    ranges = stage_ranges(offs)
    assert len(ranges) == 3 and all(ranges)
    flat = sorted(r for rs in ranges for r in rs)
    assert flat[0][0] == 0 and flat[-1][1] == n_total
    for (a0, a1), (b0, b1) in zip(flat, flat[1:]):
        assert a1 == b0, "gap or overlap between stage ranges"
    assert sum(len(rs) for rs in ranges) <= 8  # decayed + non-decayed region per stage (+ the stage-0 heads split)
    cover = torch.full((n_total,), -1, dtype=torch.int8)
    for st, rs in enumerate(ranges):
        for a, b in rs:
            cover[a:b] = st
    for key, (off, n, *_r) in offs.items():
        assert (cover[off:off + n] == lrs_stage_of(key)).all(), key


def _staged_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from syncvsr_b200.lightning import allreduce_mean_
    from syncvsr_b200.train import allreduce_ranges, stage_ranges

    offs = {f"decoder.w{i}": (i * 40, 38) for i in range(3)}
    offs.update({f"encoder.encoders.0.w{i}": (120 + i * 40, 40) for i in range(3)})
    offs.update({"encoder.frontend.w": (240, 17), "encoder.embed.0.weight": (260, 40)})
    ranges = stage_ranges(offs)
    g = torch.Generator().manual_seed(7 + rank)
    flat = torch.randn(300, generator=g)
    whole = flat.clone()
    hs = []
    for st in range(3):  # what SentenceDataParallelStep does after each svsr_lrs_backward_stage
        hs += allreduce_ranges(flat, ranges[st])
    for h in hs:
        h.wait()
    dist.all_reduce(whole, op=dist.ReduceOp.SUM)
    mean = allreduce_mean_(torch.full((5,), float(rank + 1)))  # the module-level DDP replacement: SUM / world
    ok = torch.equal(flat, whole) and torch.allclose(mean, torch.full((5,), 1.5))
    q.put((rank, bool(ok), ranges))
    dist.destroy_process_group()


def test_staged_allreduce_equals_one_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 77) % 500
    procs = [ctx.Process(target=_staged_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[:2] for r in res) == [(0, True), (1, True)]
    assert res[0][2] == [[(0, 120)], [(120, 240)], [(240, 300)]]  # (4-element padded tensors merge into one slice)
