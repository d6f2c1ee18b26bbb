"""CPU tests: the C-ABI library loads and exports every symbol include/avi_b200.h declares (no compute calls), the product
path refuses to run without CUDA, and the clip-sharding helpers behave under a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avi_talking_b200 import _lib, shard


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = _lib.load(check_symbols=True)
    names = _lib.declared_symbols()
    assert len(names) >= 25 and "avi_gemm_bf16_tc" in names and "avi_flame_lbs_fwd" in names
    assert lib.avi_version() == 100
    assert lib.avi_last_error() == b""


def test_no_cpu_fallback():
    from avi_talking_b200 import ops
    x = torch.zeros(4, 8)
    with pytest.raises(RuntimeError, match="no CPU path|CUDA"):
        ops.layernorm(x, torch.ones(8), torch.zeros(8))
    with pytest.raises(RuntimeError):
        ops.linear(x, torch.zeros(3, 8), None)


def test_clip_shard_partitions():
    for n, w in ((512, 8), (64, 1), (7, 4), (3, 8)):
        parts = [shard.clip_shard(n, r, w) for r in range(w)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard.clip_shard(4, 4, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, lr, w = shard.env_rank_world()
        mine = shard.clip_shard(9, r, w)                      # ragged: 5 + 4 clips
        counts = shard.gather_clip_counts(len(mine))
        ms = shard.max_over_ranks([10.0 + rank, 3.0 - rank])
        dist.barrier()
        out.put((rank, mine, counts, ms))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, c0, t0), (r1, m1, c1, t1) = res
    assert m0 == [0, 2, 4, 6, 8] and m1 == [1, 3, 5, 7]
    assert c0 == c1 == [5, 4]
    assert t0 == t1 == [11.0, 3.0]


# ---------------------------------------------------------------------------------------------- training: flat layout + gradient buckets
def _tiny_vert_model():
    """FaceformerVert drop-in with a 2-layer wav2vec2 encoder on CPU (layout / bucket logic only; no kernels run)."""
    from transformers import Wav2Vec2Config
    from avi_talking_b200.faceformer import FaceformerVert, make_args
    from avi_talking_b200.wav2vec import Wav2Vec2Model
    torch.manual_seed(0)
    w2v = Wav2Vec2Model(Wav2Vec2Config(num_hidden_layers=2))
    return FaceformerVert(make_args(feature_dim=64), audio_encoder=w2v, template=torch.zeros(1, 1, 15069))


def test_flat_layout_and_flatten():
    from avi_talking_b200 import train
    m = _tiny_vert_model()
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    flat, lay = train.flatten_parameters(m)
    named = dict(m.named_parameters())
    for n in lay.names:                                       # values preserved, parameters are views of the flat buffer
        assert torch.equal(named[n], before[n])
        assert named[n].data_ptr() == flat.data_ptr() + 4 * lay.offsets[n]
    assert not any(n.startswith("audio_encoder.feature_extractor.") for n in lay.names)        # frozen (faceformer_vert.py:154)
    assert lay.names[-1] == "audio_encoder.masked_spec_embed"                                   # gradient under SpecAugment only
    # fused q|k|v views are exactly cat(q, k, v)
    p = "audio_encoder.encoder.layers.1.attention."
    fused = lay.span(flat, p + "q_proj.weight", 3 * 768, 768)
    assert torch.equal(fused, torch.cat([before[p + f"{x}_proj.weight"] for x in "qkv"], 0))
    # buckets: contiguous, ordered, cover the whole buffer, end on segment boundaries
    r = lay.bucket_ranges(8)
    assert 1 <= len(r) <= 8 and r[0][0] == 0 and r[-1][1] == lay.total
    assert all(a < b for a, b in r) and all(r[i][1] == r[i + 1][0] for i in range(len(r) - 1))


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from avi_talking_b200 import train
        m = _tiny_vert_model()
        flat, lay = train.flatten_parameters(m)
        bk = train.GradBuckets(lay, max_buckets=4)
        # rank r holds (r+1) * arange, pre-divided by the world size as TrainStep.backward does with its upstream gradient
        g = torch.arange(lay.total, dtype=torch.float32) * (rank + 1) * bk.prescale
        bk.begin()
        # backward reports progress in completion order; buckets launch as soon as their range is complete
        launched = []
        for s in lay.segments:
            upto = lay.offsets[lay.names[s]] if s < len(lay.names) else lay.total
            bk.ready(g, upto)
            launched.append(len(bk.works))
        scale = bk.finish(g)
        out.put((rank, launched, scale, float((g * scale - torch.arange(lay.total) * 1.5).abs().max())))
    finally:
        dist.destroy_process_group()


def test_grad_buckets_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, launched, scale, err in res:
        assert scale == 1.0 and err == 0.0                      # the buffer itself is the mean over the two ranks: 1.5*arange
        assert launched == sorted(launched) and launched[0] >= 1 and launched[-1] == 4   # overlapped: first bucket leaves early


def test_unsupported_configurations_raise():
    """No silent approximation: configurations whose kernels are not built refuse loudly."""
    from avi_talking_b200.faceformer import Faceformer, make_args
    with pytest.raises(NotImplementedError, match="vocaset"):
        Faceformer(make_args(dataset="BIWI"), audio_encoder=_tiny_vert_model().audio_encoder, template=torch.zeros(1, 1, 15069))
    m = _tiny_vert_model()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 16000), torch.zeros(1, 4, 53), torch.zeros(1, 4, 6), torch.zeros(1, 4, 100), teacher_forcing=False)


def test_flame_refuses_to_drop_gradients():
    from avi_talking_b200.flame import _refuse_grad
    x = torch.zeros(2, 3, requires_grad=True)
    with pytest.raises(NotImplementedError, match="forward-only"):
        _refuse_grad(None, x)
    with torch.no_grad():
        _refuse_grad(x)                      # fine: the caller does not expect a gradient
    _refuse_grad(torch.zeros(2, 3), None)


def test_ctypes_struct_mirrors_the_header():
    """AviGemmArgs in avi_talking_b200/_lib.py has the header's fields in the header's order and widths (include/avi_b200.h)."""
    import ctypes as C
    import re
    src = open(_lib.HEADER_PATH).read()
    body = re.search(r"typedef struct\s+AviGemmArgs\s*\{(.*?)\}\s*AviGemmArgs;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"(const\s+)?(void|float|int32_t|int64_t)\s*(\*?)\s*(.*)", decl)
        ctype = C.c_void_p if m.group(3) == "*" or "*" in m.group(4) else {"int32_t": C.c_int32, "int64_t": C.c_int64}[m.group(2)]
        for name in m.group(4).split(","):
            fields.append((name.replace("*", "").strip(), ctype))
    assert [(n, t) for n, t in _lib.AviGemmArgs._fields_] == fields


def test_entry_points_validate_their_arguments_before_touching_the_device():
    """Error behaviour of the C ABI (no GPU needed: every entry point checks its arguments first): status != 0 and a message in
    avi_last_error(), never a launch with a bad shape. Also the process-wide switches return their previous value."""
    import ctypes as C
    lib = _lib.load(check_symbols=True)
    buf = (C.c_float * 64)()
    p = C.cast(buf, C.c_void_p)

    def err():
        return lib.avi_last_error().decode()

    assert lib.avi_dropout_masks(p, C.c_int64(6), C.c_float(0.1), p, C.c_uint32(0), None) != 0 and "multiple of 4" in err()
    assert lib.avi_dropout_masks(p, C.c_int64(8), C.c_float(1.5), p, C.c_uint32(0), None) != 0 and "0 <= p < 1" in err()
    assert lib.avi_dropout_masks(p, C.c_int64(8), C.c_float(0.1), p, C.c_uint32(0x4C440000), None) != 0 and "reserved" in err()
    assert lib.avi_mask_mul_add(p, p, None, p, C.c_int64(6), None) != 0 and "multiple of 4" in err()
    assert lib.avi_split_bf16_terms(p, p, C.c_int64(4), C.c_int32(6), C.c_int32(3), C.c_uint32(0), None) != 0 and "K % 4" in err()
    assert lib.avi_split_bf16_terms(p, p, C.c_int64(4), C.c_int32(8), C.c_int32(3), C.c_uint32(0x3), None) != 0 and "0, 1 or 2" in err()
    assert lib.avi_layerdrop_spec_draw(None, p, C.c_int32(0), C.c_int64(0), C.c_float(0.1), None, 0, 0, 0, C.c_float(0), 0, p, None) != 0
    assert lib.avi_attn_train_fwd_drop(p, p, None, None, 1, 200, 4, 16, C.c_float(0.25), 0, 1, None) != 0 and "T <= 128" in err()
    assert lib.avi_attn_train_fwd_drop(p, p, None, None, 1, 16, 3, 16, C.c_float(0.25), 1, 30, None) != 0 and "4 heads" in err()
    assert lib.avi_spec_augment_fwd(p, p, p, C.c_int64(0), C.c_int32(8), None) != 0
    assert lib.avi_flame_prologue_ex(p, p, 1, p, p, p, None, None, 0, 150, 192, None) != 0 and "bad shape" in err()
    assert lib.avi_gemm_bf16_tc(None, None) != 0 and "null args" in err()
    args = _lib.AviGemmArgs()
    args.A = args.W = args.C = p.value
    args.batch, args.rows, args.N, args.K, args.conv_taps, args.conv_stride = 1, 8, 8, 48, 1, 1      # K not a multiple of 64
    args.a_ld, args.a_rows_alloc, args.c_ld, args.a_dtype, args.c_dtype = 48, 8, 8, _lib.DT_BF16, _lib.DT_F32
    assert lib.avi_gemm_bf16_tc(C.byref(args), None) != 0 and "multiple of 64" in err()
    assert lib.avi_gemm_bf16_tc_supported(C.byref(args)) == 0
    prev = lib.avi_set_dynamic_tiles(3)
    assert lib.avi_set_dynamic_tiles(prev) == 3
    prev = lib.avi_set_pdl(1)
    assert lib.avi_set_pdl(prev) == 1
