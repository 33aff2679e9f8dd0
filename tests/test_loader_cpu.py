"""Host-side load path without a GPU: params.json / config.json / safetensors / GGUF containers are parsed and the
tensor inventory (names, dtypes, shapes) is validated BEFORE any device is touched, so a broken Model artifact fails
with a precise error on any box; a complete artifact then fails with SSB_ENODEV here (no CPU fallback)."""
import json
import os
import struct
import time

import pytest
import torch

from oracle import llama_ref, synth

EINVAL, EIO, ENODEV = -1, -2, -3


def _create(path, params=None):
    from substratus_b200 import Engine, SsbError

    with pytest.raises(SsbError) as ei:
        Engine(str(path), params or {})
    return ei.value


@pytest.fixture()
def no_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present: creation would succeed")


def test_complete_artifact_reaches_the_device_check(tmp_path, no_gpu, lib):
    cfg = synth.TINY_MHA
    llama_ref.write_hf_dir(str(tmp_path), cfg, synth.llama_state_dict(cfg, 1), shards=3)
    e = _create(tmp_path)
    assert e.code == ENODEV and "no CPU fallback" in str(e)


def test_config_errors(tmp_path, lib):
    e = _create(tmp_path, {"weights": "synthetic"})
    assert e.code == EIO and "config.json" in str(e)
    (tmp_path / "config.json").write_text("{not json")
    assert _create(tmp_path, {"weights": "synthetic"}).code == EINVAL
    (tmp_path / "config.json").write_text(json.dumps(dict(synth.TINY_MHA, model_type="mamba")))
    e = _create(tmp_path, {"weights": "synthetic"})
    assert e.code == EINVAL and "mamba" in str(e)
    (tmp_path / "config.json").write_text(json.dumps(dict(synth.TINY_MHA, num_key_value_heads=3)))
    assert _create(tmp_path, {"weights": "synthetic"}).code == EINVAL
    (tmp_path / "config.json").write_text(json.dumps(synth.TINY_MHA))
    assert _create(tmp_path, {"weights": "synthetic", "tp_size": 3}).code == EINVAL      # 2 kv heads / 3 ranks
    assert _create(tmp_path, {"weights": "bogus"}).code == EINVAL
    from substratus_b200 import load_library
    import ctypes as C

    h = C.c_void_p()
    assert load_library().ssb_engine_create(str(tmp_path).encode(), b"[1,2]", C.byref(h)) == EINVAL  # params must be an object


def test_safetensors_inventory_errors(tmp_path, lib):
    cfg = synth.TINY_MHA
    sd = synth.llama_state_dict(cfg, 1)
    missing = {k: v for k, v in sd.items() if k != "model.layers.1.mlp.up_proj.weight"}
    llama_ref.write_hf_dir(str(tmp_path / "a"), cfg, missing)
    e = _create(tmp_path / "a")
    assert e.code == EIO and "model.layers.1.mlp.up_proj.weight" in str(e)
    wrong = dict(sd)
    wrong["model.layers.0.self_attn.o_proj.weight"] = torch.zeros(256, 128, dtype=torch.bfloat16)
    llama_ref.write_hf_dir(str(tmp_path / "b"), cfg, wrong)
    e = _create(tmp_path / "b")
    assert e.code == EINVAL and "o_proj" in str(e)
    short = dict(sd)
    short["model.embed_tokens.weight"] = sd["model.embed_tokens.weight"][:100].clone()
    llama_ref.write_hf_dir(str(tmp_path / "c"), cfg, short)
    assert _create(tmp_path / "c").code == EINVAL
    llama_ref.write_hf_dir(str(tmp_path / "d"), cfg, sd)
    raw = open(tmp_path / "d" / "model.safetensors", "rb").read()
    open(tmp_path / "d" / "model.safetensors", "wb").write(raw[:4])           # truncated container
    assert _create(tmp_path / "d").code == EIO
    open(tmp_path / "d" / "model.safetensors", "wb").write(struct.pack("<Q", 1 << 40) + raw[8:64])  # absurd header length
    assert _create(tmp_path / "d").code == EIO
    # header claims a [512, 256] bf16 tensor but the offsets cover half of it: a corrupt shard must be an EIO here, not a
    # device-side read past the staging buffer
    hlen = struct.unpack("<Q", raw[:8])[0]
    hdr = json.loads(raw[8:8 + hlen])
    name = "model.embed_tokens.weight"
    b, e_ = hdr[name]["data_offsets"]
    hdr[name]["data_offsets"] = [b, b + (e_ - b) // 2]
    nh = json.dumps(hdr).encode()
    open(tmp_path / "d" / "model.safetensors", "wb").write(struct.pack("<Q", len(nh)) + nh + raw[8 + hlen:])
    e = _create(tmp_path / "d")
    assert e.code == EIO and "shape and dtype need" in str(e)
    hdr[name]["data_offsets"] = [b, e_]
    hdr[name]["shape"] = [-4, 256]
    nh = json.dumps(hdr).encode()
    open(tmp_path / "d" / "model.safetensors", "wb").write(struct.pack("<Q", len(nh)) + nh + raw[8 + hlen:])
    assert _create(tmp_path / "d").code == EIO
    os.remove(tmp_path / "d" / "model.safetensors")
    e = _create(tmp_path / "d")
    assert e.code == EIO and "safetensors" in str(e)


def test_gguf_container(tmp_path, no_gpu, lib):
    gguf = pytest.importorskip("gguf")
    from gguf import GGMLQuantizationType as T
    from test_gguf_gpu import _write_gguf

    cfg = synth.TINY_GQA
    _write_gguf(str(tmp_path / "model.bin"), cfg, synth.llama_state_dict(cfg, 9), T.Q4_0)
    e = _create(tmp_path)                      # header, metadata and all tensors resolve; only the device is missing
    assert e.code == ENODEV
    w = gguf.GGUFWriter(str(tmp_path / "x" / "model.gguf") if (tmp_path / "x").mkdir() is None else "", "mamba")
    w.add_block_count(1)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    e = _create(tmp_path / "x")
    assert e.code in (EINVAL, EIO)


# ---------------------------------------------------------------------------------------------------------------
# pytorch_model*.bin (torch.save zip + pickle) — the other HF snapshot layout, read natively by
# substratus_b200/csrc/torch_zip.cpp; ssb_model_read_tensor exposes what the engine's readers see, byte for byte.
def _bytes(t):
    return t.contiguous().view(torch.uint8).numpy().tobytes()


def _bin_dir(path, cfg, sd, save=None, **kw):
    llama_ref.write_hf_dir(str(path), cfg, {})
    for f in os.listdir(path):
        if f.endswith(".safetensors") or f.endswith(".index.json"):
            os.remove(os.path.join(path, f))
    (save or torch.save)(sd, os.path.join(path, "pytorch_model.bin"), **kw)


@pytest.mark.parametrize("variant", ["ordered_bf16", "proto4", "f16", "f32", "module_state_dict", "wrapped"])
def test_torch_bin_snapshot_reads_the_bytes_torch_saved(tmp_path, no_gpu, lib, variant):
    from collections import OrderedDict

    from substratus_b200.engine import model_read_tensor, model_tensor_count

    cfg = synth.TINY_GQA
    sd = OrderedDict(synth.llama_state_dict(cfg, 5))
    kw = {}
    if variant == "proto4":
        kw["pickle_protocol"] = 4
    if variant in ("f16", "f32"):
        sd = OrderedDict((k, v.to(torch.float16 if variant == "f16" else torch.float32)) for k, v in sd.items())
    saved = sd
    if variant == "module_state_dict":  # nn.Module.state_dict(): OrderedDict subclass state (_metadata) arrives via BUILD
        m = torch.nn.ModuleDict({"a": torch.nn.Linear(4, 4)})
        saved = m.state_dict()
        for k, v in sd.items():
            saved[k] = v
    if variant == "wrapped":
        saved = {"epoch": 3, "lr": 1e-4, "state_dict": sd, "note": "x"}
    _bin_dir(tmp_path, cfg, saved, **kw)
    assert model_tensor_count(tmp_path) == len(sd) + (2 if variant == "module_state_dict" else 0)
    for k, v in sd.items():
        dt, shape, raw = model_read_tensor(tmp_path, k)
        assert dt == {"f16": "f16", "f32": "f32"}.get(variant, "bf16") and shape == tuple(v.shape), k
        assert raw.tobytes() == _bytes(v), k
    assert _create(tmp_path).code == ENODEV  # the inventory is complete: only the device is missing


def test_torch_bin_shards_views_and_shared_storage(tmp_path, no_gpu, lib):
    from substratus_b200.engine import model_read_tensor

    cfg = synth.TINY_MHA
    sd = synth.llama_state_dict(cfg, 2)
    # shard 1 holds slices of ONE big storage (non-zero storage offsets) and a tensor stored twice under two names
    names = sorted(sd)
    half = names[: len(names) // 2]
    flat = torch.cat([sd[k].flatten() for k in half])
    views, off = {}, 0
    for k in half:
        n = sd[k].numel()
        views[k] = flat[off:off + n].view(sd[k].shape)
        off += n
    views["alias.of.norm"] = views["model.norm.weight"] if "model.norm.weight" in views else views[half[0]]
    llama_ref.write_hf_dir(str(tmp_path), cfg, sd)
    for f in os.listdir(tmp_path):
        if f.endswith(".safetensors"):
            os.remove(tmp_path / f)
    torch.save(views, tmp_path / "pytorch_model-00001-of-00002.bin")
    torch.save({k: sd[k] for k in names[len(names) // 2:]}, tmp_path / "pytorch_model-00002-of-00002.bin")
    (tmp_path / "pytorch_model.bin.index.json").write_text(json.dumps({"weight_map": {}}))
    torch.save({"unrelated": torch.zeros(3)}, tmp_path / "training_args.bin")  # not a weight shard: must be ignored
    for k, v in sd.items():
        dt, shape, raw = model_read_tensor(tmp_path, k)
        assert shape == tuple(v.shape) and raw.tobytes() == _bytes(v), k
    from substratus_b200 import SsbError

    with pytest.raises(SsbError):
        model_read_tensor(tmp_path, "unrelated")
    assert _create(tmp_path).code == ENODEV


def _stored_zip(path, members, compress=False):
    import zipfile

    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED if compress else zipfile.ZIP_STORED) as z:
        for n, b in members.items():
            z.writestr(n, b)


def test_torch_bin_refuses_what_it_does_not_understand(tmp_path, lib):
    import io
    import pickle
    import zipfile

    from substratus_b200 import SsbError
    from substratus_b200.engine import model_read_tensor

    cfg = synth.TINY_MHA
    sd = synth.llama_state_dict(cfg, 2)

    def err(d):
        with pytest.raises(SsbError) as ei:
            model_read_tensor(d, "model.norm.weight")
        return str(ei.value)

    a = tmp_path / "legacy"
    a.mkdir()
    torch.save(sd, a / "pytorch_model.bin", _use_new_zipfile_serialization=False)
    assert "legacy torch.save format" in err(a)
    b = tmp_path / "transposed"
    b.mkdir()
    torch.save({"model.norm.weight": torch.zeros(4, 6).t()}, b / "pytorch_model.bin")
    assert "not stored contiguously" in err(b)
    c = tmp_path / "hostile"  # a pickle that would run os.system under pickle.load: here nothing is ever called
    c.mkdir()

    class Boom:
        def __reduce__(self):
            return (os.system, ("touch " + str(tmp_path / "pwned"),))

    _stored_zip(c / "pytorch_model.bin", {"archive/data.pkl": pickle.dumps(Boom(), protocol=2), "archive/version": b"3\n"})
    assert "does not hold a state dict" in err(c) and not (tmp_path / "pwned").exists()
    _stored_zip(c / "pytorch_model.bin", {"archive/data.pkl": pickle.dumps({"model.norm.weight": Boom()}, protocol=2)})
    assert "no tensors" in err(c) and not (tmp_path / "pwned").exists()
    _stored_zip(c / "pytorch_model.bin", {"archive/data.pkl": b"\x80\x02\x8e\x00."})  # BYTEARRAY8: not in torch checkpoints
    assert "unsupported pickle opcode 0x8e" in err(c)
    _stored_zip(c / "pytorch_model.bin", {"archive/data.pkl": b"\x80\x02}X\x04\x00\x00"})
    assert "truncated" in err(c)
    _stored_zip(c / "pytorch_model.bin", {"archive/other": b"x"})
    assert "no data.pkl" in err(c)
    d = tmp_path / "deflated"
    d.mkdir()
    buf = io.BytesIO()
    torch.save(sd, buf)
    with zipfile.ZipFile(io.BytesIO(buf.getvalue())) as z:
        _stored_zip(d / "pytorch_model.bin", {n: z.read(n) for n in z.namelist()}, compress=True)
    assert "compressed" in err(d)
    e = tmp_path / "cut"
    e.mkdir()
    raw = buf.getvalue()
    (e / "pytorch_model.bin").write_bytes(raw[: len(raw) // 2])
    assert "zip" in err(e)
    # storage shorter than the tensor claims: re-pack with one storage truncated
    f = tmp_path / "short"
    f.mkdir()
    with zipfile.ZipFile(io.BytesIO(raw)) as z:
        members = {n: z.read(n) for n in z.namelist()}
    for n in members:
        if "/data/" in n:
            members[n] = members[n][:-2]
    _stored_zip(f / "pytorch_model.bin", members)
    with pytest.raises(SsbError) as ei:
        model_read_tensor(f, "model.embed_tokens.weight")
    assert "runs past its storage" in str(ei.value)


def test_torch_bin_zip64_directory(tmp_path, lib):
    """Checkpoints above 4 GiB use ZIP64 records (sizes/offsets 0xFFFFFFFF + extra field 0x0001, ZIP64 end record and
    locator).  Re-pack a small checkpoint with every central-directory field forced into its ZIP64 form."""
    import io
    import zipfile

    from substratus_b200.engine import model_read_tensor

    sd = {"w": torch.arange(24, dtype=torch.float32).view(4, 6), "v": torch.ones(7, dtype=torch.bfloat16)}
    buf = io.BytesIO()
    torch.save(sd, buf)
    with zipfile.ZipFile(io.BytesIO(buf.getvalue())) as z:
        members = [(n, z.read(n)) for n in z.namelist()]
    out, cd = bytearray(), bytearray()
    for name, data in members:
        nb = name.encode()
        lho = len(out)
        out += struct.pack("<IHHHHHIIIHH", 0x04034B50, 45, 0, 0, 0, 0, 0, 0xFFFFFFFF, 0xFFFFFFFF, len(nb), 20) + nb
        out += struct.pack("<HHQQ", 1, 16, len(data), len(data)) + data
        extra = struct.pack("<HHQQQ", 1, 24, len(data), len(data), lho)
        cd += struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 45, 45, 0, 0, 0, 0, 0, 0xFFFFFFFF, 0xFFFFFFFF, len(nb), len(extra), 0, 0, 0, 0,
                          0xFFFFFFFF) + nb + extra
    cd_off = len(out)
    out += cd
    z64 = len(out)
    out += struct.pack("<IQHHIIQQQQ", 0x06064B50, 44, 45, 45, 0, 0, len(members), len(members), len(cd), cd_off)
    out += struct.pack("<IIQI", 0x07064B50, 0, z64, 1)
    out += struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, 0xFFFF, 0xFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0)
    (tmp_path / "pytorch_model.bin").write_bytes(bytes(out))
    with zipfile.ZipFile(tmp_path / "pytorch_model.bin") as z:  # python agrees this is a valid archive
        assert sorted(z.namelist()) == sorted(n for n, _ in members)
    back = torch.load(tmp_path / "pytorch_model.bin")
    assert torch.equal(back["w"], sd["w"])
    for k, v in sd.items():
        dt, shape, raw = model_read_tensor(tmp_path, k)
        assert shape == tuple(v.shape) and raw.tobytes() == _bytes(v), k


# ---------------------------------------------------------------------------------------------------------------
# Corrupt artifacts must end in a clean error, never in a crash, a hang or an out-of-bounds read: seeded mutation
# fuzz of the three container readers through ssb_model_read_tensor (a GGUF header with a dimension count of 2^32-1
# used to spin for minutes — found by this fuzz).
def _mutations(raw, rng, n):
    for _ in range(n):
        m = bytearray(raw)
        for _ in range(rng.randint(1, 6)):
            if len(m) < 2:
                break
            k, pos = rng.random(), rng.randrange(len(m))
            if k < 0.5:
                m[pos] = rng.randrange(256)
            elif k < 0.7:
                del m[pos:pos + rng.randint(1, 16)]
            elif k < 0.9:
                m[pos:pos] = bytes(rng.randrange(256) for _ in range(rng.randint(1, 16)))
            else:
                del m[rng.randrange(1, len(m)):]
        yield bytes(m)


@pytest.mark.parametrize("kind", ["safetensors", "torch_bin", "gguf"])
def test_corrupt_containers_fail_cleanly(tmp_path, lib, kind):
    import ctypes as C
    import random

    import numpy as np
    from safetensors.torch import save_file

    from substratus_b200 import load_library

    sd = {"a.weight": torch.arange(64, dtype=torch.bfloat16).view(8, 8), "b": torch.ones(5, dtype=torch.float32)}
    if kind == "safetensors":
        path = tmp_path / "model.safetensors"
        save_file(sd, str(path))
        names = [b"a.weight", b"b", None]
    elif kind == "torch_bin":
        path = tmp_path / "pytorch_model.bin"
        torch.save(sd, path)
        names = [b"a.weight", b"b", None]
    else:
        gguf = pytest.importorskip("gguf")
        path = tmp_path / "model.gguf"
        w = gguf.GGUFWriter(str(path), "llama")
        w.add_block_count(1)
        w.add_embedding_length(32)
        w.add_token_list(["a", "b", "c"])
        w.add_tensor("token_embd.weight", np.arange(96, dtype=np.float32).reshape(3, 32))
        w.add_tensor("blk.0.attn_q.weight", np.ones((32, 32), dtype=np.float16))
        w.write_header_to_file()
        w.write_kv_data_to_file()
        w.write_tensors_to_file()
        w.close()
        names = [b"token_embd.weight", b"blk.0.attn_q.weight", None]
    raw = path.read_bytes()
    L = load_library()
    n, dt, nd = C.c_int64(), C.c_int(), C.c_int()
    shape = (C.c_int64 * 4)()
    out = (C.c_uint8 * 8192)()
    d = str(tmp_path).encode()
    assert all(L.ssb_model_read_tensor(d, nm, out, 8192, C.byref(n), C.byref(dt), shape, C.byref(nd)) == 0 for nm in names)
    cases = list(_mutations(raw, random.Random(11), 500))
    if kind == "gguf":  # the regression: tensor-info dimension count 0xFFFFFFFF
        i = raw.index(b"token_embd.weight") + len(b"token_embd.weight")
        cases.append(raw[:i] + b"\xff\xff\xff\xff" + raw[i + 4:])
    t0 = time.time()
    for m in cases:
        path.write_bytes(m)
        for nm in names:
            rc = L.ssb_model_read_tensor(d, nm, out, 8192, C.byref(n), C.byref(dt), shape, C.byref(nd))
            assert rc in (0, -1, -2, -4), rc  # OK or EINVAL / EIO / ENOMEM, and we are still alive
    assert time.time() - t0 < 60


def test_pathological_json_is_an_error_not_a_crash(tmp_path, lib):
    """params.json is user input (.spec.params): 200 000 nested brackets used to overflow the parser's stack."""
    import ctypes as C

    from substratus_b200 import load_library

    L = load_library()
    for payload in (b"[" * 200000, b'{"a":' * 100000, b'"' + b"\\" * 100001, b"1e999999", b'{"a":1,}', b"\xff\xfe"):
        h = C.c_void_p()
        assert L.ssb_engine_create(str(tmp_path).encode(), payload, C.byref(h)) == EINVAL
        assert b"json" in L.ssb_last_error() or b"params" in L.ssb_last_error()


@pytest.mark.parametrize("family", ["llama", "falcon"])
def test_tp_presharded_artifact(tmp_path, no_gpu, lib, family):
    """tools/tp_shard.py writes one safetensors file per tensor-parallel rank; an engine created with that
    tp_size reads ONLY that file (the original shards may be gone), its inventory passes the host-side validation and
    creation gets as far as the device check.  The slices are a partition of the original tensors."""
    import sys

    from safetensors import safe_open

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import tp_shard

    if family == "llama":
        cfg = dict(synth.TINY_GQA, num_attention_heads=8, num_key_value_heads=4, hidden_size=1024, intermediate_size=2752)
        sd = synth.llama_state_dict(cfg, 3)
    else:
        from oracle import falcon_ref as fr

        cfg = dict(fr.TINY_FALCON, hidden_size=512, num_attention_heads=8, num_kv_heads=4)
        sd = fr.falcon_state_dict(cfg, 3)
    llama_ref.write_hf_dir(str(tmp_path), cfg, sd, shards=2)
    files = tp_shard.shard(str(tmp_path), 2, quiet=True)
    assert [os.path.basename(f) for f in files] == ["rank0.safetensors", "rank1.safetensors"]
    # a partition: concatenating the rank slices along the sharded axis gives the original tensor back; replicated
    # tensors are whole in every rank file
    parts = [safe_open(f, framework="pt") for f in files]
    assert parts[1].metadata() == {"format": "ssb-tp", "tp_size": "2", "tp_rank": "1", "source": os.path.basename(str(tmp_path))}
    sharded_bytes = 0
    for name, full in sd.items():
        a, b = parts[0].get_tensor(name), parts[1].get_tensor(name)
        if a.shape == full.shape:
            assert torch.equal(a, full) and torch.equal(b, full)
        else:
            axis = 0 if a.shape[0] != full.shape[0] else 1
            assert torch.equal(torch.cat([a, b], dim=axis), full), name
            sharded_bytes += full.numel() * 2
    assert sharded_bytes > 0.5 * sum(t.numel() * 2 for t in sd.values())  # most of the checkpoint IS sharded
    # the engine takes the rank file (validated on the host, then stops at the device check) ...
    for r in (0, 1):
        e = _create(tmp_path, {"tp_size": 2, "tp_rank": r})
        assert e.code == ENODEV, str(e)
    # ... and only the rank file: without the original shards the pre-sharded ranks still load, a full load does not
    for f in os.listdir(tmp_path):
        if f.endswith(".safetensors") or f.endswith(".index.json"):
            os.remove(tmp_path / f)
    assert _create(tmp_path, {"tp_size": 2, "tp_rank": 1}).code == ENODEV
    e = _create(tmp_path, {"tp_size": 2, "tp_rank": 1, "tp_presharded": 0})
    assert e.code == EIO and "safetensors" in str(e)
    assert _create(tmp_path).code == EIO  # tp_size 1 never looks at ssb_tp*
    # a rank file in the wrong place is refused by its metadata, a foreign safetensors file too
    os.replace(tmp_path / "ssb_tp2" / "rank1.safetensors", tmp_path / "ssb_tp2" / "rank0.safetensors")
    e = _create(tmp_path, {"tp_size": 2, "tp_rank": 0})
    assert e.code == EINVAL and "pre-sharded" in str(e)


def test_tp_shard_plan_covers_the_baseline_models():
    """tools/tp_shard.py's partition at the BASELINE shapes (no tensors needed): for every TP size the engine accepts, the
    per-rank row / column ranges of each projection are contiguous, equal-sized and tile the full dimension."""
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import tp_shard
    from oracle import falcon_ref as fr

    for cfg, names in ((synth.LLAMA2_7B, "llama"), (synth.LLAMA2_13B, "llama"), (synth.LLAMA2_70B, "llama"), (fr.FALCON_40B, "falcon")):
        heads = cfg["num_attention_heads"]
        d = cfg["hidden_size"] // heads
        if names == "llama":
            kvh, inter = cfg.get("num_key_value_heads", heads), cfg["intermediate_size"]
            dims = {"model.layers.0.self_attn.q_proj.weight": ("rows", heads * d), "model.layers.0.self_attn.k_proj.weight": ("rows", kvh * d),
                    "model.layers.0.self_attn.v_proj.weight": ("rows", kvh * d), "model.layers.0.self_attn.o_proj.weight": ("cols", heads * d),
                    "model.layers.0.mlp.gate_proj.weight": ("rows", inter), "model.layers.0.mlp.up_proj.weight": ("rows", inter),
                    "model.layers.0.mlp.down_proj.weight": ("cols", inter), "model.norm.weight": ("full", 0), "lm_head.weight": ("full", 0)}
        else:
            kvh, inter = cfg["num_kv_heads"], cfg.get("ffn_hidden_size", 4 * cfg["hidden_size"])
            dims = {"transformer.h.0.self_attention.query_key_value.weight": ("rows", (heads + 2 * kvh) * d),
                    "transformer.h.0.self_attention.dense.weight": ("cols", heads * d), "transformer.h.0.mlp.dense_h_to_4h.weight": ("rows", inter),
                    "transformer.h.0.mlp.dense_4h_to_h.weight": ("cols", inter), "transformer.ln_f.weight": ("full", 0)}
        for tp in (2, 4, 8):
            if kvh % tp:
                continue
            rules = [tp_shard.plan(cfg, tp)(r) for r in range(tp)]
            for name, (kind, full) in dims.items():
                got = [rule(name) for rule in rules]
                assert all(g[0] == kind for g in got), (name, got)
                if kind != "full":
                    assert got[0][1] == 0 and got[-1][2] == full and all(got[i][2] == got[i + 1][1] for i in range(tp - 1)), (name, tp, got)
                    assert len({g[2] - g[1] for g in got}) == 1
