"""Host-side load path without a GPU: params.json / config.json / safetensors / GGUF containers are parsed and the
tensor inventory (names, dtypes, shapes) is validated BEFORE any device is touched, so a broken Model artifact fails
with a precise error on any box; a complete artifact then fails with SSB_ENODEV here (no CPU fallback)."""
import json
import os
import struct

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
