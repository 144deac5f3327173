"""Checkpoint readers: HF directories are checked against transformers itself; the CTranslate2
model.bin reader is UNPINNED (no ctranslate2 / model.bin in the image) -- only writer/reader self-consistency and the
int8 de-quantisation rule are tested here, with the writer below restating the same published layout."""
import json
import os
import struct

import numpy as np
import pytest

from willow_inference_server_b200 import loaders, weights as W


def _tiny_hf(tmp_path, safe=True):
    torch = pytest.importorskip("torch")
    tr = pytest.importorskip("transformers")
    cfg = tr.WhisperConfig(d_model=128, encoder_layers=2, decoder_layers=2, encoder_attention_heads=2,
                           decoder_attention_heads=2, encoder_ffn_dim=512, decoder_ffn_dim=512, vocab_size=51865,
                           # ids of the multilingual checkpoints (openai/whisper-*, tovera/wis-whisper-*); the class
                           # defaults are the English-only ones
                           decoder_start_token_id=50258, eos_token_id=50257, bos_token_id=50257, pad_token_id=50257,
                           begin_suppress_tokens=[220, 50257], suppress_tokens=list(W.NON_SPEECH_TOKENS_MULTI))
    torch.manual_seed(3)
    model = tr.WhisperForConditionalGeneration(cfg).eval()
    d = str(tmp_path / "hf")
    model.save_pretrained(d, safe_serialization=safe)
    return model, d


def test_hf_directory_matches_transformers_state_dict(tmp_path):
    model, d = _tiny_hf(tmp_path)
    dims, tensors = loaders.load_hf_dir(d)
    assert (dims.d_model, dims.n_heads, dims.n_enc_layers, dims.n_dec_layers, dims.n_vocab) == (128, 2, 2, 2, 51865)
    assert dims.sot == 50258 and dims.eot == 50257 and {dims.sot, dims.transcribe, dims.translate} <= set(dims.suppress_ids)
    sd = {k: v.detach().float().numpy() for k, v in model.state_dict().items()}
    want = W.pack_state_dict(sd, dims)
    assert set(want) == set(tensors)
    for k in want:
        assert want[k].dtype == tensors[k].dtype and np.array_equal(want[k], tensors[k]), k
    # the blob round-trips and the front door recognises every kind of directory
    out = loaders.convert(d, str(tmp_path / "wisb"))
    dims2, t2 = loaders.load_any(os.path.dirname(out))
    assert dims2.d_model == 128 and np.array_equal(t2["dec.tok_emb"], tensors["dec.tok_emb"])
    assert dims2.suppress_ids == sorted(set(dims.suppress_ids))


def test_hf_logits_through_the_oracle(tmp_path):
    # semantics, not just names: the oracle run on the converted blob reproduces transformers' own forward pass
    torch = pytest.importorskip("torch")
    from oracle.whisper_ref import WhisperOracle

    model, d = _tiny_hf(tmp_path)
    dims, tensors = loaders.load_hf_dir(d)
    buf = np.zeros(W.blob_nbytes(tensors), np.uint8)
    W.write_blob_into(buf, dims, tensors)
    oracle = WhisperOracle.from_blob(buf)
    rng = np.random.default_rng(0)
    mel = rng.standard_normal((1, 80, 3000)).astype(np.float32) * 0.3
    toks = [50258, 50259, 50359, 50363, 100, 2000]
    with torch.no_grad():
        want = model(input_features=torch.from_numpy(mel), decoder_input_ids=torch.tensor([toks])).logits[0].numpy()
    got = oracle.forced_logits(oracle.encode(mel)[0], toks).numpy()
    # engine tensors are fp16-rounded copies of the fp32 HF parameters: tolerance = that rounding through the network
    assert np.abs(got - want).max() < 5e-2 * max(1.0, np.abs(want).max())


def test_hf_pytorch_bin_and_errors(tmp_path):
    model, d = _tiny_hf(tmp_path, safe=False)
    dims, tensors = loaders.load_hf_dir(d)
    assert tensors["enc.conv1.w"].shape == (128, 240)
    with pytest.raises(ValueError):
        loaders.dims_from_hf_config({"d_model": 384, "encoder_layers": 4, "decoder_layers": 4, "encoder_attention_heads": 6,
                                     "decoder_attention_heads": 6, "vocab_size": 51864})  # English-only vocabulary
    with pytest.raises(FileNotFoundError):
        os.makedirs(tmp_path / "empty")
        json.dump(json.load(open(os.path.join(d, "config.json"))), open(tmp_path / "empty" / "config.json", "w"))
        loaders.load_hf_dir(str(tmp_path / "empty"))


# ---------------------------------------------------------------------------------------------- CTranslate2 (unpinned)
def _ct2_variables(sd, dims, quant):
    """HF-named state dict -> CTranslate2 WhisperSpec variables (naming as restated in loaders.py)."""
    d = dims.d_model
    z = np.zeros(d, np.float32)
    v = {}

    def dense(name, w, b=None):
        w = np.asarray(w, np.float32)
        if quant:
            scale = 127.0 / np.abs(w.reshape(w.shape[0], -1)).max(axis=1)
            v[name + "/weight"] = np.round(w * scale.reshape((-1,) + (1,) * (w.ndim - 1))).astype(np.int8)
            v[name + "/weight_scale"] = scale.astype(np.float32)
        else:
            v[name + "/weight"] = w.astype(np.float16)
        if b is not None:
            v[name + "/bias"] = np.asarray(b, np.float32)

    def ln(name, p):
        v[name + "/gamma"], v[name + "/beta"] = sd[p + ".weight"], sd[p + ".bias"]

    def self_attn(name, p):
        dense(name + "/linear_0", np.concatenate([sd[p + ".q_proj.weight"], sd[p + ".k_proj.weight"], sd[p + ".v_proj.weight"]]),
              np.concatenate([sd[p + ".q_proj.bias"], z, sd[p + ".v_proj.bias"]]))
        dense(name + "/linear_1", sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])

    def ffn(name, p):
        ln(name + "/layer_norm", p + ".final_layer_norm")
        dense(name + "/linear_0", sd[p + ".fc1.weight"], sd[p + ".fc1.bias"])
        dense(name + "/linear_1", sd[p + ".fc2.weight"], sd[p + ".fc2.bias"])

    e = "model.encoder"
    dense("encoder/conv1", sd[e + ".conv1.weight"], sd[e + ".conv1.bias"])
    dense("encoder/conv2", sd[e + ".conv2.weight"], sd[e + ".conv2.bias"])
    v["encoder/position_encodings/encodings"] = sd[e + ".embed_positions.weight"]
    ln("encoder/layer_norm", e + ".layer_norm")
    for i in range(dims.n_enc_layers):
        ln(f"encoder/layer_{i}/self_attention/layer_norm", f"{e}.layers.{i}.self_attn_layer_norm")
        self_attn(f"encoder/layer_{i}/self_attention", f"{e}.layers.{i}.self_attn")
        ffn(f"encoder/layer_{i}/ffn", f"{e}.layers.{i}")
    dd = "model.decoder"
    dense("decoder/embeddings", sd[dd + ".embed_tokens.weight"])
    v["decoder/position_encodings/encodings"] = sd[dd + ".embed_positions.weight"]
    ln("decoder/layer_norm", dd + ".layer_norm")
    for i in range(dims.n_dec_layers):
        p = f"{dd}.layers.{i}"
        ln(f"decoder/layer_{i}/self_attention/layer_norm", p + ".self_attn_layer_norm")
        self_attn(f"decoder/layer_{i}/self_attention", p + ".self_attn")
        ln(f"decoder/layer_{i}/attention/layer_norm", p + ".encoder_attn_layer_norm")
        dense(f"decoder/layer_{i}/attention/linear_0", sd[p + ".encoder_attn.q_proj.weight"], sd[p + ".encoder_attn.q_proj.bias"])
        dense(f"decoder/layer_{i}/attention/linear_1",
              np.concatenate([sd[p + ".encoder_attn.k_proj.weight"], sd[p + ".encoder_attn.v_proj.weight"]]),
              np.concatenate([z, sd[p + ".encoder_attn.v_proj.bias"]]))
        dense(f"decoder/layer_{i}/attention/linear_2", sd[p + ".encoder_attn.out_proj.weight"], sd[p + ".encoder_attn.out_proj.bias"])
        ffn(f"decoder/layer_{i}/ffn", p)
    return v


def _write_ct2(path, variables, aliases):
    ids = {np.dtype(np.float32): 0, np.dtype(np.int8): 1, np.dtype(np.int16): 2, np.dtype(np.int32): 3, np.dtype(np.float16): 4}

    def s(x):
        b = x.encode()
        return struct.pack("<H", len(b) + 1) + b + b"\0"

    with open(path, "wb") as f:
        f.write(struct.pack("<I", 6) + s("WhisperSpec") + struct.pack("<I", 3) + struct.pack("<I", len(variables)))
        for name, a in variables.items():
            a = np.ascontiguousarray(a)
            f.write(s(name) + struct.pack("<B", a.ndim) + b"".join(struct.pack("<I", n) for n in a.shape))
            f.write(struct.pack("<BI", ids[a.dtype], a.nbytes) + a.tobytes())
        f.write(struct.pack("<I", len(aliases)))
        for a, t in aliases.items():
            f.write(s(a) + s(t))


@pytest.mark.parametrize("quant", [False, True])
def test_ct2_directory_self_consistency(tmp_path, quant):
    dims = W.WhisperDims(d_model=128, n_heads=2, n_enc_layers=2, n_dec_layers=2)
    sd = W.synth_state_dict(dims, seed=5)
    sd = {k: np.asarray(v, np.float32) for k, v in sd.items()}
    d = tmp_path / "ct2"
    os.makedirs(d)
    _write_ct2(str(d / "model.bin"), _ct2_variables(sd, dims, quant), {"decoder/projection/weight": "decoder/embeddings/weight"})
    json.dump({"suppress_ids": [1, 2, 7, 50258], "suppress_ids_begin": [220, 50257], "lang_ids": list(range(50259, 50358))},
              open(d / "config.json", "w"))
    dims2, tensors = loaders.load_any(str(d))
    assert (dims2.d_model, dims2.n_heads, dims2.n_enc_layers, dims2.n_dec_layers, dims2.n_vocab) == (128, 2, 2, 2, 51865)
    assert dims2.suppress_ids == [1, 2, 7, 50258] and dims2.n_langs == 99
    want = W.pack_state_dict(sd, dims2)
    for k, a in want.items():
        b = tensors[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        if not quant or a.dtype != np.float16:
            assert np.array_equal(a, b), k          # fp16 storage: the synthetic weights are fp16-representable
        else:                                       # int8: |w - q/scale| <= 0.5/scale = max|row| / 254 (+ fp16 rounding)
            a32, b32 = a.astype(np.float32), b.astype(np.float32)
            rows = a32.reshape(a32.shape[0], -1)
            bound = np.abs(rows).max(axis=1, keepdims=True) / 254 * 1.01 + 1e-3 * np.abs(rows).max()
            if k == "enc.conv1.w" or k == "enc.conv2.w":
                bound = np.abs(rows).max(axis=1, keepdims=True) / 254 * 1.01 + 1e-3 * np.abs(rows).max()
            assert (np.abs(rows - b32.reshape(rows.shape)) <= bound).all(), k


def test_ct2_reader_rejects_garbage(tmp_path):
    p = tmp_path / "model.bin"
    p.write_bytes(struct.pack("<I", 99) + b"xx")
    with pytest.raises(ValueError):
        loaders.read_ct2_model_bin(str(p))
