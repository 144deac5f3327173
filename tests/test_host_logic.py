"""CPU tests of the host-side mirror of the reference interface and of the C-ABI surface (no GPU compute)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import willow_inference_server_b200 as ct2
from willow_inference_server_b200 import _lib, audio, weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chunk_iter_matches_reference_golden(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "host_logic.json")))
    for n, table in g["chunk_iter"].items():
        x = np.zeros(int(n), np.float32)
        got = [[int(c.shape[0]), int(s[0]), int(s[1]), int(s[2])] for c, s in audio.chunk_iter(x)]
        assert got == table, n


def test_lcs_matches_reference_golden(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "host_logic.json")))
    for case in g["lcs"]:
        class Tok:
            all_special_ids = case["special"]
        merged = audio.find_longest_common_sequence([(s, (0, 0, 0)) for s in case["in"]], Tok())
        assert [int(v) for v in merged] == case["out"]


def test_lcs_longer_second_window_does_not_raise():
    class Tok:
        all_special_ids = [50257]
    out = audio.find_longest_common_sequence([([1, 2], 0), ([1, 2, 3, 4, 5], 0)], Tok())
    assert out.tolist()[:2] == [1, 2] and out.tolist()[-1] == 5


def test_pad_or_trim():
    x = np.arange(10, dtype=np.float32)
    assert audio.pad_or_trim(x, 4).tolist() == [0, 1, 2, 3]
    assert audio.pad_or_trim(x, 12).tolist() == list(range(10)) + [0, 0]
    assert audio.pad_or_trim(np.zeros(5, np.float32)).shape == (480000,)


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "wisb200.h")).read()
    declared = sorted(set(re.findall(r"\b(wisb_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    assert os.path.exists(_lib.LIB_PATH), "libwisb200.so is not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/wisb200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared  # the ctypes table binds exactly the header
    assert _lib.lib().wisb_abi_version() == 1


def test_storage_view_and_argument_errors():
    with pytest.raises(ValueError):
        ct2.StorageView.from_array(np.zeros((1, 80, 3000), np.float64))
    sv = ct2.StorageView.from_array(np.zeros((2, 80, 3000), np.float32))
    assert sv.shape == [2, 80, 3000]
    with pytest.raises(ValueError):
        ct2.models.Whisper("/nonexistent", device="cpu")
    with pytest.raises(ValueError):
        ct2.get_supported_compute_types("cpu")
    assert ct2.get_supported_compute_types("cuda") == {"float16"}


def test_blob_roundtrip(tmp_path):
    dims = W.WhisperDims(d_model=128, n_heads=2, n_enc_layers=1, n_dec_layers=1)
    t = W.synth_engine_tensors(dims, seed=3)
    path = str(tmp_path / "model.wisb")
    n = W.write_blob(path, dims, t)
    assert n == os.path.getsize(path) and n % 256 == 0
    d2, t2 = W.read_blob(path)
    assert d2.d_model == 128 and d2.n_vocab == 51865 and d2.suppress_ids == sorted(set(dims.suppress_ids))
    assert set(t2) == set(t)
    for k in t:
        assert t2[k].dtype == t[k].dtype and np.array_equal(t2[k], t[k]), k
    assert t["dec.tok_emb"].shape == (51968, 128) and not t["dec.tok_emb"][51865:].any()
    assert t["dec.crosskv.w"].shape == (256, 128) and t["enc.conv2.w"].shape == (128, 384)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.Handle.frontend(0)
    with pytest.raises(RuntimeError):
        audio.log_mel_spectrogram(np.zeros(16000, np.float32))
