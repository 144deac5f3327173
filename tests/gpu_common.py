"""Shared helpers for the -m gpu parity tests: seeded synthetic models for engine (C ABI) and oracle alike."""
import functools

import numpy as np

from oracle import logmel as om
from oracle.whisper_ref import WhisperOracle
from willow_inference_server_b200 import _lib, weights as W

PROMPT = [50258, 50259, 50359, 50363]


@functools.lru_cache(maxsize=4)
def model_pair(d_model=128, n_heads=2, n_layers=2, seed=11, eot_ramp=(10, 8.0)):
    dims = W.WhisperDims(d_model=d_model, n_heads=n_heads, n_enc_layers=n_layers, n_dec_layers=n_layers)
    tensors = W.synth_engine_tensors(dims, seed=seed, eot_ramp=eot_ramp)
    buf = np.zeros(W.blob_nbytes(tensors), np.uint8)
    W.write_blob_into(buf, dims, tensors)
    oracle = WhisperOracle.from_blob(buf)
    handle = _lib.Handle.from_host(buf, 0)
    return dims, oracle, handle


@functools.lru_cache(maxsize=2)
def mel_inputs(n=4):
    durations = [61440, 160000, 480000, 171008, 30000, 467968][:n]
    return om.log_mel_batch([om.synth_utterance(m, 100 + i) for i, m in enumerate(durations)])
