"""GPU parity: CUDA log-mel (csrc/logmel.cu through the C ABI) vs the oracle and vs the reference goldens.
Bar from BASELINE.json north_star: log-mel features within 1e-4 of the reference."""
import os

import numpy as np
import pytest

from oracle import logmel as om
from willow_inference_server_b200 import _lib, audio

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _inputs():
    rng = np.random.default_rng(20260923)
    t30 = np.arange(480000, dtype=np.float64) / 16000.0
    return {
        "synth_3p84s": om.synth_utterance(61440, seed=1234),
        "synth_10p688s": om.synth_utterance(171008, seed=1235),
        "synth_29p248s": om.synth_utterance(467968, seed=1236),
        "zeros_30s": np.zeros(480000, np.float32),
        "sine_fullscale_30s": np.sin(2 * np.pi * 440.0 * t30).astype(np.float32),
        "noise_30s": (0.5 * rng.standard_normal(480000)).astype(np.float32),
        "long_35s_trimmed": om.synth_utterance(560000, seed=1237),
        "tiny_1sample": np.array([0.25], np.float32),
    }


def test_logmel_vs_reference_goldens_and_oracle(golden_dir):
    inputs = _inputs()
    names = sorted(inputs)
    mel = audio.log_mel_batch([inputs[n] for n in names])  # ragged batch, padding/trimming fused in the kernel
    assert mel.shape == (len(names), 80, 3000) and mel.dtype == np.float32
    for i, name in enumerate(names):
        g = np.load(os.path.join(golden_dir, f"logmel_{name}.npz"))
        assert np.abs(mel[i][:, ::16] - g["sub"]).max() <= TOL, name
        if "full" in g.files:
            assert np.abs(mel[i] - g["full"]).max() <= TOL, name
        ref = om.log_mel_spectrogram(om.pad_or_trim(inputs[name]))
        assert np.abs(mel[i] - ref).max() <= TOL, name


def test_single_call_surface_matches_reference_shape():
    pcm = om.synth_utterance(61440, seed=1234)
    out = audio.log_mel_spectrogram(audio.pad_or_trim(pcm))  # exactly how main.py:613-614 calls it
    a = out.numpy()
    assert a.shape == (80, 3000) and a.dtype == np.float32
    b = audio.log_mel_spectrogram(pcm).numpy()  # unpadded: the kernel pads
    assert np.array_equal(a, b)
    assert np.abs(a - om.log_mel_spectrogram(om.pad_or_trim(pcm))).max() <= TOL


def test_s16_input_matches_float_path():
    pcm = om.synth_utterance(100000, seed=7)
    s16 = np.clip(np.round(pcm * 32768.0), -32768, 32767).astype(np.int16)
    a = audio.log_mel_spectrogram(s16).numpy()
    ref = om.log_mel_spectrogram(om.pad_or_trim(s16.astype(np.float32) / 32768.0))
    assert np.abs(a - ref).max() <= TOL


def test_properties_full_size_batch():
    # size-independent properties on a 32-window batch of 30-s noise: range, floor rule, determinism
    rng = np.random.default_rng(5)
    pcm = [(0.3 * rng.standard_normal(480000)).astype(np.float32) for _ in range(32)]
    h = _lib.Handle.frontend(0)
    m1 = audio.log_mel_batch(pcm, h)
    m2 = audio.log_mel_batch(pcm, h)
    assert np.array_equal(m1, m2)
    mx = m1.reshape(32, -1).max(1)
    mn = m1.reshape(32, -1).min(1)
    assert np.all(mn >= mx - 2.0 - 1e-6)  # max(x, max - 8) then /4
    assert np.isfinite(m1).all()
    # scaling the signal by 10 shifts unclamped values by log10(100)/4 = 0.5
    m3 = audio.log_mel_batch([p * 10 for p in pcm[:2]], h)
    assert np.abs((m3 - m1[:2]) - 0.5).max() < 2e-4


def test_bad_arguments():
    h = _lib.Handle.frontend(0)
    with pytest.raises(ValueError):
        h.logmel(np.zeros(10, np.float64), [0], [10])
    with pytest.raises(ValueError):
        audio.log_mel_spectrogram(np.zeros((2, 100), np.float32))


def test_long_audio_windows_are_framed_on_the_gpu():
    # chunk_iter windows come straight out of one PCM buffer (offset + length), padding fused --
    # identical to the reference recipe log_mel_spectrogram(pad_or_trim(chunk)) window by window
    from willow_inference_server_b200 import audio

    rng = np.random.default_rng(5)
    x = (0.3 * np.sin(np.arange(75 * 16000) * 0.05) + 0.05 * rng.standard_normal(75 * 16000)).astype(np.float32)
    mel, strides = audio.log_mel_chunks(x)
    want = [(audio.log_mel_spectrogram(audio.pad_or_trim(c)).numpy(), s) for c, s in audio.chunk_iter(x)]
    assert mel.shape == (len(want), 80, 3000) and strides == [s for _, s in want] and len(want) == 6
    for i, (w, _) in enumerate(want):
        assert np.array_equal(mel[i], w), i
    assert audio.log_mel_chunks(np.zeros(0, np.float32))[0].shape == (0, 80, 3000)
