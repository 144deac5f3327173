"""Drop-in test through the reference's OWN orchestration code.

`do_whisper`, `do_translate` and `chunkit` are taken UNMODIFIED from the reference's main.py (main.py:91-94, 514-547,
554-770) -- extracted with `ast` at test time, nothing is copied into this repository -- and executed with
  ctranslate2            := willow_inference_server_b200              (the two-line integration diff of INTEGRATION.md)
  log_mel_spectrogram .. := willow_inference_server_b200.audio
and stubs for what stays outside the hot path (librosa, the HF tokenizer / processor, settings, the logger).

main.py itself cannot be imported (aiortc, av, librosa, ctranslate2 ... are absent), which is why the functions are cut out.
The reference source is found at /root/reference/main.py (build container) or baseline/_ref/wis_reference/main.py (staged
there, git-ignored, by scripts/stage_reference.py so that it travels to the GPU box); without either the tests skip.

  * CPU test: a recording engine with the signature of models.Whisper.generate / detect_language behind the real
    StorageView and the log-mel oracle: every call do_whisper makes binds to the shim's API.
  * GPU test: the real engine; the token ids do_whisper returns equal a direct wisb_generate on the same features.
"""
import ast
import datetime
import inspect
import logging
import math
import os
import re
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ["/root/reference/main.py", os.path.join(ROOT, "baseline", "_ref", "wis_reference", "main.py")]
MAIN_PY = next((p for p in CANDIDATES if os.path.isfile(p)), None)
needs_reference = pytest.mark.skipif(MAIN_PY is None, reason="reference main.py is not available on this box")

PROMPT = [50258, 50259, 50359, 50363]


def _extract(names):
    src = open(MAIN_PY).read()
    tree = ast.parse(src)
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {n.name for n in picked} == set(names)
    return ast.Module(body=picked, type_ignores=[])


class FakeTokenizer:
    """Stands in for the HF tokenizer WIS loads (main.py:331-333): the four prompt tokens and the special-id list."""

    def __init__(self, dims):
        from willow_inference_server_b200.languages import LANGUAGE_CODES

        self.table = {"<|startoftranscript|>": dims["sot"], "<|transcribe|>": dims["transcribe"], "<|translate|>": dims["translate"],
                      "<|notimestamps|>": dims["no_timestamps"]}
        for i, c in enumerate(LANGUAGE_CODES):
            self.table[f"<|{c}|>"] = dims["lang_first"] + i
        self.all_special_ids = list(range(dims["eot"], dims["n_vocab"]))

    def convert_tokens_to_ids(self, toks):
        return [self.table[t] for t in toks]


class FakeProcessor:
    def __init__(self, dims):
        self.tokenizer = FakeTokenizer(dims)

    def decode(self, tokens):
        return " ".join(str(int(t)) for t in tokens)


def _namespace(engine_pkg, audio_mod, model, dims, pcm, concurrent_gpu_chunks=2):
    """Globals do_whisper / do_translate read in main.py (main.py:186-232 copies them from settings)."""
    librosa = types.SimpleNamespace(load=lambda f, sr=16000, mono=True: (pcm, sr),
                                    get_duration=lambda y, sr: len(y) / float(sr))
    ns = {
        "datetime": datetime, "math": math, "re": re, "np": np, "librosa": librosa, "ctranslate2": engine_pkg,
        "logger": logging.getLogger("dropin"), "settings": types.SimpleNamespace(language="en"),
        "models": types.SimpleNamespace(whisper_model_large=model, whisper_model_medium=model, whisper_model_small=model,
                                        whisper_model_base=model, whisper_model_tiny=model, whisper_processor=FakeProcessor(dims)),
        "chunk_iter": audio_mod.chunk_iter, "pad_or_trim": audio_mod.pad_or_trim,
        "log_mel_spectrogram": audio_mod.log_mel_spectrogram, "find_longest_common_sequence": audio_mod.find_longest_common_sequence,
        "beam_size": 1, "long_beam_size": 3, "long_beam_size_threshold": 12000, "support_chunking": True,
        "concurrent_gpu_chunks": concurrent_gpu_chunks,
    }
    code = compile(_extract({"chunkit", "do_translate", "do_whisper"}), MAIN_PY, "exec")
    exec(code, ns)
    return ns


def _synth(n, seed):
    from oracle import logmel as om

    return om.synth_utterance(n, seed)


# ------------------------------------------------------------------------------------------------------------------ CPU
class RecordingWhisper:
    """Binds every call to the signature of the real shim methods and returns canned results."""

    def __init__(self):
        self.calls = []

    def generate(self, *a, **kw):
        from willow_inference_server_b200 import models

        bound = inspect.signature(models.Whisper.generate).bind(self, *a, **kw)
        feats, prompts = bound.arguments["features"], bound.arguments["prompts"]
        assert isinstance(feats, models.StorageView) and feats.shape[1:] == [80, 3000]
        assert len(prompts) == feats.shape[0] and all(len(p) == 4 for p in prompts)
        self.calls.append(("generate", feats.shape[0], list(prompts[0]), bound.arguments.get("beam_size", 5)))
        return [models.WhisperGenerationResult([[100 + i, 200 + i, 50257]]) for i in range(feats.shape[0])]

    def detect_language(self, *a, **kw):
        from willow_inference_server_b200 import models

        bound = inspect.signature(models.Whisper.detect_language).bind(self, *a, **kw)
        assert bound.arguments["features"].shape == [1, 80, 3000]
        self.calls.append(("detect_language",))
        return [[("<|de|>", 0.9), ("<|en|>", 0.1)]]


class OracleAudio:
    """wis.audio surface on the CPU: the log-mel oracle with the `.numpy()` the reference calls on the result."""

    def __init__(self):
        from oracle import logmel as om
        from willow_inference_server_b200 import audio

        self.chunk_iter, self.find_longest_common_sequence = audio.chunk_iter, audio.find_longest_common_sequence
        self.pad_or_trim = om.pad_or_trim
        self.log_mel_spectrogram = lambda x: types.SimpleNamespace(numpy=lambda: om.log_mel_spectrogram(x))


@needs_reference
def test_reference_do_whisper_drives_the_shim_api_on_cpu():
    import willow_inference_server_b200 as pkg
    from willow_inference_server_b200 import weights as W

    d = W.WhisperDims()
    dims = {"sot": d.sot, "eot": d.eot, "transcribe": d.transcribe, "translate": d.translate, "no_timestamps": d.no_timestamps,
            "lang_first": d.lang_first, "n_vocab": d.n_vocab}
    eng = RecordingWhisper()
    ns = _namespace(pkg, OracleAudio(), eng, dims, _synth(61440, 1))
    lang, text, ms, translation, speedup, dur = ns["do_whisper"]("x.flac", "large", 5, "transcribe", False, "en")
    assert (lang, text, translation, dur) == ("en", "100 200 50257", None, 3840)
    assert eng.calls == [("generate", 1, PROMPT, 5)]
    # language detection + the long-audio path: 75 s -> 6 windows, two per engine call, beam 3 (long mode), LCS merge
    eng.calls.clear()
    ns = _namespace(pkg, OracleAudio(), eng, dims, _synth(75 * 16000, 2))
    lang, text, *_ = ns["do_whisper"]("x.flac", "medium", 5, "transcribe", True, None)
    assert lang == "de"
    assert eng.calls[0] == ("detect_language",)
    assert [c[1] for c in eng.calls[1:]] == [2, 2, 2] and all(c[3] == 3 for c in eng.calls[1:])
    assert eng.calls[1][2] == [d.sot, d.lang_first + 2, d.transcribe, d.no_timestamps]  # <|de|>
    # do_translate: positional generate(features, prompts, beam_size=...) with the <|translate|> prompt (main.py:535-537)
    eng.calls.clear()
    out = ns["do_translate"](eng, pkg.StorageView.from_array(np.zeros((1, 80, 3000), np.float32)), 1, "<|de|>", 4)
    assert eng.calls == [("generate", 1, [d.sot, d.lang_first + 2, d.translate, d.no_timestamps], 4)] and out == "100 200 50257"
    # the latent reference bug is still the reference's: translate=True trips over len(int) (main.py:729)
    with pytest.raises(TypeError):
        ns["do_whisper"]("x.flac", "large", 5, "transcribe", False, "en", True)


# ------------------------------------------------------------------------------------------------------------------ GPU
@needs_reference
@pytest.mark.gpu
def test_reference_do_whisper_over_the_real_engine():
    import willow_inference_server_b200 as pkg
    from tests.gpu_common import model_pair
    from willow_inference_server_b200 import audio, models

    dims, oracle, h = model_pair()
    model = models.Whisper(None, device="cuda", _handles=[h])
    d = model.dims
    # ---- one short utterance: the tokens do_whisper decodes are the tokens of a direct engine call on the same features
    pcm = _synth(61440, 7)
    ns = _namespace(pkg, audio, model, d, pcm)
    lang, text, ms, translation, speedup, dur = ns["do_whisper"]("x.flac", "large", 5, "transcribe", False, "en")
    mel = audio.log_mel_spectrogram(audio.pad_or_trim(pcm)).numpy()[None]
    want, _ = h.generate(mel, [PROMPT], beam_size=5)
    assert text == " ".join(str(t) for t in want[0]) and lang == "en" and dur == 3840 and translation is None
    # ---- 75 s: chunked, long-mode beam 3, two windows per call, stitched by the LCS merge; detect_language first
    pcm = _synth(75 * 16000, 8)
    ns = _namespace(pkg, audio, model, d, pcm)
    lang, text, *_ = ns["do_whisper"]("x.flac", "base", 5, "transcribe", True, None)
    mels, strides = audio.log_mel_chunks(pcm)
    top = model.detect_language(models.StorageView.from_array(mels[:1]))[0][0][0]
    assert lang == top.strip("<|>")
    prompt = ns["models"].whisper_processor.tokenizer.convert_tokens_to_ids(["<|startoftranscript|>", top, "<|transcribe|>", "<|notimestamps|>"])
    seqs, _ = h.generate(mels, [prompt] * mels.shape[0], beam_size=3)
    merged = audio.find_longest_common_sequence(list(zip(seqs, strides)), ns["models"].whisper_processor.tokenizer)
    assert text == " ".join(str(int(t)) for t in merged)
    # ---- do_translate on the features do_whisper left in `gpu_features`
    out = ns["do_translate"](model, models.StorageView.from_array(mel), 1, "<|en|>", 5)
    tr, _ = h.generate(mel, [[PROMPT[0], PROMPT[1], d["translate"], PROMPT[3]]], beam_size=5)
    assert out == " ".join(str(t) for t in tr[0])
