#!/usr/bin/env python
"""Generate tests/golden/whisper_hf_tiny.npz with HF transformers' Whisper (an independent
implementation of the architecture the reference runs inside CTranslate2).

    python scripts/gen_golden_whisper_hf.py

The seeded synthetic weights (willow_inference_server_b200.weights.synth_state_dict) are loaded
into ``WhisperForConditionalGeneration``; we record its encoder output, teacher-forced logits and a
greedy decode (full re-forward per step, same suppress rules as SURVEY.md section 8a row A11).
Tests regenerate the weights/inputs from the seeds stored here and compare the oracle to these.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transformers import WhisperConfig, WhisperForConditionalGeneration  # noqa: E402

from oracle import logmel as om  # noqa: E402
from willow_inference_server_b200 import weights as W  # noqa: E402

CFG = dict(d_model=128, n_heads=2, n_enc_layers=2, n_dec_layers=2)
SEED = 11
EOT_RAMP = (10, 8.0)
PROMPT = [50258, 50259, 50359, 50363]
FORCED = PROMPT + [100, 2000, 30000, 41000, 12, 50000, 7, 999]


def main():
    dims = W.WhisperDims(**CFG)
    sd = W.synth_state_dict(dims, seed=SEED, eot_ramp=EOT_RAMP)
    cfg = WhisperConfig(
        vocab_size=dims.n_vocab, num_mel_bins=80, d_model=dims.d_model,
        encoder_layers=dims.n_enc_layers, encoder_attention_heads=dims.n_heads, encoder_ffn_dim=4 * dims.d_model,
        decoder_layers=dims.n_dec_layers, decoder_attention_heads=dims.n_heads, decoder_ffn_dim=4 * dims.d_model,
        max_source_positions=1500, max_target_positions=448, activation_function="gelu",
        pad_token_id=50257, bos_token_id=50257, eos_token_id=50257, decoder_start_token_id=50258,
    )
    model = WhisperForConditionalGeneration(cfg).eval()
    tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
    tsd["proj_out.weight"] = tsd["model.decoder.embed_tokens.weight"]
    print(model.load_state_dict(tsd, strict=False))
    mel = om.log_mel_batch([om.synth_utterance(61440, 1234), om.synth_utterance(160000, 5)])
    feats = torch.from_numpy(mel)
    with torch.no_grad():
        enc = model.model.encoder(feats).last_hidden_state  # [2,1500,d]
        logits = model(input_features=feats[:1], decoder_input_ids=torch.tensor([FORCED])).logits[0]
        sup = sorted(set(dims.suppress_ids))
        greedy = []
        for b in range(2):
            toks = list(PROMPT)
            out = []
            for s in range(60):
                lg = model(input_features=feats[b : b + 1], decoder_input_ids=torch.tensor([toks])).logits[0, -1].clone()
                lg[sup] = -float("inf")
                if s == 0:
                    lg[dims.suppress_ids_begin] = -float("inf")
                t = int(torch.argmax(lg))
                if t == dims.eot:
                    break
                out.append(t)
                toks.append(t)
            greedy.append(out)
    vocab_idx = np.unique(np.concatenate([np.arange(0, dims.n_vocab, 97), np.arange(50250, dims.n_vocab)]))
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "whisper_hf_tiny.npz"),
        cfg=np.array([CFG["d_model"], CFG["n_heads"], CFG["n_enc_layers"], CFG["n_dec_layers"]]),
        seed=np.int64(SEED), eot_ramp=np.array(EOT_RAMP, np.float64), prompt=np.array(PROMPT), forced=np.array(FORCED),
        enc_rows=np.arange(0, 1500, 25), enc_sub=enc[:, ::25].numpy(),
        vocab_idx=vocab_idx, logits_sub=logits[:, vocab_idx].numpy(),
        greedy0=np.array(greedy[0]), greedy1=np.array(greedy[1]),
    )
    print("greedy lens", [len(g) for g in greedy], greedy[0][:10])


if __name__ == "__main__":
    main()
