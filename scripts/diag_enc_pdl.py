"""Encoder with / without programmatic dependent launch: stage time and bit-identity of the output (large-v2, one window)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from willow_inference_server_b200 import _lib, weights as W  # noqa: E402

dims = W.WhisperDims.for_size("large-v2")
host, _ = bench.make_blob_host(dims, pinned=False)
h = _lib.Handle.from_host(host.numpy(), 0)
del host
pcm = bench.synth_utterance(bench.AUDIO_SAMPLES, 1234)
mel = h.logmel(pcm, [0], [len(pcm)])
P = np.asarray([bench.PROMPT], np.int32)
outs = {}
for pdl in (0, 1, 0, 1):
    h.set_option("enc_pdl", pdl)
    for _ in range(4):
        out, _ = h.generate(mel, P, bench.BEAM, 1.0, 1.0, bench.MAX_LENGTH, [dims.eot])
    t = h.timing()
    outs[pdl] = h.debug_encode(mel)
    print("enc_pdl", pdl, "encoder_ms %.3f cross_kv_ms %.3f generate_ms %.3f" % (t["encoder_ms"], t["cross_kv_ms"], t["generate_ms"]), "tokens", out[0][:5], flush=True)
print("encoder output identical:", bool(np.array_equal(outs[0], outs[1])), "max abs", float(np.abs(outs[0]).max()))
