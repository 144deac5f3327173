#!/usr/bin/env python
"""Event trace of one CTA of the persistent decoder pass over one layer (large-v2, beam 5): SM-clock stamps of thread 0,
printed as (event id, ns since the window opened, ns since the previous event).  --cta=N --layer=N --simt
Event ids: GEMV phase 1 start, 2 activations landed, 3 first weight unit there, 4 main loop done, 5 partial tiles synced,
6 epilogue done; self-attention 20 start, 21 keys gathered, 22 block done, 24 stored; cross-attention 10 start, 11 queries
there, 12 K/V there, 13 walk done, 14 warp partials written, 15 CTA partial in global memory, 16 all splits seen,
17 merged; barrier 30 entered, 31 CTA synced, 32 arrival posted, 33 barrier open."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from willow_inference_server_b200 import _lib, weights as W
import torch
opt = {a.split("=")[0][2:]: int(a.split("=")[1]) for a in sys.argv[1:] if "=" in a}
dims = W.WhisperDims.for_size("large-v2")
host, _ = bench.make_blob_host(dims)
h = _lib.Handle.from_host(host.numpy(), 0)
pcm = torch.from_numpy(bench.synth_utterance(bench.AUDIO_SAMPLES, 1234)).cuda()
off, ns = np.zeros(1, np.int64), np.array([bench.AUDIO_SAMPLES], np.int32)
prompts = np.array([bench.PROMPT], np.int32)
h.set_option("mega_mma", 0 if "--simt" in sys.argv else 1)
h.set_option("mega_trace", 1)
h.set_option("mega_trace_cta", opt.get("cta", 5))
h.set_option("mega_trace_layer", opt.get("layer", 10))
if "dbg" in opt:
    h.set_option("mega_dbg", opt["dbg"])
for i in range(3):
    h.logmel(pcm.data_ptr(), off, ns, to_host=False, keep=True, pcm_on_device=True, pcm_dtype=_lib.PCM_F32, B=1)
    ids, _ = h.generate(None, prompts, bench.BEAM, 1.0, 1.0, bench.MAX_LENGTH, [dims.eot], B=1)
print(h.timing())
tt = h.debug_read_trace(2048).astype(np.int64)
n = int(tt[1024])
ev = tt[1025:1025 + 2 * n].reshape(-1, 2)
GHZ = 1.965
t = ((ev[:, 1] - ev[0, 1]) % (1 << 32)) / GHZ
print("events:", n)
line = []
for i in range(n):
    line.append("(%d, %d, +%d)" % (ev[i, 0], t[i], t[i] - t[i - 1] if i else 0))
    if ev[i, 0] == 33:
        print(" ".join(line))
        line = []
print(" ".join(line))
