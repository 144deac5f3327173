#!/usr/bin/env python
"""Arrival of every CTA at every grid barrier of the persistent decoder pass (large-v2, beam 5): which CTAs close each phase,
and how far behind the median they are.  `--simt` traces the SIMT pass instead of the warp-MMA pass."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from willow_inference_server_b200 import _lib, weights as W
import torch
dims = W.WhisperDims.for_size("large-v2")
host, _ = bench.make_blob_host(dims)
h = _lib.Handle.from_host(host.numpy(), 0)
pcm = torch.from_numpy(bench.synth_utterance(bench.AUDIO_SAMPLES, 1234)).cuda()
off, ns = np.zeros(1, np.int64), np.array([bench.AUDIO_SAMPLES], np.int32)
prompts = np.array([bench.PROMPT], np.int32)
h.set_option("mega_trace", 1)
h.set_option("mega_mma", 0 if "--simt" in sys.argv else 1)
for a_ in sys.argv:
    if a_.startswith("--dbg="):
        h.set_option("mega_dbg", int(a_[6:]))
for i in range(3):
    h.logmel(pcm.data_ptr(), off, ns, to_host=False, keep=True, pcm_on_device=True, pcm_dtype=_lib.PCM_F32, B=1)
    ids, _ = h.generate(None, prompts, bench.BEAM, 1.0, 1.0, bench.MAX_LENGTH, [dims.eot], B=1)
print(h.timing())
G = 148
tt = h.debug_read_trace(2048 + 160 * 264).astype(np.int64)
arr = tt[2048:2048 + G * 264].reshape(G, 264)[:, :259]  # arr[cta, k] = arrival at barrier k (k = 0: after the embedding)
names = ["embed"] + ["qkv", "self", "o", "cq", "cross", "co", "fc1", "fc2"] * 32 + ["vocab"]
last = arr.max(axis=0)
dur = np.diff(np.concatenate([[tt[0]], last]))  # time between the last arrivals of consecutive barriers = phase length
print("pass (last arrival to last arrival) us: %.1f" % ((last[-1] - tt[0]) / 1e3))
import collections
by = collections.defaultdict(list)
for k, n in enumerate(names):
    a = arr[:, k]
    by[n].append((dur[k], a.max() - np.median(a), a.max() - np.percentile(a, 90), int(a.argmax())))
for n in ["embed", "qkv", "self", "o", "cq", "cross", "co", "fc1", "fc2", "vocab"]:
    v = np.array([x[:3] for x in by[n]], float)
    who = collections.Counter(x[3] for x in by[n]).most_common(4)
    print("%-6s phase %6.0f ns   last - median arrival %6.0f ns   last - p90 %6.0f ns   last CTAs %s" % (n, np.median(v[:, 0]), np.median(v[:, 1]), np.median(v[:, 2]), who))
k = 1 + 8 * 10  # layer 10
for j, n in enumerate(names[k:k + 8]):
    a = arr[:, k + j] - arr[:, k + j].min()
    print(n, "arrival offsets (ns) by CTA, layer 10:", " ".join(str(int(x)) for x in a))
