#!/usr/bin/env python
"""Per-phase timing of the persistent decoder pass (large-v2, beam 5): prints ns per phase type."""
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
if "--mma" in sys.argv:
    h.set_option("mega_mma", 1)
for i in range(3):
    h.logmel(pcm.data_ptr(), off, ns, to_host=False, keep=True, pcm_on_device=True, pcm_dtype=_lib.PCM_F32, B=1)
    ids, _ = h.generate(None, prompts, bench.BEAM, 1.0, 1.0, bench.MAX_LENGTH, [dims.eot], B=1)
print(h.timing())
tt = h.debug_read_trace(2048).astype(np.int64)
ev = tt[1024:1024+2*240].reshape(-1,2)
print('events (id, t ns, dt):', [(int(a), int(b - ev[0,1]), int(b - ev[max(i-1,0),1])) for i, (a, b) in enumerate(ev[:110])])
t = tt[:2*262]
# t[2k] = time barrier k was released (k=0: kernel start), t[2k+1] = time CTA 0 arrived at barrier k+1... (index k -> barrier k+1)
rel = t[0::2]; arr = t[1::2]
names = ["embed"] + ["qkv", "self", "o", "cq", "cross", "co", "fc1", "fc2"] * 32 + ["vocab"]
import collections
work = collections.defaultdict(list); wait = collections.defaultdict(list)
for k, n in enumerate(names):
    work[n].append(arr[k] - rel[k]); wait[n].append(rel[k + 1] - arr[k])
for n in ["embed", "qkv", "self", "o", "cq", "cross", "co", "fc1", "fc2", "vocab"]:
    print("%-6s CTA0 work median %6d ns   barrier wait median %6d ns" % (n, np.median(work[n]), np.median(wait[n])))
print("pass total us", (rel[len(names)] - rel[0]) / 1e3)
