"""N > 1 host logic on CPU: world_size-2 gloo run of the load-time weight broadcast and of the window sharding rule."""
import os
import socket
import subprocess
import sys

import pytest

from willow_inference_server_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from willow_inference_server_b200 import parallel, weights as W
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
dims = W.WhisperDims(d_model=128, n_heads=2, n_enc_layers=1, n_dec_layers=1)
blob = None
if rank == 0:
    t = W.synth_engine_tensors(dims, seed=3)
    buf = np.zeros(W.blob_nbytes(t), np.uint8); W.write_blob_into(buf, dims, t)
    blob = torch.from_numpy(buf)
else:
    blob = torch.empty(0, dtype=torch.uint8)
out = parallel.broadcast_blob(blob, torch.device("cpu"))
d2, t2 = W.read_blob(out.numpy())
assert d2.d_model == 128 and "dec.crosskv.w" in t2
sums = [None] * world
dist.all_gather_object(sums, parallel.checksum(out))
assert len(set(sums)) == 1, sums
lo, hi = parallel.shard_range(7, world, rank)
spans = [None] * world
dist.all_gather_object(spans, (lo, hi))
assert spans == [(0, 4), (4, 7)], spans
print("rank", rank, "ok", out.numel())
dist.destroy_process_group()
"""


def test_shard_range_properties():
    for n in (0, 1, 5, 64, 513):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def test_gloo_world2_broadcast_and_sharding(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
