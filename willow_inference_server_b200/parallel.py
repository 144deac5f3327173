"""Multi-GPU plumbing for the hot path: windows are independent, so ranks only share the weight
blob -- ONE broadcast at load time (NCCL over NVLink on GPUs, gloo in the CPU tests) and no data-path collective.
torch.distributed is plumbing here, nothing in it touches the compute path."""
from __future__ import annotations

import hashlib


def shard_range(n_items: int, world: int, rank: int) -> tuple:
    """Contiguous split of `n_items` independent windows over `world` ranks (the same rule models.Whisper uses for its
    in-process replicas): the first `n_items % world` ranks get one extra item."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def broadcast_blob(blob, device, src: int = 0):
    """Rank `src` passes a 1-D uint8 tensor (host, pinned or not); every rank gets a uint8 tensor on `device` with the same
    bytes.  Works without an initialised process group (single process) too."""
    import torch
    import torch.distributed as dist

    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if not multi:
        return blob.to(device)
    rank = dist.get_rank()
    n = torch.tensor([blob.numel() if rank == src else 0], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    out = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    if rank == src:
        out.copy_(blob)
    dist.broadcast(out, src)
    return out


def checksum(t) -> str:
    return hashlib.sha256(t.detach().cpu().numpy().tobytes()).hexdigest()
