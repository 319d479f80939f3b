"""Host-side plumbing for one-process-per-GPU runs (torch.distributed carries the control plane:
rendezvous, the NCCL unique id of the library's own communicator, gathering per-rank traces).
The data-path collective (one all-gather of last-accepted records per iteration) lives inside
libsmm_b200.so; nothing here touches chain data during a step."""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from ._abi import SMM_NCCL_ID_BYTES, Trace


def shard_range(n_chains: int, world_size: int, rank: int) -> "tuple[int, int]":
    """chains [lo, hi) owned by `rank`: contiguous by id."""
    if n_chains % world_size != 0:
        raise ValueError("n_chains must be a multiple of world_size")
    L = n_chains // world_size
    return rank * L, (rank + 1) * L


def owner_of(chain: int, n_chains: int, world_size: int) -> int:
    return chain // (n_chains // world_size)


def env_rank_world() -> "tuple[int, int, int]":
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_id(make_id, pg=None, device: Optional[str] = None) -> bytes:
    """Rank 0 calls make_id() (-> 128 bytes, smm_nccl_unique_id); everyone returns the same bytes."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(pg)
    t = torch.zeros(SMM_NCCL_ID_BYTES, dtype=torch.uint8, device=device or "cpu")
    if rank == 0:
        raw = bytes(make_id())[:SMM_NCCL_ID_BYTES].ljust(SMM_NCCL_ID_BYTES, b"\0")
        t = torch.tensor(list(raw), dtype=torch.uint8, device=device or "cpu")
    dist.broadcast(t, 0, group=pg)
    return bytes(t.cpu().tolist())


def gather_trace(local: Trace, pg=None, device: Optional[str] = None) -> Trace:
    """All ranks receive the full [n][N] trace: per-rank [n][L] blocks joined in rank (= chain) order.
    `device`: where the collective's tensors live ("cuda" for an NCCL process group; default: host, gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(pg)
    parts = [Trace(local.n, local.L, local.P, local.M) for _ in range(world)]
    for f in Trace.FLOAT_FIELDS + Trace.INT_FIELDS:
        mine = torch.from_numpy(np.ascontiguousarray(getattr(local, f)))
        if device:
            mine = mine.to(device)
        bufs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(bufs, mine, group=pg)
        for r in range(world):
            setattr(parts[r], f, bufs[r].cpu().numpy())
    return Trace.concat_chains(parts)


def slice_trace(full: Trace, lo: int, hi: int) -> Trace:
    out = Trace(full.n, hi - lo, full.P, full.M)
    for f in Trace.FLOAT_FIELDS + Trace.INT_FIELDS:
        setattr(out, f, np.ascontiguousarray(getattr(full, f)[:, lo:hi]))
    return out
