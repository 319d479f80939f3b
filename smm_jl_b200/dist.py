"""Host-side plumbing for one-process-per-GPU runs (torch.distributed carries the control plane:
rendezvous, the NCCL unique id of the library's own communicator, gathering per-rank traces).
The data-path collective (one all-gather of last-accepted records per iteration) lives inside
libsmm_b200.so; nothing here touches chain data during a step."""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from ._abi import SMM_NCCL_ID_BYTES, Trace


def shard_chains(n_chains: int, world_size: int, rank: int) -> np.ndarray:
    """0-based global ids of the chains `rank` owns, in the order of its local columns: rank, rank + world, ...
    Chains are dealt round robin (include/smm_b200.h): the temperature ladder runs along the global id, and a hot
    chain's truncated proposal needs more rejection attempts, so every rank should hold the same mix (SURVEY.md 8e
    shards contiguously; with contiguous blocks the last rank is the straggler of every iteration)."""
    if n_chains % world_size != 0:
        raise ValueError("n_chains must be a multiple of world_size")
    return np.arange(rank, n_chains, world_size)


def owner_of(chain: int, n_chains: int, world_size: int) -> int:
    return chain % world_size


def local_index(chain: int, world_size: int) -> int:
    return chain // world_size


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
    """All ranks receive the full [n][N] trace: column c of rank r's [n][L] block is global chain c * world + r.
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
    return Trace.interleave_ranks(parts)


def slice_trace(full: Trace, chains) -> Trace:
    """the columns `chains` (global ids, e.g. shard_chains(N, world, rank)) of a full trace"""
    chains = np.asarray(chains)
    out = Trace(full.n, len(chains), full.P, full.M)
    for f in Trace.FLOAT_FIELDS + Trace.INT_FIELDS:
        setattr(out, f, np.ascontiguousarray(getattr(full, f)[:, chains]))
    return out


def interleave(per_rank) -> np.ndarray:
    """per-chain vectors of every rank ([L] each, rank order) -> the [N] vector in global chain order"""
    return np.stack([np.asarray(v) for v in per_rank], axis=1).reshape(-1)
