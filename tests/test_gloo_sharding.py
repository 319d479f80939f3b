"""world_size-2 `gloo` test of the N > 1 host path on CPU: rendezvous, broadcast of the library's NCCL id,
chain sharding and trace gathering.  The oracle's full-run trace stands in for the per-rank device traces
(no GPU here), so the test checks that the round-robin shards, gathered, reproduce the single-process run."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle_lib
    from smm_jl_b200 import _lib, configs, dist as sd
    from smm_jl_b200._abi import Trace
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the id of the library's own communicator travels over torch.distributed
        idb = sd.broadcast_id(_lib.nccl_unique_id)
        ids = [None] * world
        dist.all_gather_object(ids, idb)
        assert len(idb) == 128 and all(i == idb for i in ids) and any(b != 0 for b in idb)
        # 2. config for this rank: same problem, own rank
        cfg = configs.mvnormal(8, 6, n_sim=200, world_size=world, rank=rank, nccl_id=idb)
        cs = cfg.c_struct()
        assert cs.rank == rank and cs.world_size == world and bytes(cs.nccl_id) == idb
        mine_ids = sd.shard_chains(cfg.n_chains, world, rank)
        assert mine_ids.tolist() == list(range(rank, 8, world))
        # 3. every rank's shard of the (deterministic) run, gathered back = the single-process run
        full = oracle_lib.run(configs.mvnormal(8, 6, n_sim=200), 6).trace
        mine = sd.slice_trace(full, mine_ids)
        got = sd.gather_trace(mine)
        for f in Trace.FLOAT_FIELDS + Trace.INT_FIELDS:
            assert np.array_equal(getattr(got, f), getattr(full, f), equal_nan=True), f
        sig = [None] * world
        dist.all_gather_object(sig, np.asarray(cfg.sigma0)[mine_ids])
        assert np.array_equal(sd.interleave(sig), np.asarray(cfg.sigma0))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
