"""N > 1 on real GPUs: one process per GPU, chains sharded, one all-gather per iteration; the gathered
trace must equal the oracle's single-process run (and therefore the 1-GPU run)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_cfg(kind, n_chains, n_iter, data_mom=None, **kw):
    from smm_jl_b200 import configs
    if kind == "panel":
        return configs.dynamic_panel(n_chains, n_iter, 8, 10, 150, data_mom=data_mom, sigma_update_steps=5, **kw)
    return configs.mvnormal(n_chains, n_iter, **kw)


def _worker(rank, world, port, n_chains, n_iter, mode, q, kind="mvnormal", data_mom=None):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from smm_jl_b200 import _lib, configs, dist as sd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idb = sd.broadcast_id(_lib.nccl_unique_id)
        cfg = _make_cfg(kind, n_chains, n_iter, data_mom, device=rank, world_size=world, rank=rank, nccl_id=idb,
                        exchange_mode=mode)
        with _lib.BGPHandle(cfg) as h:
            h.step(n_iter // 2)
            h.step(n_iter - n_iter // 2)
            tr = h.read_trace(1, n_iter)
            sigma, acc = h.chain_state()
            ctr = h.counters()
        full = sd.gather_trace(tr)
        sig = [None] * world
        dist.all_gather_object(sig, sigma)
        q.put((rank, "ok", full if rank == 0 else None, sd.interleave(sig) if rank == 0 else None, ctr))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e), None, None, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_n_gpus_match_the_oracle(smm, oracle, world, mode):
    """world ranks x 16 chains: NCCL all-gather (mode 0), fused peer stores + flag exchange (mode 1), per-chain
    completion tags to every peer (mode 2) -- the gathered trace equals the oracle's single-process run"""
    from smm_jl_b200 import configs
    if smm.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    n_chains, n_iter = 16 * world, 30
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_chains, n_iter, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == ["ok"] * world, [r[1] for r in res]
    from tests.parity import assert_trace_parity
    ref = oracle.run(configs.mvnormal(n_chains, n_iter), n_iter, n_threads=8)
    assert_trace_parity(res[0][2], ref.trace)
    np.testing.assert_array_equal(res[0][3], ref.sigma)
    assert res[0][4]["swaps"] == ref.swaps
    assert res[0][4]["collectives"] == (n_iter - 1 if mode == 0 else 0)


def test_two_gpus_panel_match_the_oracle(smm, oracle):
    """dynamic-panel objective sharded over 2 GPUs (ncclAllGather per iteration) = the oracle's single-process run"""
    from smm_jl_b200 import configs
    if smm.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, n_chains, n_iter = 2, 12, 20
    dm = configs.panel_data_moments(lambda c, p: oracle.eval_batch(c, p), 8, 10, 150)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_chains, n_iter, 0, q, "panel", dm)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == ["ok"] * world, [r[1] for r in res]
    from tests.parity import assert_trace_parity
    ref = oracle.run(_make_cfg("panel", n_chains, n_iter, dm), n_iter, n_threads=8)
    assert_trace_parity(res[0][2], ref.trace)
    np.testing.assert_array_equal(res[0][3], ref.sigma)
    assert res[0][4]["swaps"] == ref.swaps


def _seq_worker(rank, world, port, plan, q):
    """several handles one after the other in the same process: the communicator and the CUDA-IPC exchange arena of the
    first handle are handed to the later ones (smm_b200.h, smm_shutdown); the last entry of `plan` runs while another
    fused handle is still alive (its own arena)"""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from smm_jl_b200 import _lib, dist as sd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["SMM_ARENA_MIN_BYTES"] = "65536"   # so that the 1024-chain ensemble below outgrows the first arena
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idb = sd.broadcast_id(_lib.nccl_unique_id)
        out = []
        keep = None
        for i, (n_chains, n_iter, mode) in enumerate(plan):
            # later handles carry a garbage id: it must not be looked at once the communicator is cached
            cfg = _make_cfg("mvnormal", n_chains, n_iter, None, device=rank, world_size=world, rank=rank,
                            nccl_id=idb if i == 0 else b"\x01" * 128, exchange_mode=mode, n_sim=600)
            h = _lib.BGPHandle(cfg)
            h.step(n_iter // 2)
            h.step(n_iter - n_iter // 2)
            tr = h.read_trace(1, n_iter)
            sigma, _ = h.chain_state()
            full = sd.gather_trace(tr)
            sig = [None] * world
            dist.all_gather_object(sig, sigma)
            out.append((full, sd.interleave(sig)) if rank == 0 else None)
            if i == len(plan) - 2:
                keep = h          # stays alive while the last handle is created and run
            else:
                h.close()
        if keep is not None:
            keep.close()
        _lib.lib().smm_shutdown()     # and the caches can be ended and rebuilt
        q.put((rank, "ok", out))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e), None))
    finally:
        dist.destroy_process_group()


def test_handles_in_sequence_share_the_communicator_and_arena(smm, oracle):
    from smm_jl_b200 import configs
    world = 2
    if smm.device_count() < world:
        pytest.skip("needs 2 GPUs")
    # (chains, iterations, mode): fused, fused with a larger ensemble (same arena, other offsets), NCCL path on the
    # cached communicator, grid-barrier mode with an ensemble that outgrows the arena (it is remade collectively), and
    # a fused handle created while the previous one is still alive
    plan = [(32, 24, 2), (64, 16, 3), (32, 12, 0), (1024, 6, 1), (32, 20, 3), (48, 12, 2)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_seq_worker, args=(r, world, port, plan, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == ["ok"] * world, [r[1] for r in res]
    from tests.parity import assert_trace_parity
    for (n_chains, n_iter, mode), got in zip(plan, res[0][2]):
        ref = oracle.run(configs.mvnormal(n_chains, n_iter, n_sim=600), n_iter, n_threads=8)
        assert_trace_parity(got[0], ref.trace)
        np.testing.assert_array_equal(got[1], ref.sigma)
