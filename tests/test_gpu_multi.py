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
        q.put((rank, "ok", full if rank == 0 else None, np.concatenate(sig) if rank == 0 else None, ctr))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e), None, None, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", [0, 1, 2])
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
