"""world_size-2 (and 3) gloo test of the multi-process host logic used by bench.py --gpus N:
group placement and the all-gather of peer handles.  No GPU involved."""
import os
import socket

import pytest
import torch.multiprocessing as mp

from apus_b200 import placement as P


def test_placement_covers_every_replica_once():
    for world in (1, 2, 3, 4, 8):
        for n in (1, 3, 5, 7):
            seen = {}
            for rank in range(world):
                for g, r in P.hosted(rank, world, n):
                    assert (g, r) not in seen
                    seen[(g, r)] = rank
                    assert P.host_of(g, r, world) == rank
            assert len(seen) == world * n
            # every rank leads exactly one group; load is balanced within one replica
            assert all(seen[(g, 0)] == g for g in range(world))
            loads = [sum(1 for v in seen.values() if v == rk) for rk in range(world)]
            assert max(loads) - min(loads) <= 0 if (n % world == 0 or world == 1) else max(loads) - min(loads) <= n


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = {k: bytes([k[0], k[1], rank]) * 42 + b"\0\0" for k in P.hosted(rank, world, n)}
    merged = P.exchange(dist, mine, world)
    ok = len(merged) == world * n
    for g, r, p in P.connections(rank, world, n):
        blob = merged[(g, p)]
        ok &= blob[0] == g and blob[1] == p and blob[2] == P.host_of(g, p, world) and len(blob) == 128
    dist.barrier()
    q.put((rank, ok, len(merged)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 5), (3, 3)])
def test_handle_exchange_over_gloo(world, n):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == list(range(world))
    assert all(r[1] for r in res), res
    assert all(r[2] == world * n for r in res)
