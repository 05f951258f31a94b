"""CPU tier: the N>1 path (replica sharding + max-over-ranks timing) with gloo, world_size 2."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PKG = os.path.join(ROOT, "tensor-train-interior-point-method_b200")


def _worker(rank, world, port, q):
    sys.path.insert(0, PKG)
    import torch.distributed as dist
    from ttipm_b200 import replicas
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = replicas.assign(list(range(7)), rank, world)
    t = replicas.max_over_ranks([1.0 + rank, 5.0 - rank])
    res = replicas.gather_results({"rank": rank, "items": mine})
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mine, t, res))


def test_replicas_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert outs[0][1] == [0, 2, 4, 6] and outs[1][1] == [1, 3, 5]
    for _, _, t, res in outs:
        assert t == [2.0, 5.0]                              # max over ranks, never a wall-clock mix
        assert [r["rank"] for r in res] == [0, 1]
        assert sorted(res[0]["items"] + res[1]["items"]) == list(range(7))


def test_single_process_is_identity():
    sys.path.insert(0, PKG)
    from ttipm_b200 import replicas
    assert replicas.assign([1, 2, 3], 0, 1) == [1, 2, 3]
    assert replicas.max_over_ranks([3.5]) == [3.5]
    assert replicas.gather_results("x") == ["x"]
