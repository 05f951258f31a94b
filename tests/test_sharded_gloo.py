"""CPU tier: the rank-sharded local matvec with gloo, world_size 2, on the emulator build of the kernels:
every rank contracts its slab of the right-interface rank axis, the all-gathered result equals the unsharded product."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _case():
    rng = np.random.default_rng(11)
    r, R, nb, n = 5, 7, 3, 4                     # 7 does not divide by 2: the last slab is padded
    ranks = {(0, 0): (2, 3), (0, 1): (1, 2), (1, 2): (1, 1), (2, 1): (3, 2), (2, 2): (2, 2)}
    A = {k: rng.standard_normal((s, n, n, S)) for k, (s, S) in ranks.items()}
    P1 = {k: rng.standard_normal((r, s, r)) for k, (s, S) in ranks.items()}
    P2 = {k: rng.standard_normal((R, S, R)) for k, (s, S) in ranks.items()}
    x = rng.standard_normal((r, nb, n, R))
    return r, R, nb, A, P1, P2, x


def _worker(rank, world, port, q):
    sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"),
                    os.path.join(ROOT, "tensor-train-interior-point-method_b200")]
    import torch.distributed as dist
    import rt_util
    import tt_oracle as O
    from ttipm_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rt = rt_util.emu_runtime()
    r, R, nb, A, P1, P2, x = _case()
    dev = rt.to_device
    terms = []
    for (i, j) in A:
        a, p1, p2 = dev(A[i, j]), dev(P1[i, j]), dev(P2[i, j])
        terms.append((p1, a, p2, j, i))
        if (i, j) == (0, 1):                     # transposed alias, reference src/tt_als.py:196
            terms.append((p1.permute(2, 1, 0), a.permute(0, 2, 1, 3), p2.permute(2, 1, 0), 0, 1))
    op = sharded.ShardedBlockMatvec(terms, nb, (r, R), rt=rt)
    y = rt.to_host(op(dev(x)))
    bm = O.BlockMatrix({k: [v] for k, v in A.items()}, transposes={(0, 1): (1, 0)})
    want = O.block_local_product(bm, 0, P1, P2, x)
    err = float(np.linalg.norm(y - want) / np.linalg.norm(want))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, err, (op.lo, op.hi, op.per), y.shape))


def test_sharded_matvec_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert outs[0][2] == (0, 4, 4) and outs[1][2] == (4, 7, 4)
    for _, err, _, shape in outs:
        assert err < 1e-10, err                  # north-star kernel tolerance, fp64
        assert tuple(shape) == (5, 3, 4, 7)


def test_slab_partition():
    sys.path.insert(0, os.path.join(ROOT, "tensor-train-interior-point-method_b200"))
    from ttipm_b200 import sharded
    for L in (1, 7, 8, 55, 256):
        for world in (1, 2, 4, 8):
            cover = []
            for g in range(world):
                lo, hi, per = sharded.slab(L, g, world)
                assert hi - lo <= per
                cover += list(range(lo, hi))
            assert cover == list(range(L))
