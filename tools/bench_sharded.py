"""Rank-sharded local matvec (ttipm_b200.sharded) on the scaled rank grid, N GPUs of one node:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 tools/bench_sharded.py
Every rank holds x, P1 and the operator cores, 1/N of every P2; a step = one matvec + the NCCL all-gather of the output
slabs.  Time = CUDA events per rank, MAX over ranks.  Prints one JSON line per shape on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200")]
from ttipm_b200 import get_runtime, replicas, sharded  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rt = get_runtime()
    out = []
    for r, s in ((256, 16), (256, 32), (128, 16)):
        rng = np.random.default_rng(0)                      # identical operands on every rank
        ranks = {(0, 0): s, (0, 1): s, (1, 2): 1, (2, 1): s, (2, 2): s}
        up = rt.to_device
        terms, flops = [], 0.0
        for (i, j), q in ranks.items():
            A, P1, P2 = up(rng.standard_normal((q, 4, 4, q))), up(rng.standard_normal((r, q, r))), up(rng.standard_normal((r, q, r)))
            f = 2.0 * r * 4 * r * r * q + 2.0 * r * r * q * 16 * q + 2.0 * r * 4 * r * r * q
            terms.append((P1, A, P2, j, i))
            flops += f
            if (i, j) == (0, 1):
                terms.append((P1.permute(2, 1, 0), A.permute(0, 2, 1, 3), P2.permute(2, 1, 0), 0, 1))
                flops += f
        x = up(rng.standard_normal((r, 3, 4, r)))
        op = sharded.ShardedBlockMatvec(terms, 3, (r, r), rt=rt)
        if "SHARD_FAKE_WORLD" in os.environ:                 # single-GPU debugging of one rank's slab of a larger job
            fw = int(os.environ["SHARD_FAKE_WORLD"])
            op.world, op.rank = fw, fw - 1
            op.lo, op.hi, op.per = sharded.slab(r, fw - 1, fw)
            op.terms = sharded.K.TermList()
            for t in terms:
                op.terms.add(t[0], t[1], t[2][op.lo:op.hi], t[3], t[4])
            op.world = 1
            op.L = op.per
        for _ in range(3):
            y = op(x)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        iters = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            y = op(x)
        e1.record()
        torch.cuda.synchronize(dev)
        us = e0.elapsed_time(e1) * 1e3 / iters
        us_max = replicas.max_over_ranks([us], device=dev)[0]
        chk = float(y.double().abs().sum())
        if rank == 0:
            rec = dict(kernel="sharded_block_matvec", n_gpus=world, r=r, s=s, us=us_max, tflops=flops / us_max * 1e-6,
                       allgather_bytes_per_rank=8 * r * 3 * 4 * op.per, checksum=chk)
            print(json.dumps(rec), flush=True)
            out.append(rec)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"bench_sharded_n{world}.jsonl"), "w") as f:
            for rec in out:
                f.write(json.dumps(rec) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
