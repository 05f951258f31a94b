"""Rank-sharded local block matvec (BASELINE north star: "the rank dimension of the interface
contractions split across GPUs with an NCCL all-gather").

The three stages of K1 -- x P2, then the operator core, then P1 (reference cy_src/lgmres_cy.pyx:146-153,
src/tt_als.py:193) -- all carry the OUTPUT rank index Lam of the right interface P2[Lam, S, R] untouched, so a split of
Lam shards every stage with no recomputation: rank g holds the full x, P1 and operator cores and only its slab
P2[Lam_g, :, :] (for a transposed alias: the slab of the permuted view), computes y[..., Lam_g] with the ordinary
`ttipm_block_matvec` kernel, and the slabs are all-gathered because the next Krylov step needs the whole vector on
every rank.  One process per GPU, torch.distributed for the plumbing (nccl over NVLink on GPUs, gloo in the CPU tests);
the collective moves b * l * n * L / G doubles per rank and matvec.

At the shipped problem sizes (interfaces of KBs..MBs, one matvec ~ 50 us) this does not pay -- the path then runs as
independent replicas (ttipm_b200.replicas); the sharded form is for the scaled rank grid of SURVEY 8d.
"""
import torch
import torch.distributed as dist

from . import kernels as K
from .runtime import get_runtime


def _world(group):
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def slab(L, rank, world):
    """[lo, hi) of the output rank axis owned by `rank`; every rank gets ceil(L / world) or is padded."""
    per = (L + world - 1) // world
    return min(L, rank * per), min(L, (rank + 1) * per), per


class ShardedBlockMatvec:
    """y = block_local_product(x) with the right-interface rank axis split over `group`.

    terms: iterable of (P1, A, P2, in_block, out_block[, alpha]) with FULL device tensors (P2 may be a permuted view);
    the constructor keeps only this rank's P2 slab."""

    def __init__(self, terms, nb_out, out_ranks, group=None, rt=None):
        self.rt = rt or get_runtime()
        self.group = group
        self.rank, self.world = _world(group)
        self.nb_out = int(nb_out)
        self.l, self.L = out_ranks
        self.lo, self.hi, self.per = slab(self.L, self.rank, self.world)
        self.terms = K.TermList()
        for t in terms:
            P1, A, P2, ib, ob = t[:5]
            alpha = t[5] if len(t) > 5 else 1.0
            assert P2.shape[0] == self.L
            if self.hi > self.lo:
                self.terms.add(P1, A, P2[self.lo:self.hi], ib, ob, alpha)

    def local(self, x):
        """This rank's slab y[..., lo:hi] as a contiguous (l, nb_out, n, per) tensor (zero padded)."""
        n = x.shape[2]
        out = self.rt.zeros(self.l, self.nb_out, n, self.per)
        if self.hi > self.lo:
            y = K.block_matvec(self.terms, x, self.nb_out, (self.l, self.hi - self.lo), rt=self.rt)
            out[..., : self.hi - self.lo] = y
        return out

    def __call__(self, x):
        mine = self.local(x)
        if self.world == 1:
            return mine[..., : self.L]
        flat = torch.empty(self.world * mine.numel(), dtype=mine.dtype, device=mine.device)
        dist.all_gather_into_tensor(flat, mine.reshape(-1), group=self.group)
        gathered = flat.view((self.world,) + tuple(mine.shape))
        # (G, l, b, n, per) -> (l, b, n, G * per) and drop the padding of the last slab
        full = gathered.permute(1, 2, 3, 0, 4).reshape(self.l, self.nb_out, mine.shape[2], self.world * self.per)
        return full[..., : self.L].contiguous()
