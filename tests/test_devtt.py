"""Device-resident tensor trains: the lazy `TTList` the TT functions return behaves like the reference's
list[np.ndarray] for the unchanged driver code, chains of TT operations stay on the device, and the native TT-algebra
driver (csrc/tt_driver.cu) reproduces the oracle on a chain of the kind src/tt_ipm.py:404-475 builds."""
import copy

import numpy as np
import pytest

import rt_util
import tt_oracle as O
from kernel_cases import rel, _dense_tt


def _rand_tt(rng, ranks, mode):
    rr = [1] + list(ranks) + [1]
    return [rng.standard_normal((a, *mode, b)) for a, b in zip(rr[:-1], rr[1:])]


def _run(rt):
    from ttipm_b200 import tt as T, use_runtime
    from ttipm_b200.devtt import TTList
    rng = np.random.default_rng(11)
    A = _rand_tt(rng, [2, 3, 2], (2, 2))
    B = _rand_tt(rng, [3, 2, 2], (2, 2))
    X = _rand_tt(rng, [2, 2, 3], (2, 2))
    with use_runtime(rt):
        # ---- a chain stays on the device: no NumPy core exists until somebody looks ----------------------------
        np.random.seed(3)
        P = T.tt_fast_mat_mat_mul(A, B, 1e-12)
        S = T.tt_add(X, T.tt_scale(0.5, P))
        R = T.tt_rank_reduce(S, 1e-10)
        assert R is S and isinstance(R, TTList) and not R._live and not P._live
        assert len(R) == 4 and T.tt_ranks(R) == [sh[0] for sh in R.shapes()[1:]] and not R._live
        V = T.tt_reshape(R, (4,))
        assert isinstance(V, TTList) and not V._live and not R._live and V.shapes()[0] == (1, 4, R.shapes()[0][-1])
        ip = T.tt_inner_prod(R, R)
        assert not R._live
        # ---- the same chain with the oracle -----------------------------------------------------------------------
        np.random.seed(3)
        Po = O.tt_fast_mat_mat_mul([c.copy() for c in A], [c.copy() for c in B], 1e-12)
        Ro = O.tt_rank_reduce(O.tt_add(X, O.tt_scale(0.5, Po)), 1e-10)
        assert O.tt_ranks(Ro) == T.tt_ranks(R)
        assert abs(ip - O.tt_inner_prod(Ro, Ro)) <= 1e-10 * abs(ip)
        assert rel(_dense_tt(R), _dense_tt(Ro)) < 1e-10                    # first look: ONE download
        assert R._live and all(isinstance(c, np.ndarray) for c in R)
        assert rel(_dense_tt(V), _dense_tt([c.reshape(c.shape[0], 4, c.shape[-1]) for c in Ro])) < 1e-10
        # ---- list behaviour the driver / problem generators rely on ----------------------------------------------
        E = np.ones((1, 2, 2, 1))
        Q = T.tt_add(X, X)
        cat = [E] + Q + [E]                                                  # plain + lazy + plain (psd_system/graphm)
        assert type(cat) is list and len(cat) == 6 and cat[1].shape == (1, 2, 2, 4)
        Q2, Q3 = T.tt_add(X, X), T.tt_add(X, X)
        both = Q2 + Q3                                                       # lazy + lazy
        assert len(both) == 8 and both[4].shape == (1, 2, 2, 4)
        assert len(copy.deepcopy(T.tt_add(X, X))) == 4 and len(list(T.tt_add(X, X))) == 4
        a, b, c, d = T.tt_add(X, X)                                          # unpacking
        assert a.shape == (1, 2, 2, 4) and T.tt_add(X, X)[-1].shape == (6, 2, 2, 1)
        assert [x.shape for x in reversed(T.tt_add(X, X))][0] == (6, 2, 2, 1)
        # ---- cores modified in place by the caller are seen (checksum), so is a replaced core ---------------------
        Y = T.tt_add(X, X)
        before = T.tt_inner_prod(Y, Y)
        Y[0][...] *= 2.0                                                     # in-place edit of a materialised core
        assert abs(T.tt_inner_prod(Y, Y) - 4.0 * before) <= 1e-10 * abs(before)
        Y[1] = 3.0 * Y[1]
        assert abs(T.tt_inner_prod(Y, Y) - 36.0 * before) <= 1e-9 * abs(before)
        # ---- functions that mutate a plain input list still do -----------------------------------------------------
        plain = [c.copy() for c in O.tt_add(X, X)]
        out = T.tt_rank_reduce(plain, 1e-10)
        assert out is plain and type(plain) is list and T.tt_ranks(plain) == [2, 2, 3]
        # ---- psd rounding: input rebound to the rounded train, result = rounded + factor * I ---------------------
        sym = T.tt_add(X, T.tt_transpose(X))
        sym_o = O.tt_add(X, O.tt_transpose(X))
        got = T.tt_psd_rank_reduce(sym, 0.5)
        want = O.tt_psd_rank_reduce(sym_o, 0.5)
        assert T.tt_ranks(got) == O.tt_ranks(want) and rel(_dense_tt(got), _dense_tt(want)) < 1e-10
        assert T.tt_ranks(sym) == O.tt_ranks(sym_o)
        # transposes / embeddings of lazy trains
        t1 = T.tt_transpose(T.tt_add(X, X))
        assert rel(_dense_tt(t1), _dense_tt(O.tt_transpose(O.tt_add(X, X)))) < 1e-14
        k1 = T.tt_IkronM(T.tt_add(X, X))
        assert rel(_dense_tt(k1), _dense_tt(O.tt_IkronM(O.tt_add(X, X)))) < 1e-14


def test_devtt_emu():
    _run(rt_util.emu_runtime())


@pytest.mark.gpu
def test_devtt_gpu():
    _run(rt_util.cuda_runtime())
