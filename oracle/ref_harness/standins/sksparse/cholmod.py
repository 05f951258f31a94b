"""Stand-in: the reference imports sksparse.cholmod at module import time only
(src/tt_als.py:10); SpCholInv is dead code."""


def cholesky(*a, **k):
    raise NotImplementedError("sksparse stand-in")
