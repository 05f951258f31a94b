"""Harness (this container only): make the UNMODIFIED reference importable.

  import ref_env; ref = ref_env.load()   ->  namespace with the reference modules

sys.path order: stand-ins, oracle/_ref (freshly built cy_src/*.so), /root/reference
(src/, psd_system/).  Two environment deltas are neutralised:
the lgmres_cy.pyx:510 return bug (patched at build time) and SciPy >= 1.18
LinearOperator handling of (m,1) vectors in the step-size eigen sweeps.
"""
import os
import sys
import types

# TTIPM_REF_TREE: a maintainer's own checkout of the reference (e.g. on a machine that has both the
# reference tree and a B200); default is the read-only mount of the build container
REF = os.environ.get("TTIPM_REF_TREE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.abspath(os.path.join(HERE, "..", "_ref"))
_loaded = None


def available():
    return os.path.isdir(REF)


def load():
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present (GPU box?)")
    sys.path.insert(0, HERE)
    import build_ref
    build_ref.build()
    for p in (REF, OUT, os.path.join(HERE, "standins")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in list(sys.modules):
        if name == "src" or name.startswith("src.") or name == "cy_src" or name.startswith("cy_src."):
            del sys.modules[name]
    import numpy as np
    import scipy.sparse.linalg as spla
    import src.tt_ops as tt_ops
    import src.tt_als as tt_als
    _RealLO = spla.LinearOperator

    def _lo_factory(shape, matvec=None, rmatvec=None, matmat=None, dtype=None, **kw):
        def mv(v):
            return np.asarray(matvec(np.asarray(v).reshape(-1))).reshape(-1)

        def mm(M):
            M = np.asarray(M)
            return np.stack([mv(M[:, i]) for i in range(M.shape[1])], axis=1)
        return _RealLO(shape, matvec=mv, matmat=mm, dtype=np.float64)

    spla.LinearOperator = _lo_factory
    tt_als.scp.sparse.linalg.LinearOperator = _lo_factory
    import src.tt_ipm as tt_ipm
    import cy_src.tt_ops_cy as tt_ops_cy
    import cy_src.lgmres_cy as lgmres_cy
    ns = types.SimpleNamespace(tt_ops=tt_ops, tt_als=tt_als, tt_ipm=tt_ipm,
                               tt_ops_cy=tt_ops_cy, lgmres_cy=lgmres_cy)
    _loaded = ns
    return ns


def create_problem(name, dim, rank):
    """Import psd_system/<name>/<name>.py::create_problem without running its __main__."""
    import importlib.util
    load()
    path = os.path.join(REF, "psd_system", name, name + ".py")
    spec = importlib.util.spec_from_file_location("refproblem_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.create_problem(dim, rank)
