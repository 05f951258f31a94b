"""Stand-in for petsc4py 3.25.1 (absent offline).  Only the objects that
reference src/tt_ipm.py:101-162 (LGMRESSolver) touches; KSP.solve runs the
oracle's restatement of PETSc's LGMRES (oracle/lgmres_ref.py)."""
from . import PETSc  # noqa: F401
