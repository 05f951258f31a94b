"""Same module name as reference cy_src/lgmres_cy.pyx."""
from ttipm_b200.lgmres import BaseMatVec, IneqMatVecWrapper, MatVecWrapper  # noqa: F401
