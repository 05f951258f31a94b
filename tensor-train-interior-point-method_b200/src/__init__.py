"""Drop-in module names of the reference (`src.tt_ops`, `src.tt_als`) backed by ttipm_b200.

Option B of INTEGRATION.md: with this directory ahead of the reference on sys.path and TTIPM_REF_TREE pointing at a
checkout of the reference, the modules this package does NOT replace (`src.tt_ipm`, `src.utils`, `src.baselines`, ...:
the unchanged callers of the hot path) are found in the reference's own src/ directory, while `src.tt_ops` and
`src.tt_als` resolve to the files here."""
import os as _os

_ref = _os.environ.get("TTIPM_REF_TREE")
if _ref and _os.path.isdir(_os.path.join(_ref, "src")):
    __path__.append(_os.path.join(_ref, "src"))
