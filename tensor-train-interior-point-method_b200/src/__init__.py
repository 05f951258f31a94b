"""Drop-in module names of the reference (`src.tt_ops`, `src.tt_als`) backed by ttipm_b200."""
