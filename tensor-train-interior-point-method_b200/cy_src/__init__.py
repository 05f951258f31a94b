"""Drop-in module names of the reference's Cython extensions backed by ttipm_b200."""
