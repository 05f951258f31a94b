"""One traced step-size eigen sweep on the GPU (for ncu captures of k_eig_lanczos / k_eig_assemble) with timing:
python tools/prof_eig.py [fixture]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200"), os.path.join(ROOT, "tests"),
                os.path.join(ROOT, "oracle")]
import eigen_cases as EC  # noqa: E402
from ttipm_b200 import eigen as E, get_runtime  # noqa: E402

rt = get_runtime()
path = sys.argv[1] if len(sys.argv) > 1 else [f for f in EC.FILES if "maxcut_10_r1_s41_25" in f][0]
g = EC.load(path)
for rep in range(3):
    stats = {}
    rt.sync()
    t0 = time.perf_counter()
    out = EC.run(g, lambda *a, **k: E.tt_max_generalised_eigen(*a, _stats=stats, **k),
                 lambda *a, **k: E.tt_min_eig(*a, _stats=stats, **k))
    rt.sync()
    print(os.path.basename(path), "seconds", round(time.perf_counter() - t0, 4), "step", out["step"], "reference", g["scalar"], stats)
