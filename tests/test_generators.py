"""Problem-generator helpers of the host mirror (ttipm_b200/generators.py; SURVEY 8b "names the drivers actually
use") against fixtures produced by the reference (make_golden.py generators), and the Option-B layout: every
reference psd_system/<p>/<p>.py imports and runs create_problem against the replacement modules src/, cy_src/."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import golden_io as G
import rt_util

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generators.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def _tt(gold, prefix):
    return [gold[f"{prefix}/{k}"] for k in range(int(gold[prefix + "/n"]))]


def test_sampler_draws_match_reference(gold):
    """tt_random_binary_sym consumes the NumPy RNG in the reference's order and builds the same cores."""
    from ttipm_b200 import generators as GEN
    for key in [k for k in gold if k.startswith("binary_sym/") and k.endswith("/n")]:
        seed, dim, rank = map(int, key.split("/")[1].split("_"))
        np.random.seed(seed)
        ours = GEN.tt_random_binary_sym(dim, rank, skew=-1.0)
        want = _tt(gold, key[:-2])
        assert len(ours) == len(want)
        for a, b in zip(ours, want):
            assert a.shape == b.shape and np.abs(a - b).max() < 1e-13


def test_triangular_split_dense(gold):
    from ttipm_b200 import generators as GEN
    for dim in (1, 2, 4):
        for name, fn in (("tril", GEN.tt_tril_one_matrix), ("triu", GEN.tt_triu_one_matrix)):
            for a, b in zip(fn(dim), _tt(gold, f"{name}/{dim}")):
                assert np.array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))
        dense = GEN.tt_matrix_to_matrix(GEN.tt_tril_one_matrix(dim))
        assert np.array_equal(dense, np.tril(np.ones((2 ** dim, 2 ** dim))))
    m = _tt(gold, "split/in")
    assert np.abs(GEN.tt_matrix_to_matrix(m) - gold["to_matrix"]).max() < 1e-13
    ours, want = GEN.tt_split_bonds([c.copy() for c in m]), _tt(gold, "split/out")
    assert [c.shape for c in ours] == [c.shape for c in want]
    merged = GEN.tt_merge_bonds(ours)                       # gauge free: the product of each split pair
    for a, b in zip(merged, m):
        assert np.abs(a - b).max() < 1e-12


def _graph_check(gold, rt):
    from ttipm_b200 import generators as GEN, use_runtime
    with use_runtime(rt):
        for key in [k for k in gold if k.startswith("graph/") and k.endswith("/dense")]:
            seed, dim, r = map(int, key.split("/")[1].split("_"))
            np.random.seed(seed)
            g = GEN.tt_random_graph(dim, r)
            base = key[:-len("/dense")]
            assert [int(v) for v in GEN.T.tt_ranks(g)] == [int(v) for v in gold[base + "/ranks"]]
            assert np.abs(GEN.tt_matrix_to_matrix(g) - gold[key]).max() < 1e-10
            assert int(np.random.get_state()[2]) == int(gold[base + "/rng_pos"])     # same number of rejection draws


def test_random_graph_emu(gold):
    _graph_check(gold, rt_util.emu_runtime())


@pytest.mark.gpu
def test_random_graph_gpu(gold):
    _graph_check(gold, rt_util.cuda_runtime())


_CHILD = r"""
import importlib.util, json, os, sys
import numpy as np
root, problem, dim, rank, seed, mode = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
pkg = os.path.join(root, "tensor-train-interior-point-method_b200")
sys.path[:0] = [pkg, os.path.join(root, "tests")]
sys.path.append(os.path.join(root, "oracle", "ref_harness", "standins"))     # petsc4py / memory_profiler / ... absent here
import ttipm_b200.runtime as R
if mode == "emu":
    import rt_util
    R._override = rt_util.emu_runtime()
import src.tt_ops, src.tt_als, cy_src.tt_ops_cy, cy_src.lgmres_cy          # the replacement modules (Option B)
assert src.tt_ops.__file__.startswith(pkg), src.tt_ops.__file__
path = os.path.join(os.environ.get("TTIPM_REF_TREE", "/root/reference"), "psd_system", problem, problem + ".py")
spec = importlib.util.spec_from_file_location("refproblem_" + problem, path)
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
np.random.seed(seed)
res = mod.create_problem(dim, rank)
from ttipm_b200.generators import tt_to_tensor
out = []
for item in res:
    if item is None:
        out.append(None)
    elif isinstance(item, dict):
        out.append({str(k): tt_to_tensor(v).ravel().tolist() for k, v in item.items()})
    else:
        out.append(tt_to_tensor(item).ravel().tolist())
print("RESULT" + json.dumps(out))
"""


@pytest.mark.ref
@pytest.mark.parametrize("problem", ["maxcut", "corr_clust", "max_stable_set", "graphm"])
def test_option_b_create_problem(gold, problem, tmp_path):
    """The unmodified psd_system/<p>/<p>.py imports `from src.tt_ops import *` from the replacement modules and
    create_problem builds the same problem instance as with the reference's own modules."""
    from ttipm_b200.generators import tt_to_tensor
    dim, rank, seed = (int(v) for v in gold[f"problem/{problem}/args"])
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    script = tmp_path / "child.py"
    script.write_text(_CHILD)
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    env["TTIPM_REF_TREE"] = "/root/reference"       # src.utils / src.tt_ipm (unchanged callers) come from the reference
    res = subprocess.run([sys.executable, str(script), root, problem, str(dim), str(rank), str(seed), "emu"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("RESULT")][-1]
    ours = json.loads(line[len("RESULT"):])
    assert len(ours) == int(gold[f"problem/{problem}/n"])
    for q, item in enumerate(ours):
        base = f"problem/{problem}/{q}"
        if item is None:
            assert base + "/none" in gold
        elif isinstance(item, dict):
            for key, flat in item.items():
                want = tt_to_tensor(_tt(gold, f"{base}/dict/{key}")).ravel()
                assert np.abs(np.array(flat) - want).max() < 1e-9 * max(1.0, np.abs(want).max())
        else:
            want = tt_to_tensor(_tt(gold, base + "/tt")).ravel()
            flat = item
            assert np.abs(np.array(flat) - want).max() < 1e-9 * max(1.0, np.abs(want).max())
