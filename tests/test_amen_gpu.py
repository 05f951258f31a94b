"""GPU tier: device AMEn sweep vs the oracle on KKT systems traced from reference IPM runs."""
import pytest

import golden_io as G
import rt_util
import amen_cases as AC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", G.amen_files("amen_maxcut_5*") + G.amen_files("amen_maxcut_10*") +
                         G.amen_files("amen_corr_clust_8*") + G.amen_files("amen_max_stable_set_9*"),
                         ids=lambda p: p.split("amen_")[-1][:-4])
@pytest.mark.parametrize("native", [True, False], ids=["native", "python"])
def test_block_amen_matches_oracle(path, native):
    rt = rt_util.cuda_runtime()
    out = AC.run_block_amen(rt, path, native=native)
    print(out)
    assert out["sweeps_dev"] == out["sweeps_oracle"], out
    assert out["solves_dev"] == out["solves_oracle"], out
    # same termination tolerance reached; solutions agree far below the AMEn tolerance
    assert out["res_dev"] <= max(10 * out["res_oracle"], 1e-9), out
    assert out["sol_rel_diff"] < 1e-5, out
