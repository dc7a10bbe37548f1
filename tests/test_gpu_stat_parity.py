"""Trained-model parity (north_star, third check): TransE / TransH / TransR trained by the B200 path vs
(a) the unmodified reference and (b) the reference's bit-pinned update rule driven by the uniform counter
sampler, on identical synthetic KGs, 3 seeds each (tests/golden/stat_parity.npz, produced by
`python tools/stat_parity.py --stage ref` in the build container).

Findings this test pins (full numbers: profiles/stat_parity_r01.json):
  * our exact ranking of the reference's trained tables reproduces the reference's own CPU evaluation exactly;
  * against (b) the 3-seed mean filtered MeanRank agrees within the reference's own seed-to-seed spread
    (4-8 % on this 2,000-entity KG; the 2 % target of north_star is below what 3 seeds x 2,000 queries resolve);
  * against (a) the GPU models are consistently a few per cent BETTER: the reference's randMax
    (common/utils.cpp:113-120) multiplies two rand() values in int, so e.g. 75 % of its sampled indices are
    even -- a defect SURVEY.md A.1 says not to replicate."""
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_three_seed_trained_model_parity(gpu_lib, tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import stat_parity
    out = str(tmp_path / "sp.json")
    stat_parity.stage_gpu(out)
    res = json.load(open(out))
    for row in res["rows"]:
        # same tables, same exact arithmetic: our ranking of the reference's model == the reference's own evaluation
        assert abs(row["reference"]["filt_mr"] - row["reference_cpu_eval"]["filt_mr"]) < 1e-9
        assert abs(row["reference"]["raw_mr"] - row["reference_cpu_eval"]["raw_mr"]) < 1e-9
        assert abs(row["reference"]["filt_h10"] - row["reference_cpu_eval"]["filt_h10"]) < 1e-12
    for name, s in res["summary"].items():
        tol = max(0.08, 2.0 * s["ref_seed_spread_rel"])
        assert abs(s["mr_rel_diff_vs_uniform"]) <= tol, (name, s)
        assert abs(s["h10_diff_points_vs_uniform"]) <= 2.5, (name, s)
        assert s["mr_rel_diff"] <= 0.03, (name, s)   # never worse than the shipped reference beyond noise
        assert s["gpu_filt_mr"] < 0.25 * 1000          # far better than chance (N_E / 2 = 1000)


def test_trained_model_parity_at_fb15k_shape(gpu_lib, tmp_path):
    """north_star check 3 at a BASELINE shape: BASELINE configs[1] (TransE bern squared-L2 size=100, FB15k-shape KG, all
    118,142 filtered-ranking queries), 100 epochs on both sides, 3 data seeds.  The reference side was trained and ranked
    on the host in the build container (tools/stat_parity_large.py --stage ref; metrics in tests/golden/stat_parity_fb15k.json):
      shipped   the unmodified reference program                                   <- the parity target
      uniform   the reference's update rule + the uniform counter sampler
    Asserted at the tolerance north_star states -- filtered MeanRank within 2 %, Hits@10 within 0.5 points, 3-seed means:
      * kb2e_b200 with KB2E_FLAG_SAMPLER_RANDMAX (indices drawn with the distribution of the reference's randMax) vs shipped;
      * kb2e_b200 with its default uniform sampler vs the reference's update rule under the same uniform sampler.
    The default sampler against the SHIPPED reference is ~4 % better in MeanRank (reported, asserted one-sided): the
    reference's randMax multiplies two rand() values in int, so e.g. 75 % of its triple indices are even -- the host-side
    arms show the same 4 % between `shipped` and `uniform` with the update rule held fixed."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import stat_parity_large
    out = str(tmp_path / "spl.json")
    res = stat_parity_large.stage_gpu(out)
    s = res["summary"]
    for key in ("gpu_randmax_vs_shipped", "gpu_uniform_vs_uniform"):
        assert abs(s[key]["mr_rel_diff"]) <= 0.02, (key, s[key])
        assert abs(s[key]["h10_diff_points"]) <= 0.5, (key, s[key])
    assert s["gpu_uniform_vs_shipped"]["mr_rel_diff"] <= 0.0   # never worse than the shipped reference
    keep = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(keep):
        json.dump(res, open(os.path.join(keep, "r02_stat_parity_fb15k.json"), "w"), indent=1, default=float)
