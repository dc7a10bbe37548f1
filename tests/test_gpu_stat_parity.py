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
