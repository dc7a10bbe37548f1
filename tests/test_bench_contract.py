"""CPU checks of bench.py's bookkeeping (the measurement itself needs a B200): the algorithmic-bytes formula of
SURVEY.md 8d, the reference arm's multi-rank behaviour, and the flags the driver passes."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_algorithmic_bytes_match_survey_8d():
    import bench
    # TransE, s = 4: B_E = (4 + 8 alpha) * D * 4 + 12 per pair, plus 3 * U * D * 4 for the publish
    b, alpha = bench.algorithmic_bytes(1000, 1000, 0, 100)
    assert alpha == 1.0 and b == 1000 * (1612 + 3200)
    b, alpha = bench.algorithmic_bytes(1000, 0, 0, 50)
    assert alpha == 0.0 and b == 1000 * 812
    b, alpha = bench.algorithmic_bytes(2000, 500, 300, 200)
    assert alpha == 0.25 and b == 2000 * (3212 + 6400 * 0.25) + 3 * 300 * 200 * 4
    # TransH gathers / updates five rows
    b, _ = bench.algorithmic_bytes(10, 10, 0, 100, rows_per_pair=5)
    assert b == 10 * (2012 + 4000)


def test_reference_arm_is_silent_on_other_ranks():
    """Under torchrun only rank 0 runs the (single-process CPU) reference; the other ranks exit 0 without output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout)


def test_committed_bench_lines_carry_the_contract_keys():
    """The round's committed bench lines (profiles/) have every key the driver and the judge read."""
    for name in ("r01_bench.json", "r01_bench_2gpu.json", "r01_bench_8gpu.json"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        d = json.loads(open(path).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "eval"):
            assert k in d, (name, k)
        assert d["config"]["workload"] and d["gpu_launches"] > 0 and d["vs_baseline"] is None
        for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
            assert k in d["e2e"] and k in d["eval"]["e2e"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in d["roofline"] and k in d["eval"]["roofline"]
        assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
