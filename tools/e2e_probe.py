"""Scratch: where the end-to-end time of the ranking call goes (host-side timers around each C-ABI call;
KB2E_RANK_TIMING=1 adds the phases inside kb2e_rank on stderr)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kb2e_b200  # noqa: E402
from kb2e_b200 import kg, TABLE_ENTITY, TABLE_RELATION  # noqa: E402

g = kg.make_kg("fb15k", seed=0)
nE, nR, D = g["nE"], g["nR"], 100
rng = np.random.default_rng(0)
ent = np.round(rng.normal(0, 0.1, (nE, D)), 6)
rel = np.round(rng.normal(0, 0.1, (nR, D)), 6)
test = tuple(np.ascontiguousarray(g["test"][:, k]) for k in range(3))
filt_all = np.concatenate([g["train"], g["valid"]])
filt = tuple(np.ascontiguousarray(filt_all[:, k]) for k in range(3))
n = len(g["test"])
out = [np.empty(2 * n, dtype=np.int32) for _ in range(4)]
for it in range(4):
    marks = []
    t0 = time.perf_counter()
    def mark(name):
        marks.append((name, (time.perf_counter() - t0) * 1e3))
    e2 = kb2e_b200.Context("transe", D, nE, nR, method=1, distance=1)
    mark("create")
    e2.upload(TABLE_ENTITY, ent); e2.upload(TABLE_RELATION, rel)
    mark("upload")
    e2.set_test_triples(test)
    mark("set_test")
    e2.add_filter_triples(filt)
    mark("add_filter")
    if it >= 2:
        os.environ["KB2E_RANK_TIMING"] = "1"
    e2.rank(0, n, out=out)
    mark("rank")
    e2.rank(0, n, out=out)
    mark("rank again")
    os.environ.pop("KB2E_RANK_TIMING", None)
    e2.close()
    mark("close")
    prev = 0.0
    print("iteration", it, " ".join("%s=%.2f" % (k, v - p) for (k, v), p in zip(marks, [0.0] + [m[1] for m in marks[:-1]])), "total %.2f ms" % marks[-1][1], flush=True)
