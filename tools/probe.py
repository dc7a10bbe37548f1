"""Scratch performance probe (not part of the bench contract): times the training epoch loop and the
ranking at a named shape and prints the derived rates."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kb2e_b200  # noqa: E402
from kb2e_b200 import kg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="fb15k")
ap.add_argument("--model", default="transe")
ap.add_argument("--dim", type=int, default=100)
ap.add_argument("--distance", type=int, default=1)
ap.add_argument("--method", type=int, default=1)
ap.add_argument("--epochs", type=int, default=20)
ap.add_argument("--test", type=int, default=2000)
ap.add_argument("--random", action="store_true", help="random triples instead of the planted KG")
args = ap.parse_args()

nE, nR, ntr, nva, nte, _ = kg.SHAPES[args.shape]
t0 = time.time()
if args.random:
    rng = np.random.default_rng(0)
    mk = lambda n: np.stack([rng.integers(0, nE, n, dtype=np.int32), rng.integers(0, nE, n, dtype=np.int32),
                             rng.integers(0, nR, n, dtype=np.int32)], 1)
    g = {"nE": nE, "nR": nR, "train": mk(ntr), "valid": mk(nva), "test": mk(nte)}
else:
    g = kg.make_kg(args.shape, seed=0)
print("kg %.1fs" % (time.time() - t0), flush=True)
with kb2e_b200.Context(args.model, args.dim, nE, nR, method=args.method, distance=args.distance, batches=100,
                       rate=0.01, margin=1.0, seed=1) as ctx:
    t0 = time.time()
    ctx.set_train_triples(g["train"])
    ctx.set_bern(*(kg.bern_stats(g["train"], nR) if args.method == 1 else (None, None)))
    ctx.init_embeddings()
    print("setup %.3fs" % (time.time() - t0), flush=True)
    done = 0
    for chunk in ((1, 1, args.epochs, args.epochs) if args.shape != "scaled" else (1, args.epochs)):
        s0 = ctx.train_stats()
        t0 = time.time()
        loss = ctx.train_epochs(done, chunk)
        wall = time.time() - t0
        s1 = ctx.train_stats()
        done += chunk
        ms = s1["kernel_ms"] - s0["kernel_ms"]
        n = s1["samples"] - s0["samples"]
        alpha = (s1["active"] - s0["active"]) / n
        U = (s1["touched_ent"] - s0["touched_ent"]) + (s1["touched_rel"] - s0["touched_rel"])
        P = (args.dim + 3) // 4 * 4
        rows = 5 if args.model == "transh" else 4
        bytes_ = n * ((rows + 2 * rows * alpha) * args.dim * 4 + 12) + 3 * U * args.dim * 4
        print("epochs %d: kernel %.3f ms (wall %.3f ms) -> %.1f M triples/s, %.2f us/batch, alpha %.3f, U/batch %.0f, "
              "alg %.1f GB/s, loss %.2f -> %.2f" % (chunk, ms, wall * 1e3, n / ms / 1e3, ms * 1e3 / (chunk * 100), alpha,
                                                    U / (chunk * 100), bytes_ / ms / 1e6, loss[0], loss[-1]), flush=True)
    ctx.set_test_triples(g["test"][:args.test])
    ctx.add_filter_triples(g["train"])
    ctx.add_filter_triples(g["valid"])
    for _ in range(2):
        r0 = ctx.rank_stats()
        t0 = time.time()
        res = ctx.rank(want_ranks=False)
        wall = time.time() - t0
        r1 = ctx.rank_stats()
        q = r1["queries"] - r0["queries"]
        ms, mms = r1["kernel_ms"] - r0["kernel_ms"], r1["main_kernel_ms"] - r0["main_kernel_ms"]
        print("rank %d queries: kernels %.3f ms (main %.3f ms, wall %.1f ms) -> %.0f queries/s (main kernel only %.0f), "
              "raw MR %.1f filt MR %.1f hits@10 %.3f/%.3f" % (q, ms, mms, wall * 1e3, q / ms * 1e3, q / mms * 1e3,
                                                              res["sums"][0] / q, res["sums"][1] / q, res["sums"][2] / q,
                                                              res["sums"][3] / q), flush=True)
