"""Scratch performance probe of the batched training (kb2e_set_replicas): K TransE models of one configuration in one
launch at a named shape; prints per-batch time, aggregate throughput and the algorithmic GB/s."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kb2e_b200  # noqa: E402
from kb2e_b200 import kg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="fb15k")
ap.add_argument("--dim", type=int, default=100)
ap.add_argument("--distance", type=int, default=1)
ap.add_argument("--method", type=int, default=1)
ap.add_argument("--epochs", type=int, default=20)
ap.add_argument("--models", default="1,2,4,8,12")
args = ap.parse_args()
nE, nR, ntr, _, _, _ = kg.SHAPES[args.shape]
g = kg.make_kg(args.shape, seed=0)
bern = kg.bern_stats(g["train"], nR) if args.method == 1 else (None, None)
for K in [int(x) for x in args.models.split(",")]:
    with kb2e_b200.Context("transe", args.dim, nE, nR, method=args.method, distance=args.distance, batches=100, rate=0.01, margin=1.0,
                           seed=1) as ctx:
        if K > 1:
            ctx.set_replicas(K)
        ctx.set_train_triples(g["train"])
        ctx.set_bern(*bern)
        ctx.init_embeddings()
        try:
            ctx.train_epochs(0, 2)
        except kb2e_b200.Kb2eError as e:
            print("K=%d: %s" % (K, e))
            continue
        done = 2
        for chunk in (args.epochs, args.epochs):
            s0 = ctx.train_stats()
            loss = ctx.train_epochs(done, chunk)
            s1 = ctx.train_stats()
            done += chunk
            ms = s1["kernel_ms"] - s0["kernel_ms"]
            n = s1["samples"] - s0["samples"]
            alpha = (s1["active"] - s0["active"]) / n
            U = (s1["touched_ent"] - s0["touched_ent"]) + (s1["touched_rel"] - s0["touched_rel"])
            bytes_ = n * ((4 + 8 * alpha) * args.dim * 4 + 12) + 3 * U * args.dim * 4
            last = np.atleast_2d(loss)[:, -1]
            print("K=%2d epochs %d: %.3f ms -> %.1f M triples/s aggregate, %.2f us/batch, alpha %.3f, alg %.1f GB/s, final loss per model %s"
                  % (K, chunk, ms, n / ms / 1e3, ms * 1e3 / (chunk * 100), alpha, bytes_ / ms / 1e6, np.round(last[:4], 1)), flush=True)
