"""Run under torchrun on G GPUs: (1) parity of the entity-partitioned training with the single-GPU kernel on a
small KG, (2) throughput at the scaled shape (BASELINE configs[4]) or any named shape.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/dist_check.py [--shape scaled]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kb2e_b200  # noqa: E402
from kb2e_b200 import kg, TABLE_ENTITY, TABLE_RELATION  # noqa: E402
from kb2e_b200.partitioned import PartitionedTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="scaled")
ap.add_argument("--dim", type=int, default=200)
ap.add_argument("--epochs", type=int, default=2)
ap.add_argument("--skip-parity", action="store_true")
ap.add_argument("--skip-throughput", action="store_true")
ap.add_argument("--abort-test", action="store_true",
                help="failure injection: the last rank never launches; the others must give up at their first cross-GPU barrier "
                     "(KB2E_ERR_PEER after KB2E_DIST_TIMEOUT_MS) instead of hanging their GPUs")
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {"world": world}

if not args.skip_parity:
    g = kg.make_kg("small", seed=3)
    nE, nR, D = g["nE"], g["nR"], 40
    rng = np.random.default_rng(0)
    ent = rng.normal(0, 1.0 / D, (nE, D)).astype(np.float32).astype(np.float64)
    rel = rng.normal(0, 1.0 / D, (nR, D)).astype(np.float32).astype(np.float64)
    hm, tm = kg.bern_stats(g["train"], nR)
    cfg = dict(method=1, distance=1, batches=20, rate=0.01, margin=1.0, seed=5)
    pt = PartitionedTrainer(D, nE, nR, rank, world, local, **cfg)
    pt.set_training_set(g["train"], hm, tm)
    pt.upload_global(ent, rel)
    loss = pt.train_epochs(0, 5)
    e2, r2 = pt.gather_global()
    pt.close()
    if rank == 0:
        with kb2e_b200.Context("transe", D, nE, nR, device=local, **cfg) as ctx:
            ctx.set_train_triples(g["train"])
            ctx.set_bern(hm, tm)
            ctx.upload(TABLE_ENTITY, ent)
            ctx.upload(TABLE_RELATION, rel)
            loss1 = ctx.train_epochs(0, 5)
            e1, r1 = ctx.download(TABLE_ENTITY), ctx.download(TABLE_RELATION)
        out["parity"] = {"loss_single": loss1.tolist(), "loss_partitioned": loss.tolist(),
                         "max_abs_entity_diff": float(np.abs(e1 - e2).max()), "mean_abs_entity_diff": float(np.abs(e1 - e2).mean()),
                         "max_abs_relation_diff": float(np.abs(r1 - r2).max())}
        assert np.allclose(loss1, loss, rtol=1e-3), (loss1, loss)
        assert np.abs(e1 - e2).mean() < 5e-5 and np.abs(r1 - r2).max() < 1e-2

if args.abort_test and world > 1:
    os.environ["KB2E_DIST_TIMEOUT_MS"] = "1500"
    g = kg.make_kg("tiny", seed=3)
    pt = PartitionedTrainer(24, g["nE"], g["nR"], rank, world, local, method=0, distance=1, batches=10, rate=0.01, margin=1.0, seed=5)
    pt.set_training_set(g["train"], None, None)
    pt.init_embeddings()
    pt.train_epochs(0, 1)                       # a healthy collective launch first
    t0 = time.time()
    status = "skipped the launch"
    if rank != world - 1:
        try:
            pt.ctx.dist_train_epochs(1, 1)      # the last rank never shows up
            status = "returned OK (WRONG)"
        except kb2e_b200.Kb2eError as e:
            status = str(e)[:160]
    took = time.time() - t0
    again = None
    if rank != world - 1:
        try:
            pt.ctx.dist_train_epochs(2, 1)
        except kb2e_b200.Kb2eError as e:        # the context refuses further partitioned launches
            again = str(e)[:80]
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, status, round(took, 2), again))
    if rank == 0:
        out["abort_test"] = gathered
        assert all(("(5)" in s and t < 10 and a is not None) for r, s, t, a in gathered if r != world - 1), gathered
    pt.ctx.close()

nE, nR, ntr, _, _, _ = kg.SHAPES[args.shape]
if args.skip_throughput:
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0)
rng = np.random.default_rng(1)   # same seed on every rank: identical synthetic triples (throughput only)
train = (rng.integers(0, nE, ntr, dtype=np.int32), rng.integers(0, nE, ntr, dtype=np.int32), rng.integers(0, nR, ntr, dtype=np.int32))
pt = PartitionedTrainer(args.dim, nE, nR, rank, world, local, method=0, distance=1, batches=100, rate=0.01, margin=1.0, seed=1)
pt.set_training_set(train, None, None)
pt.init_embeddings()
pt.train_epochs(0, 1)
s0 = pt.ctx.train_stats()
t0 = time.time()
loss = pt.train_epochs(1, args.epochs)
wall = time.time() - t0
s1 = pt.ctx.train_stats()
ms = torch.tensor([s1["kernel_ms"] - s0["kernel_ms"]], dtype=torch.float64, device="cuda")
cnt = torch.tensor([s1["samples"] - s0["samples"], s1["active"] - s0["active"], s1["touched_ent"] - s0["touched_ent"] + s1["touched_rel"] - s0["touched_rel"]],
                   dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt)
pt.close()
if rank == 0:
    n, active, U = (float(x) for x in cnt.tolist())
    alpha = active / n
    abytes = n * ((4 + 8 * alpha) * args.dim * 4 + 12) + 3 * U * args.dim * 4
    sec = float(ms.item()) * 1e-3
    out["throughput"] = {"shape": args.shape, "dim": args.dim, "epochs": args.epochs, "kernel_ms_max_over_ranks": float(ms.item()),
                         "wall_s": wall, "triples_per_s": n / sec, "alpha": alpha, "algorithmic_GBps_total": abytes / sec / 1e9,
                         "algorithmic_GBps_per_gpu": abytes / sec / 1e9 / world, "loss": loss.tolist()}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
