#!/usr/bin/env python
"""bench.py -- the KB2E hot path on B200: margin-ranking SGD epochs + filtered link-prediction ranking.

Workload (BASELINE.json configs[1]): TransE, bern, squared-L2, size=100, rate=0.01, margin=1,
batches=100 on a synthetic planted-translation KG of FB15k shape (14,951 entities / 1,345 relations /
483,142 train / 50,000 valid / 59,071 test), then filtered MeanRank / Hits@10 over the 59,071 test
triples (118,142 queries x 14,951 candidates).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A training "step" is one epoch (100 batches x 4,831 sampled (positive, negative) pairs); the
headline `value` is training pairs ("triples") per second with everything resident in HBM, timed
per step with CUDA events on the launching stream, L2 flushed before every step, max over ranks.
`e2e` is the same epoch through the C ABI with HOST buffers: triples and fp64 tables copied in from
pinned host memory, fp64 tables copied back.  The ranking pass is reported in the `eval` block of the
same JSON line with its own value / e2e / roofline / cpu_baseline.

Multi-GPU (--gpus N under torchrun): training at FB15k shape does not shard ("replicas only",
DESIGN.md): every rank trains its own replica (different -seed), value = N x pairs / max time.
Ranking shards the test triples across ranks (tables + filter replicated) and all-reduces the four
int64 sums over NCCL.  The `partitioned` block adds BASELINE configs[4] (TransE at the scaled shape,
4 M entities x 200, 100 M triples, one timed epoch for the whole job): at N > 1 the entity-partitioned
trainer, at N = 1 the single-GPU kernel with the roofline fraction of that genuinely HBM-bound shape.

--impl reference times the UNMODIFIED reference (oracle/_ref, compiled from /root/reference by
oracle/Makefile) on the host cores of this box: one epoch of its bfgs() per step; evalCorruption on
a bounded sample of the queries.  The reference is single-threaded.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(model="transe", dim=100, method=1, distance=1, rate=0.01, margin=1.0, batches=100, shape="fb15k")
WORKLOAD = "TransE bern L2 size=100 rate=0.01 margin=1 batches=100, synthetic FB15k-shape KG (14951/1345/483142/50000/59071) + filtered ranking of 59071 test triples"
HBM_FALLBACK_GBS = 6650.0


def bench_config(world):
    """The `config` object of the JSON line -- the SAME dict in the product arm and in the reference arm."""
    return {"workload": WORKLOAD,
            "step": "one epoch = 100 batches x 4831 sampled (positive, negative) pairs",
            "parallelism": "replicas only (training), test triples sharded (ranking)" if world > 1 else "single GPU",
            "l2": "flushed before every timed step (256 MiB write); the 13 MB of tables are L2-resident within a step by nature of the workload",
            "timing": "CUDA events on the launching stream around each persistent launch, max over ranks (product arm); "
                      "wall clock of the reference's own bfgs() on one host core (reference arm)"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return HBM_FALLBACK_GBS, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8:
                self.rows.append(f)

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        busy = [x for x in sm if x > 0.5 * float(self.rows[0][2])] or sm
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[4 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": float(self.rows[0][2]), "reasons": reasons, "samples": len(self.rows)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def algorithmic_bytes(n_pairs, active, touched_rows, dim, rows_per_pair=4):
    """per pair (rows + 2*rows*alpha) * D * 4 + 12 bytes, plus 3*U*D*4 for the publish."""
    alpha = active / max(n_pairs, 1)
    return n_pairs * ((rows_per_pair + 2 * rows_per_pair * alpha) * dim * 4 + 12) + 3 * touched_rows * dim * 4, alpha


# ------------------------------------------------------------------------------------------ product arm
def run_product(args):
    import torch
    import torch.distributed as dist

    import kb2e_b200
    from kb2e_b200 import kg, TABLE_ENTITY, TABLE_RELATION

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hbm_peak, tensor_peak, peak_src = peaks()
    nE, nR, n_train, n_valid, n_test, _ = kg.SHAPES[CFG["shape"]]

    # ---- synthetic KG: generated on rank 0 (torch on the GPU for the exact nearest neighbours), broadcast
    if rank == 0:
        g = kg.make_kg(CFG["shape"], seed=0)
        blob = torch.from_numpy(np.concatenate([g["train"], g["valid"], g["test"]]).astype(np.int32)).cuda()
    else:
        blob = torch.empty((n_train + n_valid + n_test, 3), dtype=torch.int32, device="cuda")
    if world > 1:
        dist.broadcast(blob, 0)
    allt = blob.cpu().numpy()
    train, valid, test = allt[:n_train], allt[n_train:n_train + n_valid], allt[n_train + n_valid:]
    head_mean, tail_mean = kg.bern_stats(train, nR)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(a):
        if world == 1:
            return a
        t = torch.from_numpy(np.asarray(a, dtype=np.int64)).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    ctx = kb2e_b200.Context(CFG["model"], CFG["dim"], nE, nR, method=CFG["method"], distance=CFG["distance"],
                            batches=CFG["batches"], rate=CFG["rate"], margin=CFG["margin"], seed=1 + rank, device=local)
    ctx.set_train_triples(train)
    ctx.set_bern(head_mean, tail_mean)
    ctx.init_embeddings()
    pairs_per_step = CFG["batches"] * (n_train // CFG["batches"])
    K, W = args.steps, args.warmup

    clocks = ClockSampler(local)
    with clocks:
        # ---- training, device-resident -------------------------------------------------------------
        for e in range(W):
            ctx.train_epochs(e, 1)
        s0 = ctx.train_stats()
        step_ms, wall_ms = [], []
        for e in range(W, W + K):
            flush.fill_(1)          # L2 flush between timed steps
            barrier()
            t0 = time.perf_counter()
            before = ctx.train_stats()["kernel_ms"]
            ctx.train_epochs(e, 1)  # one persistent launch = 100 batches; events on the launching stream inside
            torch.cuda.synchronize()
            wall_ms.append((time.perf_counter() - t0) * 1e3)
            step_ms.append(max_over_ranks(ctx.train_stats()["kernel_ms"] - before))
        s1 = ctx.train_stats()
        total_ms = sum(step_ms)
        value = world * pairs_per_step * K / (total_ms * 1e-3)
        n_pairs = s1["samples"] - s0["samples"]
        touched = (s1["touched_ent"] - s0["touched_ent"]) + (s1["touched_rel"] - s0["touched_rel"])
        abytes, alpha = algorithmic_bytes(n_pairs, s1["active"] - s0["active"], touched, CFG["dim"])
        launches = int(s1["launches"] - s0["launches"])
        local_kernel_ms = s1["kernel_ms"] - s0["kernel_ms"]
        achieved = abytes / (local_kernel_ms * 1e-3) / 1e9

        # ---- training, end to end through the C ABI with host buffers ---------------------------------
        def pinned(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t, t.numpy()

        ent_host = ctx.download(TABLE_ENTITY)
        rel_host = ctx.download(TABLE_RELATION)
        keep = [pinned(train[:, 0]), pinned(train[:, 1]), pinned(train[:, 2]), pinned(ent_host), pinned(rel_host),
                pinned(head_mean), pinned(tail_mean), pinned(np.empty_like(ent_host)), pinned(np.empty_like(rel_host))]
        p_h, p_t, p_r, p_ent, p_rel, p_hm, p_tm, o_ent, o_rel = (k[1] for k in keep)
        p_train = (p_h, p_t, p_r)
        e2e_ms = []
        with kb2e_b200.Context(CFG["model"], CFG["dim"], nE, nR, method=CFG["method"], distance=CFG["distance"],
                               batches=CFG["batches"], rate=CFG["rate"], margin=CFG["margin"], seed=101 + rank, device=local) as c2:
            for it in range(2 + K):
                barrier()
                t0 = time.perf_counter()
                c2.set_train_triples(p_train)           # H2D: triples (+ device hash-set build)
                c2.set_bern(p_hm, p_tm)                 # H2D: corruption probabilities
                c2.upload(TABLE_ENTITY, p_ent)          # H2D: fp64 tables
                c2.upload(TABLE_RELATION, p_rel)
                loss = c2.train_epochs(it, 1)           # D2H: epoch loss
                out_e = c2.download(TABLE_ENTITY, out=o_ent)    # D2H: fp64 tables
                out_r = c2.download(TABLE_RELATION, out=o_rel)
                dt = (time.perf_counter() - t0) * 1e3
                if it >= 2:
                    e2e_ms.append(max_over_ranks(dt))
        h2d = int(p_h.nbytes * 3 + p_ent.nbytes + p_rel.nbytes + p_hm.nbytes + p_tm.nbytes)
        d2h = int(out_e.nbytes + out_r.nbytes + 8)
        e2e_value = world * pairs_per_step * K / (sum(e2e_ms) * 1e-3)

        # ---- filtered ranking: test triples sharded over ranks ------------------------------------------
        # every rank ranks with rank 0's trained tables
        ent_t = torch.from_numpy(ctx.download(TABLE_ENTITY)).cuda()
        rel_t = torch.from_numpy(ctx.download(TABLE_RELATION)).cuda()
        if world > 1:
            dist.broadcast(ent_t, 0)
            dist.broadcast(rel_t, 0)
        ent_eval, rel_eval = np.round(ent_t.cpu().numpy(), 6), np.round(rel_t.cpu().numpy(), 6)  # what "%.6lf" files hold
        lo = n_test * rank // world
        hi = n_test * (rank + 1) // world
        ev = kb2e_b200.Context(CFG["model"], CFG["dim"], nE, nR, method=CFG["method"], distance=CFG["distance"], device=local)
        ev.upload(TABLE_ENTITY, ent_eval)
        ev.upload(TABLE_RELATION, rel_eval)
        ev.set_test_triples(test)
        ev.add_filter_triples(train)
        ev.add_filter_triples(valid)
        KE, WE = max(1, min(K, 3)), max(1, min(W, 2))
        for _ in range(WE):
            ev.rank(lo, hi - lo, want_ranks=False)
        r0 = ev.rank_stats()
        ev_ms = []
        for _ in range(KE):
            flush.fill_(1)
            barrier()
            before = ev.rank_stats()["kernel_ms"]
            res = ev.rank(lo, hi - lo, want_ranks=False)
            ev_ms.append(max_over_ranks(ev.rank_stats()["kernel_ms"] - before))
        r1 = ev.rank_stats()
        sums = sum_over_ranks(res["sums"])
        nq = 2 * n_test
        ev_value = nq * KE / (sum(ev_ms) * 1e-3)
        main_ms = (r1["main_kernel_ms"] - r0["main_kernel_ms"]) / KE
        # end to end: host tables + host triples in, per-query ranks out
        filt_all = np.concatenate([train, valid])
        keep2 = [pinned(test[:, k]) for k in range(3)] + [pinned(filt_all[:, k]) for k in range(3)] + \
                [pinned(ent_eval), pinned(rel_eval)] + [pinned(np.empty(2 * (hi - lo), dtype=np.int32)) for _ in range(4)]
        p_test = tuple(k[1] for k in keep2[0:3])
        p_filt = tuple(k[1] for k in keep2[3:6])
        p_ee, p_re = keep2[6][1], keep2[7][1]
        p_out = [k[1] for k in keep2[8:12]]
        ev_e2e = []
        with kb2e_b200.Context(CFG["model"], CFG["dim"], nE, nR, method=CFG["method"], distance=CFG["distance"], device=local) as e2:
            for it in range(1 + KE):
                barrier()
                t0 = time.perf_counter()
                e2.upload(TABLE_ENTITY, p_ee)          # H2D: fp64 tables
                e2.upload(TABLE_RELATION, p_re)
                e2.set_test_triples(p_test)            # H2D: test triples
                e2.clear_filter_triples()
                e2.add_filter_triples(p_filt)          # H2D: train + valid; the hashed CSR is rebuilt on the device
                full = e2.rank(lo, hi - lo, out=p_out)  # D2H: per-query raw / filtered ranks and tie counts
                dt = (time.perf_counter() - t0) * 1e3
                if it >= 1:
                    ev_e2e.append(max_over_ranks(dt))
            # resident: tables, test triples and the filter CSR stay on the device between calls (an evaluation loop
            # over checkpoints of one KG re-uploads only what changed); per step: kb2e_rank + ranks copied back
            ev_res = []
            for it in range(1 + KE):
                barrier()
                t0 = time.perf_counter()
                e2.rank(lo, hi - lo, out=p_out)
                dt = (time.perf_counter() - t0) * 1e3
                if it >= 1:
                    ev_res.append(max_over_ranks(dt))
        ev_e2e_value = nq * KE / (sum(ev_e2e) * 1e-3)
        ev_res_value = nq * KE / (sum(ev_res) * 1e-3)
        ev_h2d = int(ent_eval.nbytes + rel_eval.nbytes + 3 * p_test[0].nbytes + 3 * p_filt[0].nbytes)
        ev_d2h = int(4 * 4 * 2 * (hi - lo) + 32)
        ev_launches = int(r1["launches"] - r0["launches"])
        # ---- batched training: 8 models of the same configuration (different seeds) in one launch ----------------------
        sweep = None
        try:
            sweep = run_sweep(local, rank, train, head_mean, tail_mean, nE, nR, hbm_peak, peak_src, flush, K, W)
        except Exception as exc:
            sweep = {"error": repr(exc)}
        # ---- entity-partitioned training at the scaled shape (BASELINE configs[4]), N > 1 only ---------------
        part = None
        if not args.no_partitioned:
            try:
                if world > 1:
                    part = run_partitioned(rank, world, local, max_over_ranks, sum_over_ranks)
                else:
                    part = run_scaled_single(local, hbm_peak, peak_src)
            except Exception as exc:  # the headline line must survive a failure here
                part = {"error": repr(exc)}
    clk = clocks.summary()

    if rank != 0:
        ctx.close(); ev.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (rank 0, N=1 only): the reference itself on this box's host cores --------------
    cpu_train = cpu_eval = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_train, cpu_eval = reference_sample(train, valid, test, nE, nR, ent_eval, rel_eval, train_epochs=5, eval_triples=800)  # ~10 s + ~10 s of CPU work

    ranking_flops = 2.0 * (2 * (hi - lo)) * nE * CFG["dim"]  # this rank's shard: the roofline is per GPU
    traffic = profile_traffic("train")
    line = {
        "metric": "train_triples_per_s", "value": value, "unit": "triples/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "kb2e_b200",
        "config": bench_config(world),
        "wall_ms_per_step": sum(wall_ms) / K,
        "gpu_launches": launches + ev_launches,
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "triples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": sum(e2e_ms) / K,
                "what": "kb2e_set_train_triples + kb2e_set_bern + kb2e_upload x2 + kb2e_train_epochs(1) + kb2e_download x2, pinned host buffers"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "kb2e::train_kernel<TransE, 16 lanes x 2 float4 per pair, 640 threads, touched-row lists>",
                     "alpha": alpha, "touched_rows_per_batch": touched / (K * CFG["batches"]),
                     "algorithmic_bytes_per_launch": abytes / max(launches, 1),
                     "note": "algorithmic bytes (SURVEY 8d) / CUDA-event time of the persistent launch; tables are L2-resident, so DRAM traffic is far below the algorithmic bytes"},
        "cpu_baseline": cpu_train,
        "sweep": sweep,
        "partitioned": part,
        "eval": {
            "metric": "eval_queries_per_s", "value": ev_value, "unit": "queries/s", "ms_per_step": sum(ev_ms) / KE, "steps": KE,
            "queries_per_step": nq, "candidates": nE, "scaling": "strong",
            "raw_mean_rank": float(sums[0]) / nq, "filtered_mean_rank": float(sums[1]) / nq,
            "raw_hits10": float(sums[2]) / nq, "filtered_hits10": float(sums[3]) / nq,
            "e2e": {"value": ev_e2e_value, "unit": "queries/s", "h2d_bytes_per_step": ev_h2d, "d2h_bytes_per_step": ev_d2h,
                    "ms_per_step": sum(ev_e2e) / KE,
                    "what": "COLD: per step, on one long-lived context: kb2e_upload x2 (fp64 tables) + kb2e_set_test_triples + kb2e_add_filter_triples "
                            "(train + valid; filter CSR rebuilt on the device) + kb2e_rank with per-query ranks copied back; pinned host buffers",
                    "resident": {"value": ev_res_value, "unit": "queries/s", "ms_per_step": sum(ev_res) / KE, "d2h_bytes_per_step": ev_d2h,
                                 "what": "tables, test triples and filter CSR already on the device: kb2e_rank + per-query ranks copied back"}},
            "roofline": {"bound": "tensor", "achieved": ranking_flops / (main_ms * 1e-3) / 1e12, "peak": tensor_peak,
                         "unit": "TFLOP/s", "frac": ranking_flops / (main_ms * 1e-3) / 1e12 / tensor_peak, "traffic": profile_traffic("rank"),
                         "peak_source": peak_src + ", bf16 burst", "main_kernel_ms": main_ms,
                         "kernel": "kb2e::tc::rank_l2_tc_kernel (tcgen05.mma kind::f16, bf16 x3 split, fused compare-and-count epilogue)",
                         "executed_tflops": ranking_flops * 3 * 112 / CFG["dim"] / (main_ms * 1e-3) / 1e12,
                         "rechecked_pairs_per_step": (r1["rechecked"] - r0["rechecked"]) / KE,
                         "note": "achieved = algorithmic 2*Q*N*D FLOPs / CUDA-event time of the tensor-core kernel; the kernel executes 3 bf16 products per term "
                                 "(hi*hi + hi*lo + lo*hi) over K padded 100 -> 112 so that only a handful of candidates per query need the exact fp64 recheck"},
            "cpu_baseline": cpu_eval,
        },
    }
    print(json.dumps(line))
    ctx.close(); ev.close()
    if world > 1:
        dist.destroy_process_group()


def run_sweep(local, rank, train, head_mean, tail_mean, nE, nR, hbm_peak, peak_src, flush, K, W, n_models=8):
    """n_models independent TransE models of the headline configuration (seeds seed .. seed + n_models - 1)
    trained in ONE persistent launch per epoch (kb2e_set_replicas).  One model leaves the B200 latency-bound (a batch is a chain
    of L2 round trips and two grid barriers, profiles/README.md); stacked models share launch, phases and barriers, each
    computing exactly what it computes alone (tests/test_gpu_train.py).  value = pairs of ALL models per second."""
    import torch

    import kb2e_b200
    with kb2e_b200.Context(CFG["model"], CFG["dim"], nE, nR, method=CFG["method"], distance=CFG["distance"], batches=CFG["batches"],
                           rate=CFG["rate"], margin=CFG["margin"], seed=1 + rank, device=local) as ctx:
        ctx.set_replicas(n_models)
        ctx.set_train_triples(train)
        ctx.set_bern(head_mean, tail_mean)
        ctx.init_embeddings()
        for e in range(W):
            ctx.train_epochs(e, 1)
        s0 = ctx.train_stats()
        for e in range(W, W + K):
            flush.fill_(1)
            torch.cuda.synchronize()
            loss = ctx.train_epochs(e, 1)
        s1 = ctx.train_stats()
    ms = s1["kernel_ms"] - s0["kernel_ms"]
    n = s1["samples"] - s0["samples"]
    touched = (s1["touched_ent"] - s0["touched_ent"]) + (s1["touched_rel"] - s0["touched_rel"])
    abytes, alpha = algorithmic_bytes(n, s1["active"] - s0["active"], touched, CFG["dim"])
    achieved = abytes / (ms * 1e-3) / 1e9
    return {"metric": "train_triples_per_s", "value": n / (ms * 1e-3), "unit": "triples/s", "models": n_models, "steps": K,
            "ms_per_step": ms / K, "alpha": alpha,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                         "peak_source": peak_src, "kernel": "kb2e::train_sweep_kernel<16 lanes x 2 float4, 640 threads>",
                         "note": "aggregate algorithmic bytes of the stacked models (SURVEY 8d) / CUDA-event time of the launch; "
                                 "8 x 13 MB of tables stay L2-resident"},
            "config": {"workload": WORKLOAD + " -- %d models, seeds %d..%d, one launch per epoch" % (n_models, 1 + rank, rank + n_models),
                       "step": "one epoch of every model = %d x 100 batches x 4831 pairs" % n_models},
            "final_loss_per_model": [float(x) for x in np.atleast_2d(loss)[:, -1]]}


def run_scaled_single(local, hbm_peak, peak_src):
    """BASELINE configs[4] on ONE GPU (the whole 3.2 GB table in one HBM): the same persistent kernel as the headline,
    at the one shape where the training step is genuinely HBM-bound.  It is the N = 1 point of the `partitioned` strong-
    scaling series and carries its own roofline fraction, measured live (CUDA events, algorithmic bytes of SURVEY 8d)."""
    import kb2e_b200
    from kb2e_b200 import kg
    nE, nR, ntr, _, _, _ = kg.SHAPES["scaled"]
    dim = 200
    rng = np.random.default_rng(1)
    train = (rng.integers(0, nE, ntr, dtype=np.int32), rng.integers(0, nE, ntr, dtype=np.int32), rng.integers(0, nR, ntr, dtype=np.int32))
    with kb2e_b200.Context("transe", dim, nE, nR, method=0, distance=1, batches=100, rate=0.01, margin=1.0, seed=1, device=local) as ctx:
        ctx.set_train_triples(train)
        del train
        ctx.set_bern(None, None)
        ctx.init_embeddings()
        ctx.train_epochs(0, 1)       # warm-up epoch
        s0 = ctx.train_stats()
        loss = ctx.train_epochs(1, 1)
        s1 = ctx.train_stats()
    ms = s1["kernel_ms"] - s0["kernel_ms"]
    n = s1["samples"] - s0["samples"]
    touched = (s1["touched_ent"] - s0["touched_ent"]) + (s1["touched_rel"] - s0["touched_rel"])
    abytes, alpha = algorithmic_bytes(n, s1["active"] - s0["active"], touched, dim)
    achieved = abytes / (ms * 1e-3) / 1e9
    return {"metric": "train_triples_per_s", "value": n / (ms * 1e-3), "unit": "triples/s", "n_gpus": 1, "scaling": "strong",
            "ms_per_epoch": ms, "alpha": alpha,
            "series_note": "N = 1 point of the partitioned series: the single-GPU kernel (kb2e_train_epochs), same triples, seed and "
                           "initial tables as the N > 1 points (kb2e_dist_init_embeddings seeds rows exactly like kb2e_init_embeddings)",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": profile_traffic("scaled"), "peak_source": peak_src, "kernel": "kb2e::train_kernel<TransE,32,2,768>",
                         "note": "traffic = dram bytes of one 100-batch launch from the committed ncu capture (profiles/r01_ncu_full_summary.txt)"},
            "config": {"workload": "TransE L2 size=200, 4,000,000 entities x 1,345 relations, 100,000,000 random triples, batches=100 "
                                   "(1,000,000 pairs per batch), single GPU: the N = 1 point of the partitioned series",
                       "timing": "CUDA events around the persistent launch of one epoch; 1 warm-up epoch"},
            "loss": float(loss[0])}


def run_partitioned(rank, world, local, max_over_ranks, sum_over_ranks):
    """BASELINE configs[4]: TransE squared-L2 size=200 on 4 M entities / 100 M (random) triples, entity rows
    partitioned over the ranks, relation rows replicated (kb2e_b200/csrc/train_dist.cu).  One timed epoch =
    100 batches x 1,000,000 pairs for the whole job (strong scaling: total work fixed)."""
    from kb2e_b200 import kg
    from kb2e_b200.partitioned import PartitionedTrainer
    nE, nR, ntr, _, _, _ = kg.SHAPES["scaled"]
    dim = 200
    rng = np.random.default_rng(1)   # same seed on every rank: identical synthetic triples (throughput only)
    train = (rng.integers(0, nE, ntr, dtype=np.int32), rng.integers(0, nE, ntr, dtype=np.int32), rng.integers(0, nR, ntr, dtype=np.int32))
    pt = PartitionedTrainer(dim, nE, nR, rank, world, local, method=0, distance=1, batches=100, rate=0.01, margin=1.0, seed=1)
    try:
        pt.set_training_set(train, None, None)
        del train
        pt.init_embeddings()
        pt.train_epochs(0, 1)        # warm-up epoch
        s0 = pt.ctx.train_stats()
        loss = pt.train_epochs(1, 1)
        s1 = pt.ctx.train_stats()
        ms = max_over_ranks(s1["kernel_ms"] - s0["kernel_ms"])
        cnt = sum_over_ranks([s1["samples"] - s0["samples"], s1["active"] - s0["active"],
                              s1["touched_ent"] - s0["touched_ent"] + s1["touched_rel"] - s0["touched_rel"]])
        n, active, touched = (float(x) for x in cnt)
        abytes, alpha = algorithmic_bytes(n, active, touched, dim)
        hbm_peak, _, peak_src = peaks()
        per_gpu = abytes / world / (ms * 1e-3) / 1e9
        # bytes that must cross NVLink per GPU and epoch: a kept sample has its head local; tail and corrupting rows are remote
        # with probability (world - 1) / world each: one row served in, and (if the hinge is active) one update pushed out
        P4 = (dim + 3) // 4 * 4 * 4
        nvlink_bytes = n / world * 2.0 * (world - 1) / world * (P4 + 8 + alpha * (P4 + 16))
        return {"metric": "train_triples_per_s", "value": n / (ms * 1e-3), "unit": "triples/s", "n_gpus": world, "scaling": "strong",
                "ms_per_epoch": ms, "alpha": alpha, "algorithmic_GBps_total": abytes / (ms * 1e-3) / 1e9,
                "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": hbm_peak, "unit": "GB/s", "frac": per_gpu / hbm_peak, "traffic": None,
                             "peak_source": peak_src, "kernel": "kb2e::train_dist_kernel",
                             "nvlink_GBps_per_gpu_each_way": nvlink_bytes / (ms * 1e-3) / 1e9, "nvlink_peak_GBps": 640.0,
                             "nvlink_frac": nvlink_bytes / (ms * 1e-3) / 1e9 / 640.0,
                             "note": "per-GPU algorithmic bytes (SURVEY 8d) / epoch time against the measured HBM peak; NVLink: rows served in + "
                                     "updates pushed out per GPU against the 640 GB/s posted-write rate of profiles/r01_p2p_bench.txt"},
                "config": {"workload": "TransE L2 size=200, 4,000,000 entities x 1,345 relations, 100,000,000 random triples, batches=100 "
                                       "(1,000,000 pairs per batch), entity rows partitioned by id mod N, relation rows replicated",
                           "exchange": "row requests, rows and updates as posted peer stores / vector REDs over NVLink (CUDA IPC peer memory); no NCCL on the data path",
                           "timing": "CUDA events around the persistent launch of one epoch, max over ranks; 1 warm-up epoch"},
                "loss": float(loss[0])}
    finally:
        pt.close()


def profile_traffic(which):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one matches."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(which)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------ reference arm
def reference_sample(train, valid, test, nE, nR, ent_eval, rel_eval, train_epochs, eval_triples, steps=1, warmup=0):
    """Times the unmodified reference (oracle/_ref) on this box: `train_epochs` epochs of bfgs() per step and
    evalCorruption over the first `eval_triples` test triples.  Returns (train block, eval block)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from kb2e_oracle import Reference, ReferenceTrainer
    from kb2e_b200 import kg

    if not Reference.available():
        return ({"unavailable": "oracle/_ref/libkb2e_ref.so not built"}, None)
    ref = Reference()
    tmp = tempfile.mkdtemp(prefix="kb2e_bench_")
    try:
        kg.write_kg({"nE": nE, "nR": nR, "train": train, "valid": valid, "test": test}, tmp)
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)  # the reference prints its per-epoch line; keep the JSON line clean
        try:
            tr = ReferenceTrainer(ref, 0, tmp, tmp, CFG["dim"], CFG["rate"], CFG["margin"], CFG["method"], CFG["distance"],
                                  CFG["batches"], 1)
            secs = []
            for it in range(warmup + steps):
                s = tr.epochs(train_epochs)
                if it >= warmup:
                    secs.append(s)
            tr.close()
        finally:
            os.dup2(saved, 1)
            os.close(devnull)
        pairs = train_epochs * CFG["batches"] * (len(train) // CFG["batches"])
        train_block = {"value": pairs * len(secs) / sum(secs), "unit": "triples/s", "cores": 1, "kind": "reference",
                       "sample": f"{train_epochs} epoch(s) of the reference's own bfgs() per step ({pairs} pairs), load/init/write excluded; single-threaded program",
                       "seconds_per_step": sum(secs) / len(secs), "step_seconds": secs}
        t0 = time.perf_counter()
        ref.rank(0, CFG["distance"], ent_eval, rel_eval, None, test[:eval_triples], np.concatenate([train, valid]))
        dt = time.perf_counter() - t0
        eval_block = {"value": 2 * eval_triples / dt, "unit": "queries/s", "cores": 1, "kind": "reference",
                      "sample": f"evalCorruption on the first {eval_triples} test triples ({2 * eval_triples} queries x {nE} candidates), "
                                "including building its filter map; the N_R x N_E^2 cache reset of run() is NOT included (it would add about 0.5 s per relation)",
                      "seconds": dt}
        return train_block, eval_block
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return  # the reference is a single-process CPU program: rank 0 alone runs it
    from kb2e_b200 import kg
    nE, nR, n_train, n_valid, n_test, _ = kg.SHAPES[CFG["shape"]]
    g = kg.make_kg(CFG["shape"], seed=0)
    rng = np.random.default_rng(0)
    # the ranking sample needs tables; random N(0, 1/D) rows rounded like the "%.6lf" files are enough for timing
    ent = np.round(rng.normal(0, 1.0 / CFG["dim"], (nE, CFG["dim"])), 6)
    rel = np.round(rng.normal(0, 1.0 / CFG["dim"], (nR, CFG["dim"])), 6)
    tb, eb = reference_sample(g["train"], g["valid"], g["test"], nE, nR, ent, rel, train_epochs=1, eval_triples=100,
                              steps=args.steps, warmup=args.warmup)
    if "unavailable" in tb:
        print(json.dumps({"impl": "reference", "unavailable": tb["unavailable"]}))
        return
    line = {
        "metric": "train_triples_per_s", "value": tb["value"], "unit": "triples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tb["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": bench_config(world),
        "cpu_baseline": {k: tb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": tb["value"], "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "eval": {"metric": "eval_queries_per_s", "value": eb["value"], "unit": "queries/s", "cpu_baseline": eb,
                 "e2e": {"value": eb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kb2e_b200", choices=["kb2e_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-partitioned", action="store_true", help="skip the scaled-shape (BASELINE configs[4]) run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3  # timing rule: at least 3 warm-up steps
        run_product(args)


if __name__ == "__main__":
    main()
