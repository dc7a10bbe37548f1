"""Trained-model parity at a BASELINE shape (north_star: filtered MeanRank within 2 %, Hits@10 within 0.5 points,
mean over 3 seeds): BASELINE configs[1] -- TransE bern squared-L2 size=100 rate=0.01 margin=1 batches=100 on the
synthetic FB15k-shape KG (14,951 / 1,345 / 483,142 / 50,000 / 59,071), EPOCHS epochs on both sides, all 118,142
filtered-ranking queries.

  stage ref  (build container, CPU; needs oracle/_ref): for every data seed trains
       shipped   the UNMODIFIED reference (its own bfgs(), std::rand, randMax)                       [the parity target]
       uniform   the reference's update rule (bitwise-pinned restatement) + the uniform counter sampler
       randmax   the reference's update rule + the counter sampler with randMax's index DISTRIBUTION
       dfr       the deferred-renormalisation rule the CUDA kernels implement (fp64 CPU twin) + uniform sampler
     and ranks every model with the oracle (exact fp64, the reference's tie rule).  One (seed, arm) per process:
         python tools/stat_parity_large.py --stage ref --seed 0 --arm shipped --out /tmp/piece.json
     `--stage merge --pieces DIR` collects the pieces into tests/golden/stat_parity_fb15k.json (metrics only: the
     KGs are regenerated from their seeds, the reference's trained tables stay out of the repo).
  stage gpu  (GPU box): trains the same configuration with kb2e_b200 (default uniform sampler, and with
     KB2E_FLAG_SAMPLER_RANDMAX), ranks with kb2e_rank, compares the 3-seed means with the stored metrics.
"""
import argparse
import glob
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from kb2e_b200 import kg  # noqa: E402

SHAPE, D, LR, MARGIN, BATCHES, METHOD, DIST = "fb15k", 100, 0.01, 1.0, 100, 1, 1
EPOCHS = 100
SEEDS = (0, 1, 2)
ARMS = ("shipped", "uniform", "randmax", "dfr")
GOLD = os.path.join(ROOT, "tests", "golden", "stat_parity_fb15k.json")


def metrics_from_ranks(raw, filt):
    raw, filt = np.asarray(raw, dtype=np.int64), np.asarray(filt, dtype=np.int64)
    return {"raw_mr": float(raw.mean()), "filt_mr": float(filt.mean()), "raw_h10": float((raw <= 10).mean()),
            "filt_h10": float((filt <= 10).mean()), "n": int(len(raw))}


def stage_ref(seed, arm, out_path, epochs):
    from kb2e_oracle import Oracle, Reference
    oracle = Oracle()
    g = kg.make_kg(SHAPE, seed=seed)
    nE, nR = g["nE"], g["nR"]
    t0 = time.time()
    if arm == "shipped":
        ref = Reference()
        tmp = tempfile.mkdtemp(prefix=f"kb2e_spl{seed}_")
        kg.write_kg(g, tmp)
        secs = ref.train_files(0, tmp, tmp, D, LR, MARGIN, METHOD, DIST, BATCHES, epochs, seed + 1, write=True)
        ent = kg.read_embeddings(os.path.join(tmp, "entity2vec.bern"), nE, D)
        rel = kg.read_embeddings(os.path.join(tmp, "relation2vec.bern"), nR, D)
    else:
        rng = np.random.default_rng(500 + seed)
        ent = rng.normal(0, 1.0 / D, (nE, D))
        rel = rng.normal(0, 1.0 / D, (nR, D))
        smp = oracle.sampler(g["train"], nE, nR, METHOD)
        if arm == "randmax":
            smp.set_mode(1)
        t1 = time.time()
        if arm == "dfr":
            smp.train_epochs_dfr(0, DIST, LR, MARGIN, BATCHES, 0, epochs, 2000 + seed, ent, rel, None)
        else:
            smp.train_epochs_ref(0, DIST, LR, MARGIN, BATCHES, 0, epochs, 2000 + seed, ent, rel, None)
        secs = time.time() - t1
        ent, rel = np.round(ent, 6), np.round(rel, 6)   # what the "%.6lf" files hold
    t_train = time.time() - t0
    lo, hi, flo, fhi = oracle.rank(0, DIST, ent, rel, None, g["test"], np.concatenate([g["train"], g["valid"]]))
    m = metrics_from_ranks(lo, flo)
    m.update({"seed": seed, "arm": arm, "epochs": epochs, "train_seconds": secs, "total_seconds": time.time() - t0,
              "ties": int((np.asarray(hi) != np.asarray(lo)).sum())})
    json.dump(m, open(out_path, "w"))
    print(m, f"(train {t_train:.0f}s)", flush=True)


def stage_merge(pieces_dir):
    rows = [json.load(open(p)) for p in sorted(glob.glob(os.path.join(pieces_dir, "*.json")))]
    out = {"config": {"shape": SHAPE, "D": D, "lr": LR, "margin": MARGIN, "batches": BATCHES, "method": METHOD, "distance": DIST,
                      "epochs": rows[0]["epochs"], "seeds": sorted({r["seed"] for r in rows}),
                      "note": "metrics of models trained on CPU by the reference / the oracle on kg.make_kg('fb15k', seed); "
                              "ranked with oracle.rank (exact fp64, bitwise-pinned to the reference)"},
           "rows": rows}
    json.dump(out, open(GOLD, "w"), indent=1)
    print("wrote", GOLD, len(rows), "rows")


def summarize(rows, arm):
    sel = [r for r in rows if r["arm"] == arm]
    return (float(np.mean([r["filt_mr"] for r in sel])), float(np.mean([r["filt_h10"] for r in sel])),
            float(np.std([r["filt_mr"] for r in sel])), len(sel))


def stage_gpu(out_path, seeds=None):
    import kb2e_b200
    from kb2e_b200 import TABLE_ENTITY, TABLE_RELATION
    gold = json.load(open(GOLD))
    epochs = gold["config"]["epochs"]
    seeds = list(seeds if seeds is not None else gold["config"]["seeds"])
    nE, nR = kg.SHAPES[SHAPE][0], kg.SHAPES[SHAPE][1]
    rows = []
    for s in seeds:
        g = kg.make_kg(SHAPE, seed=s)
        hm, tm = kg.bern_stats(g["train"], nR)
        for arm, flags in (("gpu_uniform", 0), ("gpu_randmax", kb2e_b200.FLAG_SAMPLER_RANDMAX)):
            with kb2e_b200.Context(0, D, nE, nR, method=METHOD, distance=DIST, batches=BATCHES, rate=LR, margin=MARGIN,
                                   seed=1000 + s, flags=flags) as ctx:
                ctx.set_train_triples(g["train"])
                ctx.set_bern(hm, tm)
                ctx.init_embeddings()
                t0 = time.time()
                loss = ctx.train_epochs(0, epochs)
                secs = time.time() - t0
                ent, rel = ctx.download(TABLE_ENTITY), ctx.download(TABLE_RELATION)
            with kb2e_b200.Context(0, D, nE, nR, distance=DIST) as ev:
                ev.upload(TABLE_ENTITY, np.round(ent, 6))
                ev.upload(TABLE_RELATION, np.round(rel, 6))
                ev.set_test_triples(g["test"])
                ev.add_filter_triples(g["train"])
                ev.add_filter_triples(g["valid"])
                sm = ev.rank(want_ranks=False)["sums"]
            n = 2 * len(g["test"])
            row = {"seed": s, "arm": arm, "epochs": epochs, "raw_mr": sm[0] / n, "filt_mr": sm[1] / n, "raw_h10": sm[2] / n,
                   "filt_h10": sm[3] / n, "n": n, "train_seconds": secs, "final_loss": float(loss[-1])}
            rows.append(row)
            print(row, flush=True)
    ref_rows = [r for r in gold["rows"] if r["seed"] in seeds]
    summary = {}
    for garm, targets in (("gpu_uniform", ("uniform", "dfr", "shipped")), ("gpu_randmax", ("shipped", "randmax"))):
        gm, gh, gsd, _ = summarize(rows, garm)
        for t in targets:
            rm, rh, rsd, nref = summarize(ref_rows, t)
            if nref == 0:
                continue
            summary[f"{garm}_vs_{t}"] = {"gpu_filt_mr": gm, "ref_filt_mr": rm, "mr_rel_diff": (gm - rm) / rm, "gpu_filt_h10": gh,
                                         "ref_filt_h10": rh, "h10_diff_points": 100 * (gh - rh), "ref_seed_std_rel": rsd / rm,
                                         "gpu_seed_std_rel": gsd / gm, "seeds": nref}
            print(f"{garm:12s} vs {t:8s}: filt MR {gm:8.3f} vs {rm:8.3f} ({100 * (gm - rm) / rm:+.2f} %)   "
                  f"H@10 {gh:.4f} vs {rh:.4f} ({100 * (gh - rh):+.2f} pt)", flush=True)
    res = {"config": gold["config"], "rows": rows, "reference_rows": ref_rows, "summary": summary}
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1, default=float)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", required=True, choices=["ref", "merge", "gpu"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--arm", default="shipped", choices=ARMS)
    ap.add_argument("--epochs", type=int, default=EPOCHS)
    ap.add_argument("--pieces", default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.stage == "ref":
        stage_ref(a.seed, a.arm, a.out, a.epochs)
    elif a.stage == "merge":
        stage_merge(a.pieces)
    else:
        stage_gpu(a.out)
