"""Trained-model (statistical) parity: models trained by the reference vs by the B200 path on identical
synthetic data, 3 seeds (north_star: filtered MeanRank within 2 %, Hits@10 within 0.5 points, mean of 3).

  stage ref  (build container, CPU, needs oracle/_ref):  trains TransE / TransH / TransR with the UNMODIFIED
             reference (TransR: zero-patched energies, seeded from a reference TransE-unif run) and stores the
             KGs, the trained tables and their metrics in tests/golden/stat_parity.npz
  stage gpu  (GPU box): trains the same configurations with kb2e_b200 from the same KGs, ranks both sides'
             tables with kb2e_rank, prints and stores the comparison (profiles/stat_parity_r01.json)
"""
import argparse
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

SHAPE, D, LR, MARGIN, BATCHES = "small", 20, 0.01, 1.0, 100
EPOCHS = {"transe": 300, "transe_unif": 300, "transh": 300, "transr": 100}
SEEDS = (0, 1, 2)
GOLD = os.path.join(ROOT, "tests", "golden", "stat_parity.npz")


def metrics(oracle, model, dist, ent, rel, w, g):
    lo, hi, flo, fhi = oracle.rank(model, dist, ent, rel, w, g["test"], np.concatenate([g["train"], g["valid"]]))
    n = len(lo)
    return {"raw_mr": float(lo.mean()), "filt_mr": float(flo.mean()), "raw_h10": float((lo <= 10).mean()),
            "filt_h10": float((flo <= 10).mean()), "n": n}


def stage_ref():
    from kb2e_oracle import Oracle, Reference
    ref, oracle = Reference(), Oracle()
    out = {}
    report = {}
    for s in SEEDS:
        g = kg.make_kg(SHAPE, seed=s)
        nE, nR = g["nE"], g["nR"]
        for k in ("train", "valid", "test"):
            out[f"s{s}_{k}"] = g[k]
        tmp = tempfile.mkdtemp(prefix=f"kb2e_sp{s}_")
        kg.write_kg(g, tmp)
        for name, model, method, dist in (("transe", 0, 1, 0), ("transe_unif", 0, 0, 0), ("transh", 1, 1, 0), ("transr", 2, 1, 0)):
            od = os.path.join(tmp, name)
            os.makedirs(od, exist_ok=True)
            t0 = time.time()
            secs = ref.train_files(model, tmp, od, D, LR, MARGIN, method, dist, BATCHES, EPOCHS[name], s + 1,
                                   seeddir=os.path.join(tmp, "transe_unif"), seedmethod=0, zero_work=True, write=True)
            sfx = "unif" if method == 0 else "bern"
            ent = kg.read_embeddings(os.path.join(od, "entity2vec." + sfx), nE, D)
            rel = kg.read_embeddings(os.path.join(od, "relation2vec." + sfx), nR, D)
            w = None
            if model == 1:
                w = kg.read_embeddings(os.path.join(od, "weights." + sfx), nR, D)
            if model == 2:
                w = kg.read_embeddings(os.path.join(od, "weights." + sfx), nR * D, D).reshape(nR, D, D)
            m = metrics(oracle, model, dist, ent, rel, w, g)
            m["bfgs_seconds"] = secs
            report[f"s{s}_{name}"] = m
            out[f"s{s}_{name}_ent"], out[f"s{s}_{name}_rel"] = ent.astype(np.float32), rel.astype(np.float32)
            if w is not None:
                out[f"s{s}_{name}_w"] = np.asarray(w, dtype=np.float32)
            print(f"seed {s} {name}: {m}  ({time.time() - t0:.0f}s)", file=sys.stderr, flush=True)
        # (B) the reference's update rule (bitwise-pinned restatement, sequential per-update renormalisation)
        # driven by the UNIFORM counter sampler: isolates the effect of the reference's non-uniform randMax
        # (common/utils.cpp:113-120: an int-overflowing product of two rand() values, e.g. 75 % even indices).
        smp_cache = {}
        btab = {}
        for name, model, method, dist in (("transe", 0, 1, 0), ("transe_unif", 0, 0, 0), ("transh", 1, 1, 0), ("transr", 2, 1, 0)):
            t0 = time.time()
            rng = np.random.default_rng(500 + s)
            ent = rng.normal(0, 1.0 / D, (nE, D))
            rel = rng.normal(0, 1.0 / D, (nR, D))
            w = None
            if model == 1:
                w = rng.normal(0, 1.0 / D, (nR, D))
                w /= np.linalg.norm(w, axis=1, keepdims=True)
            if model == 2:
                e0, r0 = btab["transe_unif"]
                ent = np.round(e0, 6)
                ent = ent / np.linalg.norm(ent, axis=1, keepdims=True)
                rel = np.round(r0, 6).copy()
                w = np.ascontiguousarray(np.tile(np.eye(D), (nR, 1, 1)))
            if method not in smp_cache:
                smp_cache[method] = oracle.sampler(g["train"], nE, nR, method)
            smp_cache[method].train_epochs_ref(model, dist, LR, MARGIN, BATCHES, 0, EPOCHS[name], 2000 + s, ent, rel, w)
            btab[name] = (ent, rel)
            m = metrics(oracle, model, dist, np.round(ent, 6), np.round(rel, 6), None if w is None else np.round(w, 6), g)
            report[f"s{s}_{name}_uniform"] = m
            print(f"seed {s} {name} [reference update rule + uniform sampler]: {m}  ({time.time() - t0:.0f}s)", file=sys.stderr, flush=True)
    out["report"] = np.frombuffer(json.dumps(report).encode(), dtype=np.uint8)
    np.savez_compressed(GOLD, **out)
    print("wrote", GOLD, os.path.getsize(GOLD))


def stage_gpu(out_path):
    import kb2e_b200
    from kb2e_b200 import TABLE_ENTITY, TABLE_RELATION, TABLE_WEIGHTS
    gold = np.load(GOLD)
    ref_report = json.loads(bytes(gold["report"]).decode())
    rows = []

    def rank_tables(model, dist, nE, nR, ent, rel, w, g):
        with kb2e_b200.Context(model, D, nE, nR, distance=dist) as ev:
            ev.upload(TABLE_ENTITY, ent)
            ev.upload(TABLE_RELATION, rel)
            if w is not None:
                ev.upload(TABLE_WEIGHTS, np.asarray(w, dtype=np.float64).reshape(ev.table_shape(TABLE_WEIGHTS)))
            ev.set_test_triples(g["test"])
            ev.add_filter_triples(g["train"])
            ev.add_filter_triples(g["valid"])
            s = ev.rank(want_ranks=False)["sums"]
            n = 2 * len(g["test"])
            return {"raw_mr": s[0] / n, "filt_mr": s[1] / n, "raw_h10": s[2] / n, "filt_h10": s[3] / n}

    for s in SEEDS:
        g = {k: gold[f"s{s}_{k}"] for k in ("train", "valid", "test")}
        nE, nR = kg.SHAPES[SHAPE][0], kg.SHAPES[SHAPE][1]
        hm, tm = kg.bern_stats(g["train"], nR)
        trained = {}
        for name, model, method, dist in (("transe", 0, 1, 0), ("transe_unif", 0, 0, 0), ("transh", 1, 1, 0), ("transr", 2, 1, 0)):
            with kb2e_b200.Context(model, D, nE, nR, method=method, distance=dist, batches=BATCHES, rate=LR, margin=MARGIN,
                                   seed=1000 + s) as ctx:
                ctx.set_train_triples(g["train"])
                ctx.set_bern(hm, tm)
                ctx.init_embeddings()
                if model == 2:
                    # transr/trainer.cpp:88-113: seed from the TransE-unif run (6-decimal text), entity rows unit length
                    e0 = np.round(trained["transe_unif"][0], 6)
                    e0 = e0 / np.linalg.norm(e0, axis=1, keepdims=True)
                    ctx.upload(TABLE_ENTITY, e0)
                    ctx.upload(TABLE_RELATION, np.round(trained["transe_unif"][1], 6))
                t0 = time.time()
                loss = ctx.train_epochs(0, EPOCHS[name])
                secs = time.time() - t0
                ent, rel = ctx.download(TABLE_ENTITY), ctx.download(TABLE_RELATION)
                w = ctx.download(TABLE_WEIGHTS) if model != 0 else None
            trained[name] = (ent, rel, w)
            # both sides ranked by the same exact kernel from 6-decimal values (what the eval programs read)
            mine = rank_tables(model, dist, nE, nR, np.round(ent, 6), np.round(rel, 6), None if w is None else np.round(w, 6), g)
            # the fixture keeps the reference's 6-decimal file values as float32; round(x * 1e6) / 1e6 recovers the
            # exact doubles the reference's fscanf("%lf") produced (float32 error * 1e6 < 0.1)
            exact = lambda a: np.round(np.asarray(a, dtype=np.float64) * 1e6) / 1e6
            rw = exact(gold[f"s{s}_{name}_w"]) if model != 0 else None
            theirs = rank_tables(model, dist, nE, nR, exact(gold[f"s{s}_{name}_ent"]), exact(gold[f"s{s}_{name}_rel"]), rw, g)
            rr = ref_report[f"s{s}_{name}"]
            ru = ref_report.get(f"s{s}_{name}_uniform")
            rows.append({"seed": s, "model": name, "gpu": mine, "reference": theirs, "reference_cpu_eval": rr,
                         "reference_rule_uniform_sampler": ru,
                         "gpu_train_seconds": secs, "reference_bfgs_seconds": rr["bfgs_seconds"], "final_loss": float(loss[-1])})
            print(f"seed {s} {name:12s} gpu filt MR {mine['filt_mr']:8.2f} H@10 {mine['filt_h10']:.4f} | ref filt MR {theirs['filt_mr']:8.2f} "
                  f"H@10 {theirs['filt_h10']:.4f} (ref CPU eval {rr['filt_mr']:.2f}) | train {secs:.2f}s vs {rr['bfgs_seconds']:.1f}s", flush=True)
    summary = {}
    for name in ("transe", "transe_unif", "transh", "transr"):
        sel = [r for r in rows if r["model"] == name]
        gm = np.mean([r["gpu"]["filt_mr"] for r in sel]); rm = np.mean([r["reference"]["filt_mr"] for r in sel])
        gh = np.mean([r["gpu"]["filt_h10"] for r in sel]); rh = np.mean([r["reference"]["filt_h10"] for r in sel])
        spread = np.std([r["reference"]["filt_mr"] for r in sel]) / rm
        um = np.mean([r["reference_rule_uniform_sampler"]["filt_mr"] for r in sel])
        uh = np.mean([r["reference_rule_uniform_sampler"]["filt_h10"] for r in sel])
        summary[name] = {"gpu_filt_mr": gm, "ref_filt_mr": rm, "mr_rel_diff": (gm - rm) / rm, "gpu_filt_h10": gh, "ref_filt_h10": rh,
                         "h10_diff_points": 100 * (gh - rh), "ref_seed_spread_rel": spread,
                         "uniform_filt_mr": um, "uniform_filt_h10": uh, "mr_rel_diff_vs_uniform": (gm - um) / um,
                         "h10_diff_points_vs_uniform": 100 * (gh - uh)}
        print(f"{name:12s} mean filt MR: gpu {gm:.2f} | reference as shipped {rm:.2f} ({100 * (gm - rm) / rm:+.2f}%; its seed spread {100 * spread:.1f}%) | "
              f"reference rule + uniform sampler {um:.2f} ({100 * (gm - um) / um:+.2f}%)   H@10: gpu {gh:.4f} | shipped {rh:.4f} ({100 * (gh - rh):+.2f} pt) | "
              f"uniform {uh:.4f} ({100 * (gh - uh):+.2f} pt)")
    if out_path:
        json.dump({"config": {"shape": SHAPE, "D": D, "lr": LR, "margin": MARGIN, "batches": BATCHES, "epochs": EPOCHS, "seeds": SEEDS},
                   "rows": rows, "summary": summary}, open(out_path, "w"), indent=1, default=float)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", required=True, choices=["ref", "gpu"])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    stage_ref() if a.stage == "ref" else stage_gpu(a.out)
