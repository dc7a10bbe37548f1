"""Many-seed CPU study of what separates the trained models of the B200 path from the reference's
(north_star: filtered MeanRank within 2 %, Hits@10 within 0.5 points): everything here runs on the host, in fp64,
with the oracle (bitwise-pinned restatement) and the compiled reference, so that each semantic difference can be
switched on alone and measured with a confidence interval instead of a 3-seed guess.

Arms, per data seed (2,000-entity "small" KG, D=20, rate 0.01, margin 1, 100 batches):
   shipped      the unmodified reference program (its bfgs(), std::rand, randMax)
   ref_randmax  the reference's update rule + counter sampler with randMax's index distribution
   ref_uniform  the reference's update rule + uniform counter sampler
   dfr          deferred renormalisation (what the CUDA kernels compute, fp64 twin) + uniform counter sampler
   dfr_nocarry  dfr without handing the constraint's perturbation of w_r / M_r to the next batch (TransH / TransR)

    python tools/stat_parity_cpu.py --seeds 12 --procs 6 --out profiles/r02_stat_parity_cpu.json
"""
import argparse
import json
import os
import sys
import tempfile
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from kb2e_b200 import kg  # noqa: E402

SHAPE, D, LR, MARGIN, BATCHES = "small", 20, 0.01, 1.0, 100
EPOCHS = {"transe": 300, "transe_unif": 300, "transh": 300, "transr": 100}
MODELS = (("transe", 0, 1, 0), ("transe_unif", 0, 0, 0), ("transh", 1, 1, 0), ("transr", 2, 1, 0))
ARMS = ("shipped", "ref_randmax", "ref_uniform", "dfr", "dfr_nocarry")


def metrics(oracle, model, dist, ent, rel, w, g):
    lo, hi, flo, fhi = oracle.rank(model, dist, ent, rel, w, g["test"], np.concatenate([g["train"], g["valid"]]))
    return {"filt_mr": float(flo.mean()), "filt_h10": float((flo <= 10).mean()), "raw_mr": float(lo.mean())}


def job(args):
    seed, arm = args
    from kb2e_oracle import Oracle, Reference
    oracle = Oracle()
    g = kg.make_kg(SHAPE, seed=seed)
    nE, nR = g["nE"], g["nR"]
    out = {}
    t0 = time.time()
    if arm == "shipped":
        ref = Reference()
        tmp = tempfile.mkdtemp(prefix=f"kb2e_spc{seed}_")
        kg.write_kg(g, tmp)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        for name, model, method, dist in MODELS:
            od = os.path.join(tmp, name)
            os.makedirs(od, exist_ok=True)
            ref.train_files(model, tmp, od, D, LR, MARGIN, method, dist, BATCHES, EPOCHS[name], seed + 1,
                            seeddir=os.path.join(tmp, "transe_unif"), seedmethod=0, zero_work=True, write=True)
            sfx = "unif" if method == 0 else "bern"
            ent = kg.read_embeddings(os.path.join(od, "entity2vec." + sfx), nE, D)
            rel = kg.read_embeddings(os.path.join(od, "relation2vec." + sfx), nR, D)
            w = None
            if model == 1:
                w = kg.read_embeddings(os.path.join(od, "weights." + sfx), nR, D)
            if model == 2:
                w = kg.read_embeddings(os.path.join(od, "weights." + sfx), nR * D, D).reshape(nR, D, D)
            out[name] = metrics(oracle, model, dist, ent, rel, w, g)
    else:
        oracle.lib.orc_set_dfr_no_carry(1 if arm == "dfr_nocarry" else 0)
        samplers = {}
        tabs = {}
        for name, model, method, dist in MODELS:
            if arm == "dfr_nocarry" and model == 0 and name == "transe":
                continue   # TransE has no carry; transe_unif is still needed to seed TransR
            rng = np.random.default_rng(500 + seed)
            ent = rng.normal(0, 1.0 / D, (nE, D))
            rel = rng.normal(0, 1.0 / D, (nR, D))
            w = None
            if model == 1:
                w = rng.normal(0, 1.0 / D, (nR, D))
                w /= np.linalg.norm(w, axis=1, keepdims=True)
            if model == 2:   # transr/trainer.cpp:88-113: seeded from the TransE-unif run's 6-decimal files
                e0, r0 = tabs["transe_unif"]
                ent = np.round(e0, 6)
                ent = ent / np.linalg.norm(ent, axis=1, keepdims=True)
                rel = np.round(r0, 6).copy()
                w = np.ascontiguousarray(np.tile(np.eye(D), (nR, 1, 1)))
            if method not in samplers:
                samplers[method] = oracle.sampler(g["train"], nE, nR, method)
                samplers[method].set_mode(1 if arm == "ref_randmax" else 0)
            fn = samplers[method].train_epochs_dfr if arm.startswith("dfr") else samplers[method].train_epochs_ref
            fn(model, dist, LR, MARGIN, BATCHES, 0, EPOCHS[name], 2000 + seed, ent, rel, w)
            tabs[name] = (ent, rel)
            out[name] = metrics(oracle, model, dist, np.round(ent, 6), np.round(rel, 6), None if w is None else np.round(w, 6), g)
    return {"seed": seed, "arm": arm, "seconds": time.time() - t0, "models": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=12)
    ap.add_argument("--procs", type=int, default=6)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_stat_parity_cpu.json"))
    a = ap.parse_args()
    jobs = [(s, arm) for arm in ARMS for s in range(a.seeds)]
    with Pool(a.procs) as pool:
        rows = []
        for r in pool.imap_unordered(job, jobs):
            rows.append(r)
            print(r["seed"], r["arm"], f"{r['seconds']:.0f}s", {k: round(v["filt_mr"], 2) for k, v in r["models"].items()}, file=sys.stderr, flush=True)
    summary = {}
    for name, *_ in MODELS:
        per_arm = {}
        for arm in ARMS:
            sel = sorted((r["seed"], r["models"][name]) for r in rows if r["arm"] == arm and name in r["models"])
            if sel:
                per_arm[arm] = (np.array([m["filt_mr"] for _, m in sel]), np.array([m["filt_h10"] for _, m in sel]))
        summary[name] = {}
        for arm, (mr, h10) in per_arm.items():
            summary[name][arm] = {"filt_mr_mean": float(mr.mean()), "filt_mr_std": float(mr.std(ddof=1)), "filt_h10_mean": float(h10.mean()),
                                  "n": int(len(mr))}
        # paired differences (same data seed on both sides): mean and 95 % CI of the relative MeanRank difference
        for x, y in (("dfr", "ref_uniform"), ("dfr_nocarry", "dfr"), ("ref_randmax", "shipped"), ("ref_uniform", "shipped"), ("dfr", "shipped")):
            if x in per_arm and y in per_arm and len(per_arm[x][0]) == len(per_arm[y][0]):
                rel = (per_arm[x][0] - per_arm[y][0]) / per_arm[y][0]
                dh = 100 * (per_arm[x][1] - per_arm[y][1])
                n = len(rel)
                summary[name][f"{x}_vs_{y}"] = {"mr_rel_diff_mean": float(rel.mean()), "mr_rel_diff_ci95": float(1.96 * rel.std(ddof=1) / np.sqrt(n)),
                                               "h10_diff_points_mean": float(dh.mean()), "h10_diff_points_ci95": float(1.96 * dh.std(ddof=1) / np.sqrt(n)), "n": n}
    json.dump({"config": {"shape": SHAPE, "D": D, "lr": LR, "margin": MARGIN, "batches": BATCHES, "epochs": EPOCHS, "seeds": a.seeds},
               "rows": rows, "summary": summary}, open(a.out, "w"), indent=1)
    for name in summary:
        print(name)
        for k, v in summary[name].items():
            print("   ", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()})


if __name__ == "__main__":
    main()
