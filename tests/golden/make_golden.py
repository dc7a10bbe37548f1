"""Generate tests/golden/kb2e_golden.npz from the UNMODIFIED reference compiled here.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
It drives oracle/_ref/libkb2e_ref.so (oracle/ref_harness.cpp + the reference's own sources, built by
oracle/Makefile) on seeded random tables and stores inputs and the reference's outputs.  The
reference ships no fixtures of its own, so these files are what pins the oracle:
tests/test_oracle_vs_reference.py demands bitwise equality on every array.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from kb2e_oracle import Reference, build  # noqa: E402


def tables(rng, model, nE, nR, D):
    ent = rng.normal(0, 0.35, (nE, D))
    ent /= np.maximum(1.0, np.linalg.norm(ent, axis=1, keepdims=True))
    rel = rng.normal(0, 0.3, (nR, D))
    w = np.zeros((1, 1))
    if model == 1:
        w = rng.normal(0, 1, (nR, D))
        w /= np.linalg.norm(w, axis=1, keepdims=True)
    if model == 2:
        ent /= np.linalg.norm(ent, axis=1, keepdims=True)
        rel /= np.linalg.norm(rel, axis=1, keepdims=True)
        w = np.tile(np.eye(D), (nR, 1, 1)) + rng.normal(0, 0.05, (nR, D, D))
        w /= np.linalg.norm(w, axis=2, keepdims=True)
    # duplicate two entity rows so that exact energy ties exist
    ent[nE - 1] = ent[3]
    ent[nE - 2] = ent[5]
    return ent, rel, w


def main():
    build()
    ref = Reference()
    rng = np.random.default_rng(20261018)
    out = {}
    cases = []
    for model in (0, 1, 2):
        for D in (6, 20):
            for dist in (0, 1):
                if model == 1 and dist == 1:
                    continue
                cases.append((model, D, dist))
    out["cases"] = np.array(cases, dtype=np.int32)
    nE, nR, n = 48, 5, 40
    for ci, (model, D, dist) in enumerate(cases):
        ent, rel, w = tables(rng, model, nE, nR, D)
        wq = None if model == 0 else w
        h = rng.integers(0, nE, n).astype(np.int32)
        t = rng.integers(0, nE, n).astype(np.int32)
        r = rng.integers(0, nR, n).astype(np.int32)
        c = rng.integers(0, nE, n).astype(np.int32)
        side = rng.integers(0, 2, n)
        pairs = np.stack([h, t, r, np.where(side == 1, h, c), np.where(side == 1, c, t), r], 1).astype(np.int32)
        k = f"c{ci}_"
        out[k + "ent"], out[k + "rel"], out[k + "w"] = ent, rel, w
        out[k + "pairs"] = pairs
        out[k + "energy"] = ref.energy(model, dist, ent, rel, wq, h, t, r)
        if model == 2:
            out[k + "energy_shipped"] = ref.energy(model, dist, ent, rel, wq, h, t, r, zero_work=False)
        for corrupted in (0, 1):
            en, rn, wn = ref.grad(model, dist, 0.01, ent, rel, wq, h[0], t[0], r[0], corrupted)
            out[k + f"grad{corrupted}_ent"], out[k + f"grad{corrupted}_rel"] = en, rn
            out[k + f"grad{corrupted}_w"] = np.zeros((1, 1)) if wn is None else wn
        en, rn, wn, losses = ref.train_batch(model, dist, 0.01, 1.0, ent, rel, wq, pairs)
        out[k + "batch_ent"], out[k + "batch_rel"], out[k + "batch_losses"] = en, rn, losses
        out[k + "batch_w"] = np.zeros((1, 1)) if wn is None else wn
        test = np.stack([h[:12], t[:12], r[:12]], 1).astype(np.int32)
        # make the first test triples hit the duplicated rows (ties) and share (h, r) keys with the filter set
        test[0] = (3, 7, 1)
        test[1] = (9, 5, 2)
        filt = np.stack([h[12:], t[12:], r[12:]], 1).astype(np.int32)
        filt = np.concatenate([filt, np.array([[3, 8, 1], [3, 9, 1], [11, 7, 1], [nE - 1, 7, 1]], dtype=np.int32)])
        raw, flt = ref.rank(model, dist, ent, rel, wq, test, filt)
        out[k + "test"], out[k + "filt"], out[k + "rank_raw"], out[k + "rank_filt"] = test, filt, raw, flt
    # normalisation helpers
    a = rng.normal(0, 0.4, 24)
    b = rng.normal(0, 1.0, 24)
    out["norm_in"] = a
    out["norm_clip"] = ref.norm(a * 3, True)
    out["norm_short"] = ref.norm(a * 0.1, True)
    out["norm_unit"] = ref.norm(a, False)
    na, nb = ref.norm2(np.abs(a), np.abs(b), 0.01)  # positive overlap -> the corrective loop runs
    out["norm2_a_in"], out["norm2_b_in"], out["norm2_a"], out["norm2_b"] = np.abs(a), np.abs(b), na, nb
    path = os.path.join(ROOT, "tests", "golden", "kb2e_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(cases), "cases")


if __name__ == "__main__":
    main()
