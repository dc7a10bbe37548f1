"""Seeded synthetic "planted translation" knowledge graphs in the reference's on-disk format.

Recipe: latent z_e, z_r ~ N(0, I_d0); relation frequencies Zipf-like (FB15k shape) or
uniform (WN18 shape); per-relation fan-out k_r in 1..8 (a mix of 1-1 and 1-N relations so that
bern != unif and filtered != raw); a triple (h, r, t) draws h uniformly, r from the relation
distribution and t uniformly from the k_r entities nearest to z_h + z_r (exact nearest neighbours,
h excluded); duplicates removed, shuffled, split into train / valid / test of the exact counts.

Files (common/loader.cpp:15-62, common/constants.h:19-23): entity2id.txt / relation2id.txt hold
``name<TAB>id``; train/valid/test.txt hold ``head<TAB>tail<TAB>relation`` names.
"""
import os

import numpy as np

SHAPES = {
    # name: (entities, relations, train, valid, test, zipf)
    "fb15k": (14951, 1345, 483142, 50000, 59071, True),
    "wn18": (40943, 18, 141442, 5000, 5000, False),
    # throughput-only shape of BASELINE configs[4]; the reference cannot run it
    "scaled": (4_000_000, 1345, 100_000_000, 10_000, 10_000, True),
    "tiny": (500, 12, 6000, 300, 300, True),
    "small": (2000, 20, 40000, 1000, 1000, True),
}


def _nearest(zq, zE, k, exclude, use_torch):
    """Indices (n, k) of the k nearest rows of zE to each row of zq (excluding `exclude`), ascending distance."""
    n = len(zq)
    out = np.empty((n, k), dtype=np.int64)
    if use_torch:
        import torch

        dev = "cuda" if torch.cuda.is_available() else "cpu"
        E = torch.from_numpy(zE).to(dev)
        e2 = (E * E).sum(1)
        step = 16384 if dev == "cuda" else 4096
        for s in range(0, n, step):
            q = torch.from_numpy(zq[s:s + step]).to(dev)
            d = e2[None, :] - 2.0 * (q @ E.T)
            ex = torch.from_numpy(exclude[s:s + step]).to(dev)
            d[torch.arange(len(q), device=dev), ex] = float("inf")
            out[s:s + step] = torch.topk(d, k, dim=1, largest=False).indices.cpu().numpy()
        return out
    e2 = (zE * zE).sum(1)
    step = 2048
    for s in range(0, n, step):
        q = zq[s:s + step]
        d = e2[None, :] - 2.0 * (q @ zE.T)
        d[np.arange(len(q)), exclude[s:s + step]] = np.inf
        part = np.argpartition(d, k - 1, axis=1)[:, :k]
        order = np.argsort(np.take_along_axis(d, part, 1), axis=1)
        out[s:s + step] = np.take_along_axis(part, order, 1)
    return out


def make_kg(shape="small", seed=0, d0=16, use_torch=None):
    """Returns dict(nE, nR, train, valid, test) with (n, 3) int32 arrays of (head, tail, relation)."""
    if isinstance(shape, str):
        nE, nR, n_train, n_valid, n_test, zipf = SHAPES[shape]
    else:
        nE, nR, n_train, n_valid, n_test, zipf = shape
    if use_torch is None:
        use_torch = nE * (n_train + n_valid + n_test) > 2e9
    rng = np.random.default_rng(seed)
    zE = rng.standard_normal((nE, d0)).astype(np.float32)
    zR = rng.standard_normal((nR, d0)).astype(np.float32)
    if zipf:
        p = 1.0 / np.arange(1, nR + 1) ** 0.8
        p = p[rng.permutation(nR)]
    else:
        p = np.ones(nR)
    p = p / p.sum()
    fan = rng.integers(1, 9, size=nR)
    total = n_train + n_valid + n_test
    kmax = 8
    triples = np.empty((0, 3), dtype=np.int64)
    seen = None
    rounds = 0
    while len(triples) < total:
        rounds += 1
        if rounds > 20:
            raise RuntimeError("could not generate enough distinct triples for this shape")
        need = int((total - len(triples)) * 1.35) + 64
        h = rng.integers(0, nE, size=need)
        r = rng.choice(nR, size=need, p=p)
        nn = _nearest(zE[h] + zR[r], zE, kmax, h, use_torch)
        pick = (rng.random(need) * fan[r]).astype(np.int64)
        t = nn[np.arange(need), pick]
        new = np.stack([h, t, r], axis=1)
        triples = np.concatenate([triples, new])
        key = (triples[:, 2] * nE + triples[:, 0]) * nE + triples[:, 1]
        _, first = np.unique(key, return_index=True)
        triples = triples[np.sort(first)]
        del seen
        seen = None
    triples = triples[rng.permutation(len(triples))[:total]].astype(np.int32)
    # every relation and (as far as possible) every entity should appear in train; not required by the format
    return {
        "nE": nE, "nR": nR,
        "train": np.ascontiguousarray(triples[:n_train]),
        "valid": np.ascontiguousarray(triples[n_train:n_train + n_valid]),
        "test": np.ascontiguousarray(triples[n_train + n_valid:]),
    }


def write_kg(kg, datadir):
    """Write the five text files the reference's loaders read."""
    os.makedirs(datadir, exist_ok=True)
    with open(os.path.join(datadir, "entity2id.txt"), "w") as f:
        f.write("".join(f"e{i}\t{i}\n" for i in range(kg["nE"])))
    with open(os.path.join(datadir, "relation2id.txt"), "w") as f:
        f.write("".join(f"r{i}\t{i}\n" for i in range(kg["nR"])))
    for name in ("train", "valid", "test"):
        tr = kg[name]
        with open(os.path.join(datadir, name + ".txt"), "w") as f:
            f.write("".join(f"e{h}\te{t}\tr{r}\n" for h, t, r in tr.tolist()))


def bern_stats(train, nR):
    """Host-side restatement of common/trainer.cpp:163-194 (mean triples per distinct head / tail)."""
    head_mean = np.zeros(nR)
    tail_mean = np.zeros(nR)
    tr = np.asarray(train)
    for col, out in ((0, head_mean), (1, tail_mean)):
        key = tr[:, 2].astype(np.int64) * (tr[:, col].max() + 1) + tr[:, col]
        uniq = np.unique(key)
        distinct = np.bincount((uniq // (tr[:, col].max() + 1)).astype(np.int64), minlength=nR)
        count = np.bincount(tr[:, 2], minlength=nR)
        nz = distinct > 0
        out[nz] = count[nz] / distinct[nz]
    return head_mean, tail_mean


def write_embeddings(path, table):
    """The reference's "%.6lf\\t" cells, one row per line (common/trainer.cpp:109-127)."""
    with open(path, "w") as f:
        for row in np.asarray(table, dtype=np.float64):
            f.write("".join("%.6f\t" % v for v in row) + "\n")


def read_embeddings(path, rows, cols):
    return np.loadtxt(path, dtype=np.float64).reshape(rows, cols)
