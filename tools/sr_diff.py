"""Tuning aid: TransH shared-memory-resident kernel vs the three-barrier list kernel, deterministic mode, table differences."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kb2e_b200
from kb2e_b200 import kg

g = kg.make_kg("tiny", seed=6)
nE, nR = g["nE"], g["nR"]
hm, tm = kg.bern_stats(g["train"], nR)


def run(sr, D, batches, epochs):
    os.environ["KB2E_TRANSH_SR"] = "1" if sr else "0"
    with kb2e_b200.Context("transh", D, nE, nR, method=1, distance=0, batches=batches, rate=0.01, margin=1.0, seed=77,
                           flags=kb2e_b200.FLAG_DETERMINISTIC, device=0) as ctx:
        ctx.set_train_triples(g["train"])
        ctx.set_bern(hm, tm)
        ctx.init_embeddings()
        loss = ctx.train_epochs(0, epochs)
        return loss, ctx.download(kb2e_b200.TABLE_ENTITY), ctx.download(kb2e_b200.TABLE_RELATION), ctx.download(kb2e_b200.TABLE_WEIGHTS)


for D in (100, 50):
    for batches, epochs in ((6000, 1), (7, 1), (7, 3)):
        a, b, c = run(True, D, batches, epochs), run(False, D, batches, epochs), run(True, D, batches, epochs)
        msg = []
        for name, x, y, z in zip(("loss", "ent", "rel", "w"), a, b, c):
            d = np.abs(np.asarray(x) - np.asarray(y))
            msg.append("%s: %d differ, max %.3g (sr vs sr: %d differ)" % (name, int((d > 0).sum()), d.max(), int((np.asarray(x) != np.asarray(z)).sum())))
        print("D=%d batches=%d epochs=%d | " % (D, batches, epochs) + " | ".join(msg))
