"""Summarise a KB2E_TRAIN_TRACE file: per-phase times (ns) across CTAs for the traced batches."""
import sys
import numpy as np
t = np.loadtxt(sys.argv[1], dtype=np.uint64).astype(np.int64)
stamps = int(sys.argv[2]) if len(sys.argv) > 2 else 5   # stamps per batch (5 for TransE, 5 for H too)
names = {12: ["1a sample+request", "barrier", "1s serve", "barrier", "1b compute", "barrier", "2a push", "barrier", "2a' absorb",
              "2b entities", "2b relations", "barrier"],
         2: ["phase1 + fold", "barrier"],
         11: ["1a sample+request", "barrier", "1s serve", "barrier", "1b compute", "barrier", "2a push", "barrier", "2b entities",
              "2b relations", "barrier"],
         6: ["phase1", "barrier", "2a relations", "barrier", "2b entities", "barrier"],
         5: ["phase1", "barrier1", "phase2a(+bar)", "phase2b", "barrier2"],
         7: ["phase1 gather+compute", "barrier", "2a push", "barrier", "2b entities", "2b relations+prefetch", "barrier"],
         8: ["1a gather", "1b compute", "barrier", "2a push", "barrier", "2b entities", "2b relations+prefetch", "barrier"]}[stamps]
t0 = t[:, :1].min()
nb = (t.shape[1]) // stamps
for b in range(min(nb - 1, 6)):
    seg = t[:, b * stamps:(b + 1) * stamps + 1]
    if (seg == 0).any():
        break
    d = np.diff(seg, axis=1)
    print("batch", b, "start spread %d ns" % (seg[:, 0].max() - seg[:, 0].min()), "total (cta0) %d ns" % (seg[0, -1] - seg[0, 0]))
    for k, n in enumerate(names):
        print("   %-14s min %6d  median %6d  max %6d" % (n, d[:, k].min(), np.median(d[:, k]), d[:, k].max()))
