import numpy as np, sys
t = np.loadtxt(sys.argv[1], dtype=np.uint64).astype(np.int64)
# per batch 9 slots: start, p1 end, bar1 end, [fine0 loads, fine1 norm, fine2 loop, fine3 steps], 2a end, 2b end
S = 9
nb = t.shape[1] // S
for b in range(2, min(8, nb - 1)):
    seg = t[:, b * S:(b + 1) * S + 1]
    bar1 = seg[:, 2]
    ok = seg[:, 3] > 0
    print("batch", b, "ctas with stamps", ok.sum(),
          "loads arrived +%d" % np.median(seg[ok, 3] - bar1[ok]), "norm +%d" % np.median(seg[ok, 4] - seg[ok, 3]),
          "hyperplane +%d" % np.median(seg[ok, 5] - seg[ok, 4]), "steps med %d max %d" % (np.median(seg[ok, 6] - 1000000), (seg[ok, 6] - 1000000).max()),
          "2a end +%d" % np.median(seg[ok, 7] - seg[ok, 5]), "2a total %d" % np.median(seg[:, 7] - bar1))
