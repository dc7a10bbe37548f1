"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (our kernels only)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
cols = rows[hdr]
ix = {c: i for i, c in enumerate(cols)}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) != len(cols):
        continue
    name = r[ix["Kernel Name"]]
    ours = ("train_kernel", "train_fused", "train_transr", "train_dist", "train_lazy", "rank_", "filter_", "recheck", "etrue", "prep_",
            "finalize", "build_queries", "transpose_kernel", "widen_kernel", "narrow_kernel", "segment_", "hash_insert", "pack_triples",
            "init_rows", "count_chunks", "project_", "score", "sample_kernel", "DeviceRadixSort", "train_sweep", "query_kernel",
            "thresholds_kernel", "validate_triples", "fill_threshold", "widen_delta", "fill_int", "identity_kernel")
    if not any(k in name for k in ours):
        continue   # torch kernels of the synthetic-KG generator / L2 flush
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    short = name.split("(")[0][:70]
    tot[short] += us
    cnt[short] += 1
total = sum(tot.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline")
print("(cold-cache, serialised per-launch times: compare SHARES, not absolutes; torch kernels of the KG generator / L2 flush are filtered out)\n")
for k, v in tot.most_common():
    print("%12.1f us  %6.2f%%  x%4d  %s" % (v, 100 * v / total, cnt[k], k))
