"""SASS instruction histogram of the product library for profiles/ (cuobjdump -sass; needs no GPU).

    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kb2e_b200", "lib", "libkb2e_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UBLKCP", "SYNCS", "REDG.E.ADD.F32x4", "RED.E.ADD.F32x4", "ATOMG.E.EXCH", "MUFU.RCP",
         "MUFU.RSQ", "FCHK", "DADD", "DFMA", "FFMA", "BAR.SYNC", "HMMA", "CALL.REL"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
per, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if m and name:
        per[name]["instr"] += 1
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                per[name][w] += 1
total = collections.Counter()
for c in per.values():
    total.update(c)
print("# SASS instruction histogram of kb2e_b200/lib/libkb2e_b200.so (cuobjdump -sass, sm_100a); tools/sass_histogram.py")
print("# Blackwell-only mnemonics: UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit -> mbarrier, UBLKCP / UTMALDG = TMA loads,")
print("# SYNCS = mbarrier ops; REDG.E.ADD.F32x4 = vector float RED, ATOMG.E.EXCH = publish-time row claim; MUFU.RCP / FCHK / CALL.REL: IEEE division sequences.")
print("# whole library: %d kernels, %d instructions; " % (len(per), total["instr"]) + "  ".join("%s=%d" % (w, total[w]) for w in WATCH if total[w]))
keep = ("rank_l2_tc", "project_tc", "train_kernelILi0ELi16ELi2ELi640ELb1ELb0", "train_kernelILi0ELi32ELi2ELi768ELb0ELb0", "train_sweep_kernelILi16ELi2ELi640ELb0",
        "train_transh_sr_kernelILi32ELi1ELi512ELb0", "train_kernelILi1ELi32ELi1ELi512ELb1ELb0", "train_transr_kernelILi2", "rank_f32_kernelILi0", "train_dist", "filter_pairs", "recheck")
for n, c in per.items():
    if any(k in n for k in keep):
        print("%-110s instr %6d  " % (demangle(n)[:110], c["instr"]) + "  ".join("%s=%d" % (w, c[w]) for w in WATCH if c[w]))
