#!/bin/bash
# GPU session 2 of round 2: full GPU test suite (complete log kept), training probes after the publish-time claim and the
# pipelined sampler, per-phase trace, bench line.
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -s > $O/r02_s2_pytest_full.txt 2>&1
tail -15 $O/r02_s2_pytest_full.txt
grep -E "transr projection|vs |filt MR" $O/r02_s2_pytest_full.txt | head -40
{
  echo "# config 0: TransE unif L1 size=50, FB15k shape"; timeout 300 python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  echo "# config 1: TransE bern L2 size=100, FB15k shape"; timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  echo "# config 1, 768 threads"; KB2E_TRAIN_NO640=1 timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  echo "# config 2: TransH bern size=100, WN18 shape"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  echo "# TransH bern size=100, FB15k shape"; timeout 300 python tools/probe.py --model transh --dim 100 --distance 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  echo "# TransE L1 size=100, WN18 shape"; timeout 300 python tools/probe.py --model transe --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|Error|error"
  echo "# TransE L1 size=100, WN18 shape, fused off"; KB2E_TRAIN_FUSED=0 timeout 300 python tools/probe.py --model transe --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|Error|error"
} > $O/r02_s2_probes.txt 2>&1
cat $O/r02_s2_probes.txt
KB2E_TRAIN_TRACE=$O/r02_s2_trace.txt timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 10 --test 10 > /dev/null 2>&1
python tools/trace_report.py $O/r02_s2_trace.txt 5 > $O/r02_s2_trace_report.txt 2>/dev/null
tail -12 $O/r02_s2_trace_report.txt
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02_s2_bench.json 2> $O/r02_s2_bench.err || tail -5 $O/r02_s2_bench.err
cut -c1-1500 $O/r02_s2_bench.json
