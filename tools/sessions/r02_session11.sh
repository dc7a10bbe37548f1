#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_train.py -m gpu -q -k "transh or resident or deterministic or transr or batch_matches" > $O/r02_s11_pytest.txt 2>&1
tail -5 $O/r02_s11_pytest.txt
{
  echo "# config 2: TransH bern size=100, WN18 shape, relations resident in shared memory (half-warp relation groups)"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  echo "# the same with KB2E_TRANSH_SR=0 (three-barrier list kernel)"; KB2E_TRANSH_SR=0 timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  echo "# TransH size=50 WN18"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 50 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  echo "# config 3: TransR (single-window relation compaction)"; timeout 300 python tools/probe.py --model transr --dim 50 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  KB2E_TRAIN_TRACE=$O/r02_s11_trace_transh.txt timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1
  python tools/trace_report.py $O/r02_s11_trace_transh.txt 5 > $O/r02_trace_transh_sr_report.txt 2>&1
  head -13 $O/r02_trace_transh_sr_report.txt
  KB2E_TRAIN_TRACE=$O/r02_s11_trace_transr.txt timeout 300 python tools/probe.py --model transr --dim 50 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1
  python tools/trace_report.py $O/r02_s11_trace_transr.txt 6 2>&1 | head -15
  timeout 120 tools/microbench 2>&1 | grep -E "barrier"
} > $O/r02_s11_probes.txt 2>&1
cat $O/r02_s11_probes.txt
