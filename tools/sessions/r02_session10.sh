#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_stat_parity.py tests/test_gpu_programs.py -m gpu -q -k "transh or TransH or resident or deterministic or stat or programs" > $O/r02_s10_pytest.txt 2>&1
tail -8 $O/r02_s10_pytest.txt
{
  echo "# config 2: TransH bern size=100, WN18 shape, relations resident in shared memory"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 5000 2>&1 | grep -E "epochs|rank|rror"
  echo "# the same with KB2E_TRANSH_SR=0 (three-barrier list kernel)"; KB2E_TRANSH_SR=0 timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  echo "# TransH size=50 WN18"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 50 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  KB2E_TRAIN_TRACE=$O/r02_s10_trace_transh.txt timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1
  python tools/trace_report.py $O/r02_s10_trace_transh.txt 5 2>&1 | head -19
} > $O/r02_s10_probes.txt 2>&1
cat $O/r02_s10_probes.txt
