#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_train.py tests/test_gpu_stat_parity.py tests/test_gpu_programs.py -m gpu -q > $O/r02_s15_pytest.txt 2>&1
tail -4 $O/r02_s15_pytest.txt
{
  echo "# config 2 (TransH WN18, relations resident in shared memory)"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 5000 2>&1 | grep -E "epochs|rank|rror"
  echo "# config 2 with KB2E_TRANSH_SR=0 (three-barrier list kernel)"; KB2E_TRANSH_SR=0 timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  echo "# TransH size=100 FB15k shape (1,345 relations: three-barrier list kernel)"; timeout 300 python tools/probe.py --model transh --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  echo "# TransH size=50 WN18"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 50 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
} > $O/r02_transh_probes.txt 2>&1
cat $O/r02_transh_probes.txt
KB2E_TRAIN_TRACE=$O/r02_s15_trace_transh.txt timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1
python tools/trace_report.py $O/r02_s15_trace_transh.txt 5 > $O/r02_trace_transh_sr_report.txt 2>&1
KB2E_TRAIN_TRACE=$O/r02_s15_fine.txt KB2E_TRAIN_TRACE_FINE=1 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1; python tools/trace_transh_sr_fine.py $O/r02_s15_fine.txt >> $O/r02_trace_transh_sr_report.txt
head -13 $O/r02_trace_transh_sr_report.txt; tail -4 $O/r02_trace_transh_sr_report.txt
KB2E_TRANSH_SR=0 KB2E_TRAIN_TRACE=$O/r02_s15_trace_transh0.txt timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1
python tools/trace_report.py $O/r02_s15_trace_transh0.txt 5 > $O/r02_trace_transh_report.txt 2>&1
head -7 $O/r02_trace_transh_report.txt
