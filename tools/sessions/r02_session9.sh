#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/r02_s9_pytest.txt 2>&1
tail -8 $O/r02_s9_pytest.txt
{
  echo "# config 3: TransR size=50 L1, FB15k shape (grid-wide phase 2b list)"; timeout 300 python tools/probe.py --model transr --dim 50 --distance 0 --epochs 20 --test 59071 2>&1 | grep -E "epochs|rank|rror"
  KB2E_TRAIN_TRACE=$O/r02_s9_trace_transr.txt timeout 300 python tools/probe.py --model transr --dim 50 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1
  python tools/trace_report.py $O/r02_s9_trace_transr.txt 5 2>&1 | tail -30
} > $O/r02_s9_probes.txt 2>&1
cat $O/r02_s9_probes.txt
