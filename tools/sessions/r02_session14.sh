#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/r02_s14_pytest.txt 2>&1
tail -8 $O/r02_s14_pytest.txt
{
  echo "# config 1"; timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 20 --test 100 2>&1 | grep -E "epochs|rror"
  echo "# config 0"; timeout 300 python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|rror"
  echo "# config 2 (TransH WN18, SR)"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  echo "# config 2 with KB2E_TRANSH_SR=0"; KB2E_TRANSH_SR=0 timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  echo "# config 3 (TransR)"; timeout 300 python tools/probe.py --model transr --dim 50 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  echo "# TransE L1 WN18"; timeout 300 python tools/probe.py --model transe --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  echo "# sweep"; timeout 300 python tools/probe_sweep.py --models 8 2>&1 | tail -2
  KB2E_TRAIN_TRACE=$O/r02_s14_fine.txt KB2E_TRAIN_TRACE_FINE=1 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1; python tools/trace_transh_sr_fine.py $O/r02_s14_fine.txt
} > $O/r02_s14_probes.txt 2>&1
cat $O/r02_s14_probes.txt
