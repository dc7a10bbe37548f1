#!/bin/bash
# GPU session 1 of round 2: the GPU test suite, then A/B probes of the training kernels (touched-row lists on / off)
# and the TransR ranking path.  Outputs under gpurun_out/.
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > $O/r02_s1_pytest.txt
cat $O/r02_s1_pytest.txt
{
  for list in 1 0; do
    export KB2E_TRAIN_LIST=$list
    echo "## KB2E_TRAIN_LIST=$list"
    echo "# config 0: TransE unif L1 size=50, FB15k shape"; timeout 300 python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
    echo "# config 1: TransE bern L2 size=100, FB15k shape"; timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
    echo "# config 2: TransH bern size=100, WN18 shape"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
    echo "# TransH bern size=100, FB15k shape"; timeout 300 python tools/probe.py --model transh --dim 100 --distance 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  done
  unset KB2E_TRAIN_LIST
  echo "# config 3: TransR size=50 L1, FB15k shape, all 59071 test triples"; timeout 600 python tools/probe.py --model transr --dim 50 --distance 0 --epochs 5 --test 59071 2>&1 | grep -E "epochs|rank|Error|error"
  echo "# TransR size=50 squared L2, FB15k shape"; timeout 600 python tools/probe.py --model transr --dim 50 --distance 1 --epochs 5 --test 59071 2>&1 | grep -E "epochs|rank|Error|error"
  echo "# config 1 ranking: TransE L2 size=100"; timeout 600 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 5 --test 59071 2>&1 | grep -E "rank|Error|error"
} > $O/r02_s1_probes.txt 2>&1
cat $O/r02_s1_probes.txt
KB2E_TRAIN_TRACE=$O/r02_s1_trace.txt timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 10 --test 10 > /dev/null 2>&1
python tools/trace_report.py $O/r02_s1_trace.txt 5 > $O/r02_s1_trace_report.txt 2>/dev/null
cat $O/r02_s1_trace_report.txt | tail -30
