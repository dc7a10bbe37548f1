#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -s > $O/r02_s3_pytest_full.txt 2>&1
tail -12 $O/r02_s3_pytest_full.txt
grep -E "^gpu_|mean filt MR" $O/r02_s3_pytest_full.txt | head -20
{
  echo "# config 0: TransE unif L1 size=50, FB15k shape"; timeout 300 python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  echo "# config 1: TransE bern L2 size=100, FB15k shape"; timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  echo "# config 2: TransH bern size=100, WN18 shape"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  echo "# TransE L1 size=100, WN18 shape"; timeout 300 python tools/probe.py --model transe --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|Error|error"
  echo "# TransE L1 size=100, WN18 shape, fused off"; KB2E_TRAIN_FUSED=0 timeout 300 python tools/probe.py --model transe --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|Error|error"
} > $O/r02_s3_probes.txt 2>&1
cat $O/r02_s3_probes.txt
KB2E_TRAIN_TRACE=$O/r02_s3_trace.txt timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 10 --test 10 > /dev/null 2>&1
python tools/trace_report.py $O/r02_s3_trace.txt 5 > $O/r02_s3_trace_report.txt 2>/dev/null
tail -12 $O/r02_s3_trace_report.txt
KB2E_TRAIN_TRACE_FINE=1 KB2E_TRAIN_TRACE=$O/r02_s3_trace_fine.txt timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 10 --test 10 > /dev/null 2>&1
{
  echo "# config 1 ranking: TransE L2 size=100 (flat filter pairs, filter beside the tensor-core kernel)"; timeout 600 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 5 --test 59071 2>&1 | grep -E "rank|Error|error"
  echo "# same, filter after the tensor-core kernel"; KB2E_RANK_NO_OVERLAP=1 timeout 600 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 5 --test 59071 2>&1 | grep -E "rank|Error|error"
  echo "# config 0 ranking: TransE L1 size=50"; timeout 600 python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 5 --test 59071 2>&1 | grep -E "rank|Error|error"
  echo "# config 3 ranking: TransR L1 size=50"; timeout 600 python tools/probe.py --model transr --dim 50 --distance 0 --epochs 5 --test 59071 2>&1 | grep -E "rank|Error|error"
  echo "# config 2 ranking: TransH size=100 WN18"; timeout 600 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 5 --test 5000 2>&1 | grep -E "rank|Error|error"
} > $O/r02_s3_rank_probes.txt 2>&1
cat $O/r02_s3_rank_probes.txt
timeout 600 python tools/e2e_probe.py > $O/r02_s3_e2e_probe.txt 2>&1; tail -25 $O/r02_s3_e2e_probe.txt
