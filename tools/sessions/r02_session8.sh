#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r02_s8_pytest.txt 2>&1
tail -6 $O/r02_s8_pytest.txt
{
  echo "# config 4 on ONE GPU: TransE L2 size=200, scaled shape, random triples"; timeout 600 python tools/probe.py --shape scaled --dim 200 --random --epochs 2 --test 10 2>&1 | grep -E "epochs|rror"
  echo "# config 1: TransE bern L2 size=100, FB15k shape"; timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 20 --test 100 2>&1 | grep -E "epochs|rror"
  echo "# config 0: TransE unif L1 size=50, FB15k shape"; timeout 300 python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|rror"
  echo "# config 2: TransH bern size=100, WN18 shape"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|rror"
  echo "# batched"; timeout 300 python tools/probe_sweep.py --models 8 2>&1 | tail -2
} > $O/r02_s8_probes.txt 2>&1
cat $O/r02_s8_probes.txt
