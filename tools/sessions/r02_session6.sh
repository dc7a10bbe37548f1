#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "batched or deterministic or sampler" > $O/r02_s6_pytest.txt 2>&1
tail -25 $O/r02_s6_pytest.txt
timeout 900 python tools/probe_sweep.py --models 1,2,4,8,12,16 > $O/r02_s6_sweep_fb15k_d100.txt 2>&1; cat $O/r02_s6_sweep_fb15k_d100.txt
timeout 900 python tools/probe_sweep.py --dim 50 --distance 0 --method 0 --models 1,4,8,16 > $O/r02_s6_sweep_fb15k_d50.txt 2>&1; cat $O/r02_s6_sweep_fb15k_d50.txt
timeout 900 python tools/probe_sweep.py --shape wn18 --dim 100 --distance 0 --models 1,4,16,32 > $O/r02_s6_sweep_wn18.txt 2>&1; cat $O/r02_s6_sweep_wn18.txt
