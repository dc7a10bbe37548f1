set -x
python -m pytest tests/test_gpu_rank.py tests/test_gpu_programs.py tests/test_gpu_stat_parity.py -m gpu -x -q 2>&1 | tail -5
python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 5 --test 59071 2>&1 | grep rank
python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 5 --test 5000 2>&1 | grep rank
python tools/probe.py --model transr --dim 50 --distance 0 --epochs 5 --test 5000 2>&1 | grep rank
python tools/probe.py --model transr --dim 50 --distance 1 --epochs 5 --test 5000 2>&1 | grep rank
python tools/stat_parity.py --stage gpu --out gpurun_out/stat_parity_r01.json > gpurun_out/r01_stat_parity.txt 2>&1; tail -5 gpurun_out/r01_stat_parity.txt
