set -x
timeout 300 python -m pytest tests/test_gpu_rank.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 5 --test 59071 2>&1 | grep rank
timeout 120 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 5 --test 5000 2>&1 | grep rank
