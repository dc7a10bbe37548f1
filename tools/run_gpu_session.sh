set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/probe.py --model transe --dim 100 --distance 1 --epochs 2 --test 59071 2>&1 | tail -2
python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 5 --test 59071 2>&1 | tail -3
python tools/e2e_probe.py 2>&1 | grep iteration
