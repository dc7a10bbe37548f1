set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/probe.py --model transr --dim 50 --distance 0 --epochs 5 --test 2000 2>&1 | tail -8
KB2E_TRAIN_TRACE=gpurun_out/trace_transr.txt python tools/probe.py --model transr --dim 50 --distance 0 --epochs 1 --test 10 > /dev/null 2>&1
python tools/trace_report.py gpurun_out/trace_transr.txt 5 | head -24
python tools/probe.py --model transe --dim 100 --distance 1 --epochs 5 --test 59071 2>&1 | tail -5
python tools/e2e_probe.py 2>&1 | tail -40
