set -x
O=gpurun_out; R=r01
python -m pytest tests/test_gpu_train.py tests/test_gpu_programs.py -m gpu -x -q 2>&1 | tail -3
python tools/probe.py --model transe --dim 100 --distance 1 --epochs 20 --test 1000 2>&1 | grep epochs
python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 20 --test 1000 2>&1 | grep epochs
python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep epochs
KB2E_TRAIN_TRACE=$O/${R}_trace.txt python tools/probe.py --model transe --dim 100 --distance 1 --epochs 10 --test 10 > /dev/null 2>&1
python tools/trace_report.py $O/${R}_trace.txt 5 2>/dev/null | head -12
