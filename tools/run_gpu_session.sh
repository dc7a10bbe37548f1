set -x
python -m pytest tests/test_gpu_train.py -m gpu -x -q -k partitioned 2>&1 | tail -3
KB2E_TRAIN_TRACE=gpurun_out/dtrace3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py --shape scaled --epochs 1 2>&1 | tail -4
python tools/trace_report.py gpurun_out/dtrace3.0 11 | head -13
