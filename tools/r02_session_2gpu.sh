#!/bin/bash
# 2-GPU session: partitioned-trainer parity after the robustness changes, failure injection (a rank that never launches),
# throughput at the scaled shape, eval programs with --gpus 2, and the bench line at N = 2.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dist_check.py --abort-test --skip-throughput > $O/r02_dist_check_2gpu_abort.json 2> $O/r02_dist_check_2gpu.err || tail -20 $O/r02_dist_check_2gpu.err
cat $O/r02_dist_check_2gpu_abort.json | cut -c1-1500
timeout 900 $TR tools/dist_check.py --skip-parity --epochs 2 > $O/r02_dist_check_2gpu.json 2>> $O/r02_dist_check_2gpu.err || tail -20 $O/r02_dist_check_2gpu.err
cat $O/r02_dist_check_2gpu.json | cut -c1-800
timeout 600 python -m pytest tests/test_gpu_programs.py -m gpu -q -k "shards or binding" 2>&1 | tail -3
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02_bench_2gpu.json 2> $O/r02_bench_2gpu.err || tail -20 $O/r02_bench_2gpu.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02_bench_2gpu.json') if l.startswith('{')][-1]   # (NCCL prints its version line first)
print('N=2 train value %.1f M/s | eval value %.1f M q/s e2e cold %.1f resident %.1f | partitioned %s' % (d['value']/1e6, d['eval']['value']/1e6, d['eval']['e2e']['value']/1e6, d['eval']['e2e']['resident']['value']/1e6, {k: d['partitioned'].get(k) for k in ('value','ms_per_epoch','error')} if d.get('partitioned') else None))
PY
