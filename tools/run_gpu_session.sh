# scratch driver for one gpurun call: `bash tools/run_gpu_session.sh N` runs bench.py on N GPUs (N = 1: plus the reference arm, smoke and the GPU tests)
set -x
N=${1:-1}
O=gpurun_out
if [ "$N" = "1" ]; then
  ( time python bench.py ) > $O/r01_bench.json 2> $O/r01_bench.err; tail -4 $O/r01_bench.err
  python bench.py --impl reference --steps 1 --warmup 0 > $O/r01_bench_reference.json 2>> $O/r01_bench.err
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
  python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/r01_pytest_gpu.txt; cat $O/r01_pytest_gpu.txt
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > $O/r01_bench_${N}gpu.json 2> $O/bench${N}.err
  tail -2 $O/bench${N}.err
fi
F=$O/r01_bench_${N}gpu.json; [ "$N" = "1" ] && F=$O/r01_bench.json
python -c "
import json
d=json.loads(open('$F').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'eval', d['eval']['value'], d['eval']['ms_per_step'], 'eval e2e', d['eval']['e2e']['value'], 'frac', d['roofline']['frac'])
p=d['partitioned']; print({k:p.get(k) for k in ('value','ms_per_epoch','n_gpus','error')}, p.get('roofline',{}).get('frac'))
"
