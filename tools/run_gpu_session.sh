# scratch driver for one gpurun call: `bash tools/run_gpu_session.sh N`: bench.py on N GPUs (N = 1: plus reference arm, launch list, ncu capture of the ranking kernel, config probes, GPU tests)
set -x
N=${1:-1}
O=gpurun_out; R=r01
if [ "$N" = "1" ]; then
  python bench.py > $O/${R}_bench.json 2> $O/${R}_bench.err; tail -2 $O/${R}_bench.err
  python bench.py --impl reference --steps 1 --warmup 0 > $O/${R}_bench_reference.json 2>> $O/${R}_bench.err
  python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${R}_pytest_gpu.txt; cat $O/${R}_pytest_gpu.txt
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"train_kernel|rank_l2_tc|rank_exact|rank_f32|filter_|recheck|etrue|prep_|finalize|build_queries|transpose_kernel|widen_kernel|narrow_kernel|segment_|hash_insert|pack_triples|init_rows|count_chunks|DeviceRadixSort" -c 400 --csv --log-file $O/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-partitioned > $O/${R}_ncu_launch.log 2>&1
  python tools/launch_list_summary.py $O/${R}_launches.csv > $O/${R}_launch_list_summary.txt 2>&1
  ncu --set full --clock-control none --import-source on -k regex:rank_l2_tc -s 1 -c 1 -f -o $O/${R}_rank python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-partitioned > $O/${R}_ncu_rank.log 2>&1
  python tools/ncu_summary.py $O/${R}_ncu_rank_summary.txt rank=$O/${R}_rank.ncu-rep > /dev/null 2>&1; cp profiles/traffic.json $O/${R}_traffic.json
  {
    echo "# config 0: TransE unif L1 size=50, FB15k shape"; python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 20 --test 59071 2>&1 | grep -E "epochs|rank"
    echo "# config 1: TransE bern L2 size=100, FB15k shape"; python tools/probe.py --model transe --dim 100 --distance 1 --epochs 20 --test 59071 2>&1 | grep -E "epochs|rank"
    echo "# config 2: TransH bern size=100, WN18 shape"; python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 5000 2>&1 | grep -E "epochs|rank"
    echo "# config 3: TransR size=50 L1, FB15k shape"; python tools/probe.py --model transr --dim 50 --distance 0 --epochs 20 --test 5000 2>&1 | grep -E "epochs|rank"
    echo "# config 4 on ONE GPU: TransE L2 size=200, scaled shape, random triples"; python tools/probe.py --shape scaled --dim 200 --random --epochs 2 --test 10 2>&1 | grep -E "epochs"
    echo "# TransE L1 size=100, WN18 shape (one-barrier kernel, small batch)"; python tools/probe.py --model transe --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs"
  } > $O/${R}_config_probes.txt 2>&1
  grep -E "rank " $O/${R}_config_probes.txt | cut -c1-200
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > $O/${R}_bench_${N}gpu.json 2> $O/bench${N}.err
  tail -2 $O/bench${N}.err
fi
F=$O/${R}_bench_${N}gpu.json; [ "$N" = "1" ] && F=$O/${R}_bench.json
python -c "
import json
d=json.loads(open('$F').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], 'eval', d['eval']['value'], d['eval']['ms_per_step'], 'eval e2e', d['eval']['e2e']['value'], 'frac', d['roofline']['frac'], 'eval roof', d['eval']['roofline']['frac'], d['eval']['roofline']['executed_tflops'], d['eval']['roofline']['main_kernel_ms'])
p=d['partitioned']; print({k:p.get(k) for k in ('value','ms_per_epoch','n_gpus','error')}, p.get('roofline',{}).get('frac'))
print('cpu', d.get('cpu_baseline'), d['eval'].get('cpu_baseline'))
"
