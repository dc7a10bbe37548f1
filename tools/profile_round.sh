#!/bin/bash
# Runs on the GPU box (gpurun): the round's bench line (both arms), the ncu launch list of the same command, one
# `ncu --set full` capture per dominant kernel, and the per-config probes quoted in profiles/README.md.  Outputs under gpurun_out/.
R=${1:-r02}
O=gpurun_out
mkdir -p $O
set -x
python bench.py --steps 5 --warmup 3 > $O/${R}_bench.json 2> $O/${R}_bench.err || tail -5 $O/${R}_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > $O/${R}_bench_reference.json 2>> $O/${R}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${R}_launches_all.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${R}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:train_kernel -s 3 -c 1 -f -o $O/${R}_train python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-partitioned > $O/${R}_ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rank_l2_tc -s 1 -c 1 -f -o $O/${R}_rank python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-partitioned > $O/${R}_ncu_rank.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:train_sweep -s 2 -c 1 -f -o $O/${R}_sweep python tools/probe_sweep.py --models 8 --epochs 2 > $O/${R}_ncu_sweep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:project_tc -s 1 -c 1 -f -o $O/${R}_project python tools/probe.py --model transr --dim 50 --distance 0 --epochs 2 --test 59071 > $O/${R}_ncu_project.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rank_f32 -s 1 -c 1 -f -o $O/${R}_rankf32 python tools/probe.py --model transr --dim 50 --distance 0 --epochs 2 --test 59071 > $O/${R}_ncu_rankf32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:train_transr -s 3 -c 1 -f -o $O/${R}_transr python tools/probe.py --model transr --dim 50 --distance 0 --epochs 6 --test 10 > $O/${R}_ncu_transr.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:train_transh_sr|train_kernel" -s 3 -c 1 -f -o $O/${R}_transh python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 6 --test 10 > $O/${R}_ncu_transh.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:train_kernel -s 1 -c 1 -f -o $O/${R}_scaled python tools/probe.py --shape scaled --dim 200 --random --epochs 1 --test 10 > $O/${R}_ncu_scaled.log 2>&1
# summaries are made here (the reports are too large to travel back together: gpurun_out/ is capped at 64 MiB)
python tools/ncu_summary.py $O/${R}_ncu_full_summary.txt train=$O/${R}_train.ncu-rep rank=$O/${R}_rank.ncu-rep sweep=$O/${R}_sweep.ncu-rep project=$O/${R}_project.ncu-rep rankf32=$O/${R}_rankf32.ncu-rep transr=$O/${R}_transr.ncu-rep transh=$O/${R}_transh.ncu-rep scaled=$O/${R}_scaled.ncu-rep > /dev/null 2>&1
cp profiles/traffic.json $O/${R}_traffic.json
python tools/launch_list_summary.py $O/${R}_launches_all.csv > $O/${R}_launch_list_summary.txt 2>&1
head -c 400000 $O/${R}_launches_all.csv > $O/${R}_launches.csv
rm -f $O/${R}_transr.ncu-rep $O/${R}_transh.ncu-rep $O/${R}_scaled.ncu-rep $O/${R}_rankf32.ncu-rep $O/${R}_launches_all.csv $O/${R}_project.ncu-rep $O/${R}_sweep.ncu-rep
[ $(du -sm $O | cut -f1) -gt 55 ] && rm -f $O/${R}_train.ncu-rep $O/${R}_rank.ncu-rep
{
  echo "# config 0: TransE unif L1 size=50, FB15k shape"; python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 20 --test 59071 2>&1 | grep -E "epochs|rank"
  echo "# config 1: TransE bern L2 size=100, FB15k shape"; python tools/probe.py --model transe --dim 100 --distance 1 --epochs 20 --test 59071 2>&1 | grep -E "epochs|rank"
  echo "# config 2: TransH bern size=100, WN18 shape"; python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 5000 2>&1 | grep -E "epochs|rank"
  echo "# config 3: TransR size=50 L1, FB15k shape"; python tools/probe.py --model transr --dim 50 --distance 0 --epochs 20 --test 59071 2>&1 | grep -E "epochs|rank"
  echo "# config 4 on ONE GPU: TransE L2 size=200, scaled shape, random triples"; python tools/probe.py --shape scaled --dim 200 --random --epochs 2 --test 10 2>&1 | grep -E "epochs"
  echo "# TransE L1 size=100, WN18 shape"; python tools/probe.py --model transe --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs"
  echo "# batched training, config 1 shape"; python tools/probe_sweep.py --models 1,4,8,16 2>&1
} > $O/${R}_config_probes.txt 2>&1
KB2E_TRAIN_TRACE=$O/${R}_trace.txt python tools/probe.py --model transe --dim 100 --distance 1 --epochs 10 --test 10 > /dev/null 2>&1
python tools/trace_report.py $O/${R}_trace.txt 5 > $O/${R}_trace_report.txt 2>/dev/null
KB2E_TRAIN_TRACE=$O/${R}_trace_sweep.txt python tools/probe_sweep.py --models 8 --epochs 4 > /dev/null 2>&1
python tools/trace_report.py $O/${R}_trace_sweep.txt 5 > $O/${R}_trace_sweep_report.txt 2>/dev/null
KB2E_TRAIN_TRACE=$O/${R}_trace_transr.txt python tools/probe.py --model transr --dim 50 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1
python tools/trace_report.py $O/${R}_trace_transr.txt 6 > $O/${R}_trace_transr_report.txt 2>/dev/null
rm -f $O/${R}_trace_transr.txt
KB2E_TRANSR_STATS=1 python tools/probe.py --model transr --dim 50 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|transRNorm" >> $O/${R}_trace_transr_report.txt
KB2E_TRAIN_TRACE=$O/${R}_trace_transh.txt python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1
python tools/trace_report.py $O/${R}_trace_transh.txt 5 > $O/${R}_trace_transh_sr_report.txt 2>/dev/null
KB2E_TRAIN_TRACE=$O/${R}_trace_transh.txt KB2E_TRAIN_TRACE_FINE=1 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1
python tools/trace_transh_sr_fine.py $O/${R}_trace_transh.txt >> $O/${R}_trace_transh_sr_report.txt 2>/dev/null
KB2E_TRANSH_SR=0 KB2E_TRAIN_TRACE=$O/${R}_trace_transh.txt python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1
python tools/trace_report.py $O/${R}_trace_transh.txt 5 > $O/${R}_trace_transh_report.txt 2>/dev/null
rm -f $O/${R}_trace_transh.txt
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${R}_pytest_gpu.txt
cat $O/${R}_pytest_gpu.txt; cut -c1-600 $O/${R}_bench.json; cat $O/${R}_config_probes.txt; tail -14 $O/${R}_trace_sweep_report.txt
