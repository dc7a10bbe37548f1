#!/bin/bash
# the `ncu --set full` captures of tools/profile_round.sh on their own (the reports stay on the box: only the summary travels back)
R=${1:-r02}
O=gpurun_out
mkdir -p $O
set -x
ncu --set full --clock-control none --import-source on -k regex:train_kernel -s 3 -c 1 -f -o $O/${R}_train python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-partitioned > $O/${R}_ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rank_l2_tc -s 1 -c 1 -f -o $O/${R}_rank python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-partitioned > $O/${R}_ncu_rank.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:train_sweep -s 2 -c 1 -f -o $O/${R}_sweep python tools/probe_sweep.py --models 8 --epochs 2 > $O/${R}_ncu_sweep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:project_tc -s 1 -c 1 -f -o $O/${R}_project python tools/probe.py --model transr --dim 50 --distance 0 --epochs 2 --test 59071 > $O/${R}_ncu_project.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rank_f32 -s 1 -c 1 -f -o $O/${R}_rankf32 python tools/probe.py --model transr --dim 50 --distance 0 --epochs 2 --test 59071 > $O/${R}_ncu_rankf32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:train_transr -s 3 -c 1 -f -o $O/${R}_transr python tools/probe.py --model transr --dim 50 --distance 0 --epochs 6 --test 10 > $O/${R}_ncu_transr.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:train_transh_sr|train_kernel" -s 3 -c 1 -f -o $O/${R}_transh python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 6 --test 10 > $O/${R}_ncu_transh.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:train_kernel -s 1 -c 1 -f -o $O/${R}_scaled python tools/probe.py --shape scaled --dim 200 --random --epochs 1 --test 10 > $O/${R}_ncu_scaled.log 2>&1
python tools/ncu_summary.py $O/${R}_ncu_full_summary.txt train=$O/${R}_train.ncu-rep rank=$O/${R}_rank.ncu-rep sweep=$O/${R}_sweep.ncu-rep project=$O/${R}_project.ncu-rep rankf32=$O/${R}_rankf32.ncu-rep transr=$O/${R}_transr.ncu-rep transh=$O/${R}_transh.ncu-rep scaled=$O/${R}_scaled.ncu-rep > /dev/null 2>&1
cp profiles/traffic.json $O/${R}_traffic.json
rm -f $O/${R}_*.ncu-rep
grep -E '^##|gpu__time_duration|dram traffic' $O/${R}_ncu_full_summary.txt
