set -x
O=gpurun_out; R=r01
python tools/probe.py --model transe --dim 100 --distance 1 --epochs 20 --test 1000 2>&1 | grep epochs
python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 20 --test 1000 2>&1 | grep epochs
KB2E_TRAIN_TRACE=$O/${R}_trace.txt python tools/probe.py --model transe --dim 100 --distance 1 --epochs 10 --test 10 > /dev/null 2>&1
python tools/trace_report.py $O/${R}_trace.txt 5 2>/dev/null | head -12
python -m pytest tests/test_gpu_train.py tests/test_gpu_rank.py -m gpu -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches_all.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${R}_ncu_launch.log 2>&1
python tools/launch_list_summary.py $O/${R}_launches_all.csv > $O/${R}_launch_list_summary.txt 2>&1
grep -E "train_kernel|rank_|filter_|recheck|etrue|prep_|finalize|build_queries|transpose_kernel|widen|narrow|segment_|hash_insert|pack_triples|init_rows|count_chunks|RadixSort|^\"ID\"" $O/${R}_launches_all.csv > $O/${R}_launches.csv
rm -f $O/${R}_launches_all.csv
cat $O/${R}_launch_list_summary.txt
