#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r02_s5_pytest.txt 2>&1
tail -6 $O/r02_s5_pytest.txt
{
  echo "# config 0: TransE unif L1 size=50, FB15k shape"; timeout 300 python tools/probe.py --model transe --dim 50 --distance 0 --method 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  echo "# config 1: TransE bern L2 size=100, FB15k shape"; timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  echo "# config 2: TransH bern size=100, WN18 shape"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  echo "# TransH bern size=100, FB15k shape"; timeout 300 python tools/probe.py --model transh --dim 100 --distance 0 --epochs 20 --test 100 2>&1 | grep -E "epochs|Error|error"
  echo "# TransE L1 size=100, WN18 shape"; timeout 300 python tools/probe.py --model transe --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|Error|error"
} > $O/r02_s5_probes.txt 2>&1
cat $O/r02_s5_probes.txt
KB2E_TRAIN_TRACE=$O/r02_s5_trace.txt timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 10 --test 10 > /dev/null 2>&1
python tools/trace_report.py $O/r02_s5_trace.txt 5 > $O/r02_s5_trace_report.txt 2>/dev/null
tail -12 $O/r02_s5_trace_report.txt
KB2E_TRAIN_TRACE_FINE=1 KB2E_TRAIN_TRACE=$O/r02_s5_trace_fine.txt timeout 300 python tools/probe.py --model transe --dim 100 --distance 1 --epochs 10 --test 10 > /dev/null 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r02_s5_bench.json 2> $O/r02_s5_bench.err || tail -5 $O/r02_s5_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s5_bench.json'))
print('train value %.1f M/s e2e %.1f M/s frac %.3f | eval value %.1f M q/s e2e cold %.1f resident %.1f' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['eval']['value']/1e6, d['eval']['e2e']['value']/1e6, d['eval']['e2e']['resident']['value']/1e6))
PY
