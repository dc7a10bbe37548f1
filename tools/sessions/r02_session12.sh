#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/r02_s12_pytest.txt 2>&1
tail -8 $O/r02_s12_pytest.txt
{
  echo "# config 2: TransH bern size=100, WN18 shape (train_transh_sr_kernel, rsqrt soft-constraint loop)"; timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 5000 2>&1 | grep -E "epochs|rank|rror"
  echo "# the same with KB2E_TRANSH_SR=0 (three-barrier list kernel)"; KB2E_TRANSH_SR=0 timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  echo "# TransH size=100 FB15k shape (1,345 relations: three-barrier list kernel)"; timeout 300 python tools/probe.py --model transh --dim 100 --distance 0 --epochs 20 --test 10 2>&1 | grep -E "epochs|rror"
  KB2E_TRAIN_TRACE=$O/r02_s12_trace_transh.txt timeout 300 python tools/probe.py --model transh --shape wn18 --dim 100 --distance 0 --epochs 6 --test 10 > /dev/null 2>&1
  python tools/trace_report.py $O/r02_s12_trace_transh.txt 5 > $O/r02_trace_transh_sr_report.txt 2>&1
  head -13 $O/r02_trace_transh_sr_report.txt
  python -c "import __graft_entry__ as g; g.smoke()"
} > $O/r02_s12_probes.txt 2>&1
cat $O/r02_s12_probes.txt
