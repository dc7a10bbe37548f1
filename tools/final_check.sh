# final check of the in-tree build on the GPU box: smoke + the GPU test suite (+ optional TransR ranking probe)
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -q 2>&1 | tail -2
[ "$1" = "transr" ] && python tools/probe.py --model transr --dim 50 --distance 0 --epochs 5 --test 5000 2>&1 | grep rank
