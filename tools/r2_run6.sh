#!/bin/bash
# GPU test-suite + default bench (100k instances) + reference arm
set -u
O=gpurun_out; T=${1:-r2r}
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -s > $O/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 $O/${T}_gpu_tests.log; grep -E "TSP[0-9]+ x|fp16 range|above fp16|below fp16" $O/${T}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${T}_smoke.log
( time timeout 1200 python bench.py > $O/${T}_bench.log 2> $O/${T}_bench.err ) 2> $O/${T}_bench.time; echo "bench rc=$?"; tail -1 $O/${T}_bench.log | cut -c1-600; cat $O/${T}_bench.time; tail -3 $O/${T}_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 $O/${T}_bench_ref.log | cut -c1-400
