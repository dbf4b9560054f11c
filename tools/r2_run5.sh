#!/bin/bash
# full check of the round-2 state: kernel check, GPU test-suite, default bench, n=20/50 benches
set -u
O=gpurun_out; T=${1:-r2n}
mkdir -p $O
timeout 600 python tools/kn_check.py > $O/${T}_kn_full.log 2>&1; echo "kn full rc=$?"; tail -2 $O/${T}_kn_full.log
timeout 120 python tools/kn_bench.py 100 256 20
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -5 $O/${T}_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/${T}_bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/${T}_bench.log | cut -c1-1500
