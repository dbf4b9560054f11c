#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_gpu.txt
timeout 300 python tools/kn_check.py --quick > $O/r2a_kn_quick.log 2>&1; echo "kn quick rc=$?"
tail -4 $O/r2a_kn_quick.log
timeout 600 python tools/kn_check.py > $O/r2a_kn_full.log 2>&1; echo "kn full rc=$?"
tail -12 $O/r2a_kn_full.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -5 $O/r2a_gpu_tests.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2a_bench.log 2>&1; echo "bench rc=$?"
tail -1 $O/r2a_bench.log | cut -c1-300
GNNGLS_KN_WARPS_PER_HEAD=4 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2a_bench_w4.log 2>&1; echo "bench w4 rc=$?"
tail -1 $O/r2a_bench_w4.log | cut -c1-300
