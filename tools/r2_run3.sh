#!/bin/bash
set -u
O=gpurun_out; T=${1:-r2d}
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $O/${T}_gpu.txt
timeout 600 python tools/kn_check.py > $O/${T}_kn_full.log 2>&1; echo "kn full rc=$?"
grep -c " ok" $O/${T}_kn_full.log; grep -v " ok" $O/${T}_kn_full.log | head -20; tail -4 $O/${T}_kn_full.log
timeout 120 python tools/kn_bench.py 100 256 20
timeout 120 python tools/kn_bench.py 50 1024 20
timeout 120 python tools/kn_bench.py 20 4096 20
GNNGLS_KN_IMPL=scan timeout 120 python tools/kn_bench.py 100 256 20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_kn_tc -s 2 -c 1 -f -o $O/${T}_kn python tools/kn_bench.py 100 256 2 > $O/${T}_ncu.log 2>&1; echo ncu rc=$?
