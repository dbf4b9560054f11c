#!/bin/bash
set -u
O=gpurun_out; T=${1:-r2c}
mkdir -p $O
timeout 600 python tools/kn_check.py > $O/${T}_kn_full.log 2>&1; echo "kn full rc=$?"
grep -c " ok" $O/${T}_kn_full.log; grep -v " ok" $O/${T}_kn_full.log | head; tail -4 $O/${T}_kn_full.log
python tools/kn_bench.py 100 256 20
GNNGLS_KN_WARPS_PER_HEAD=4 python tools/kn_bench.py 100 256 20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_kn_scan -s 2 -c 1 -f -o $O/${T}_kn python tools/kn_bench.py 100 256 2 > $O/${T}_ncu.log 2>&1; echo ncu rc=$?
