#!/bin/bash
# quick: correctness on the risky sizes, repeated; then timing; optional ncu (arg2=ncu)
set -u
O=gpurun_out; T=${1:-r2g}
mkdir -p $O
for c in "3 3" "4 5" "20 40" "50 9" "100 3" "101 4" "127 3" "128 3" "128 40 0" "100 256 0 20" "127 3" "128 5" "100 30 0"; do timeout 120 python tools/kn_case.py $c 2>&1 | tail -1; done
timeout 120 python tools/kn_bench.py 100 256 20
timeout 120 python tools/kn_bench.py 50 1024 20
timeout 120 python tools/kn_bench.py 20 4096 20
if [ "${2:-}" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_kn_tc -s 2 -c 1 -f -o $O/${T}_kn python tools/kn_bench.py 100 256 2 > $O/${T}_ncu.log 2>&1; echo ncu rc=$?
fi
