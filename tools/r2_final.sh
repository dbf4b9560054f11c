#!/bin/bash
# final pass of the round: GPU test-suite, smoke, default bench (both arms), ncu capture of the final K_n kernel, secondary configs
set -u
O=gpurun_out; T=${1:-r2final}
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -2 $O/${T}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${T}_smoke.log
timeout 1200 python bench.py > $O/${T}_bench.log 2> $O/${T}_bench.err; echo "bench rc=$?"; tail -1 $O/${T}_bench.log | cut -c1-300
timeout 900 python bench.py --impl reference > $O/${T}_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 $O/${T}_bench_ref.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_kn_tc -s 2 -c 1 -f -o $O/${T}_kn python tools/kn_bench.py 100 256 2 > $O/${T}_ncu_kn.log 2>&1; echo "ncu kn rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${T}_launches.csv python bench.py --no-cpu-baseline --no-e2e --global-instances 512 --steps 1 --warmup 1 > $O/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
B="python bench.py --no-cpu-baseline --no-e2e"
$B --n 20 --global-instances 100 --steps 20 --warmup 5 > $O/${T}_bench_tsp20.log 2>&1
$B --n 50 --global-instances 10000 --steps 5 --warmup 3 > $O/${T}_bench_tsp50.log 2>&1
for f in tsp20 tsp50; do tail -1 $O/${T}_bench_$f.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', round(d['value'],1), d['stage_ms_per_step'])"; done
