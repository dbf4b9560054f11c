#!/bin/bash
# Round-2 evidence pass (run through gpurun; outputs land in gpurun_out/TAG_*): launch list, ncu captures of the dominant kernels,
# TSP500 captures, compute-sanitizer summaries, the config-5 move sweep and the secondary-configuration bench lines.
set -u
T=${1:-r2z}
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${T}_launches.csv \
    $B --global-instances 512 --steps 1 --warmup 1 > $O/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gat_kn_tc|ff_fused|gemm_tf32|gls_kernel' \
    --launch-skip 12 --launch-count 7 -f -o $O/${T}_prof $B --global-instances 512 --steps 1 --warmup 1 > $O/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_kn_tc -s 2 -c 1 -f -o $O/${T}_kn \
    python tools/kn_bench.py 100 256 2 > $O/${T}_ncu_kn.log 2>&1; echo "ncu kn rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:'gat_kn_scan|gls_kernel' --launch-skip 8 --launch-count 2 -f -o $O/${T}_n500 \
    $B --n 500 --global-instances 8 --steps 1 --warmup 1 --micro-batch 4 > $O/${T}_ncu_n500.log 2>&1; echo "ncu n500 rc=$?"
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/kn_case.py 20 4 > $O/${T}_san_${tool}_kn.log 2>&1; echo "$tool kn rc=$?"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/gls_diff.py 30 4 2 > $O/${T}_san_${tool}_gls.log 2>&1; echo "$tool gls rc=$?"
done
timeout 900 python tools/moves_sweep.py > $O/${T}_sweep.jsonl 2> $O/${T}_sweep.err; echo "sweep rc=$?"
$B --n 20 --global-instances 100 --steps 20 --warmup 5 > $O/${T}_bench_tsp20.log 2>&1; echo "tsp20 rc=$?"
$B --n 50 --global-instances 10000 --steps 5 --warmup 3 > $O/${T}_bench_tsp50.log 2>&1; echo "tsp50 rc=$?"
$B --n 500 --global-instances 8 --steps 3 --warmup 2 --micro-batch 4 > $O/${T}_bench_tsp500.log 2>&1; echo "tsp500 rc=$?"
GNNGLS_OP_DTYPE=tf32 GNNGLS_FT_DTYPE=tf32 $B --global-instances 12500 --steps 3 --warmup 2 > $O/${T}_bench_tf32.log 2>&1; echo "tf32 rc=$?"
for f in tsp20 tsp50 tsp500 tf32; do tail -1 $O/${T}_bench_$f.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', round(d['value'],1), d['stage_ms_per_step'])"; done
grep -h "ERROR SUMMARY\|RACECHECK SUMMARY" $O/${T}_san_*.log
