#!/bin/bash
# Round-3 evidence pass (run through gpurun; outputs land in gpurun_out/TAG_*): GPU tests, smoke, default bench + reference arm, the
# secondary configurations, ncu captures of the cluster-tier search kernels at n = 500, compute-sanitizer on the cluster tier, the
# config-5 move sweep on the current kernels.
set -u
T=${1:-r3z}
O=gpurun_out
mkdir -p $O
B="python bench.py --no-cpu-baseline --no-e2e"
timeout 900 python -m pytest tests -x -q -m gpu --timeout 120 2>&1 | tail -5 > $O/${T}_gpu_tests.log; echo "tests rc=$?"; tail -2 $O/${T}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > $O/${T}_bench.log 2> $O/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/${T}_bench_ref.log 2>&1; echo "bench ref rc=$?"
python bench.py --no-cpu-baseline --n 500 --global-instances 8 --steps 5 --warmup 3 > $O/${T}_bench_tsp500x8.log 2>&1; echo "tsp500x8 rc=$?"
python bench.py --no-cpu-baseline --n 500 --global-instances 1 --steps 5 --warmup 3 > $O/${T}_bench_tsp500x1.log 2>&1; echo "tsp500x1 rc=$?"
GNNGLS_CLUSTER=0 python bench.py --no-cpu-baseline --n 500 --global-instances 8 --steps 3 --warmup 2 > $O/${T}_bench_tsp500x8_solo.log 2>&1; echo "tsp500x8 solo rc=$?"
GNNGLS_CLUSTER=0 python bench.py --no-cpu-baseline --n 500 --global-instances 1 --steps 3 --warmup 2 > $O/${T}_bench_tsp500x1_solo.log 2>&1; echo "tsp500x1 solo rc=$?"
$B --n 20 --global-instances 100 --steps 20 --warmup 5 > $O/${T}_bench_tsp20.log 2>&1; echo "tsp20 rc=$?"
$B --n 50 --global-instances 10000 --steps 5 --warmup 3 > $O/${T}_bench_tsp50.log 2>&1; echo "tsp50 rc=$?"
SETTINGS=0,auto timeout 200 python tools/gls_cluster_bench.py 500 1 10 500 8 10 1000 1 5 200 8 10 100 8 10 > $O/${T}_cluster_bench.log 2>&1; echo "cluster bench rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gls_cluster_kernel|local_search_cluster_kernel|nn_init_block' -c 3 -f -o $O/${T}_cluster \
    env SETTINGS=auto python tools/gls_cluster_bench.py 500 1 10 > $O/${T}_ncu_cluster.log 2>&1; echo "ncu cluster rc=$?"
for tool in memcheck racecheck; do
  GNNGLS_ROWCACHE=1 timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/gls_diff.py 64 2 2 > $O/${T}_san_${tool}_cluster.log 2>&1; echo "$tool cluster rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${T}_launches.csv \
    $B --global-instances 512 --steps 1 --warmup 1 > $O/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gat_kn_tc|ff_fused|gemm_tf32|gls_kernel' \
    --launch-skip 12 --launch-count 7 -f -o $O/${T}_prof $B --global-instances 512 --steps 1 --warmup 1 > $O/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gat_kn_tc -s 2 -c 1 -f -o $O/${T}_kn \
    python tools/kn_bench.py 100 256 2 > $O/${T}_ncu_kn.log 2>&1; echo "ncu kn rc=$?"
timeout 900 python tools/moves_sweep.py > $O/${T}_sweep.jsonl 2> $O/${T}_sweep.err; echo "sweep rc=$?"
