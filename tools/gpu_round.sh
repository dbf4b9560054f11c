#!/bin/bash
# One GPU-box pass that produces everything committed under profiles/ and quoted in DESIGN.md:
#   tools/gpu_round.sh TAG     (run through gpurun; outputs land in gpurun_out/TAG_*)
# 1. GPU parity tests + smoke, 2. default bench (both arms), 3. ncu launch list of a short bench run,
# 4. one `ncu --set full` capture of the three dominant kernels (launches after warm-up).
set -u
T=${1:-rX}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/${T}_gpu_tests.log 2>&1; echo "tests rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > $O/${T}_bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/${T}_bench_ref.log 2>&1; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${T}_launches.csv \
    python bench.py --instances-per-gpu 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gat_kn_star|ff_fused|gemm_tf32|gls_kernel' \
    --launch-skip 12 --launch-count 7 -f -o $O/${T}_prof \
    python bench.py --instances-per-gpu 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -1 $O/${T}_bench.log | cut -c1-600
