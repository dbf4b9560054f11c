#!/bin/bash
# A/B several prebuilt variants of libgnngls_b200.so on ONE GPU box, alternating runs so that clock / power-cap drift
# cancels:   tools/ab_bench.sh "libV0.so libV1.so ..." [rounds] [bench args...]
LIBS=$1; ROUNDS=${2:-2}; shift 2
ARGS=${*:---instances-per-gpu 4096 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e}
cp gnngls_b200/_lib/libgnngls_b200.so /tmp/lib_orig.so
for r in $(seq $ROUNDS); do
  for l in $LIBS; do
    cp $l gnngls_b200/_lib/libgnngls_b200.so
    echo -n "$l  "
    timeout 120 python bench.py $ARGS 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; print('value %.0f  gat %.1f  ff %.1f  fc %.1f  gls %.1f  mhz %.0f' % (d['value'], s.get('gat_kn',0), s['ff'], s['fc'], s['gls'], d['clocks']['sm_mhz']))"
  done
done
cp /tmp/lib_orig.so gnngls_b200/_lib/libgnngls_b200.so
