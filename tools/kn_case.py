#!/usr/bin/env python
"""One kn_check case with wall time: python tools/kn_case.py n B [hot] [reps]"""
import sys, time
sys.path.insert(0, __file__.rsplit('/', 2)[0])
from tools import kn_check
n, B = int(sys.argv[1]), int(sys.argv[2])
hot = (sys.argv[3] != '0') if len(sys.argv) > 3 else True
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 0
t0 = time.time()
try:
    ok = kn_check.case(n, B, True, hot=hot, reps=reps)
except Exception as e:
    print('EXC', type(e).__name__, str(e)[:80]); ok = False
print(f'n={n} B={B} wall {time.time() - t0:.1f}s ok={ok}', flush=True)
