"""Phase timers of the cluster tier's local_search (library built with GNNGLS_GLS_STAMPS=1):

    GNNGLS_GLS_STAMPS=1 python -m gnngls_b200.build --force;  python tools/gls_stamps.py [n]
Prints, per cluster member, the cycles thread 0 spent in each phase of the sweeps of one local_search from the nearest-neighbour tour."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnngls_b200 import algorithms, build, instances  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
rng = np.random.default_rng(n + 1)
_, D = instances.random_instances(1, n, seed=n)
N = n * (n - 1) // 2
regret = np.maximum(rng.random((1, N)).astype(np.float32) - np.float32(0.4), 0)
Dd, rd = torch.as_tensor(D).cuda(), torch.as_tensor(regret).cuda()
tours, costs = algorithms.nearest_neighbor_batch(rd, Dd)
for _ in range(2):
    t, c, info = algorithms.local_search_batch(tours, costs, Dd)
torch.cuda.synchronize()
cnt = info['counters'].cpu().numpy()[0]
sweeps = int(cnt[0] + cnt[1])
raw = ctypes.CDLL(build.LIB_PATH)
buf = (ctypes.c_ulonglong * 128)()
assert raw.gnngls_debug_gls_stamps(buf, 128) == 0, 'library was not built with GNNGLS_GLS_STAMPS=1'
a = np.array(buf[:], dtype=np.float64).reshape(16, 8)
names = ['E + first rows', 'rows', 'CTA reduce', 'exchange + cluster barrier', 'combine', 'apply + bookkeeping']
print(f'n={n}: {sweeps} sweeps; cycles per sweep (thread 0 of each member)')
print('member ' + ' '.join(f'{x[:14]:>15s}' for x in names) + '           total')
for r in range(16):
    if a[r].sum() == 0:
        continue
    print(f'{r:6d} ' + ' '.join(f'{a[r, k] / sweeps:15.0f}' for k in range(6)) + f' {a[r, :6].sum() / sweeps:15.0f}')
