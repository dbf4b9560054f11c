#!/usr/bin/env python
"""Where does the GPU guided_local_search leave the oracle's trajectory?  python tools/gls_diff.py n [B] [K]"""
import sys
import numpy as np
import torch
sys.path.insert(0, __file__.rsplit('/', 2)[0])
from gnngls_b200 import algorithms, instances  # noqa: E402
from oracle import gls_port  # noqa: E402

n = int(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else 2; K = int(sys.argv[3]) if len(sys.argv) > 3 else 1
gls_port.build()
rng = np.random.default_rng(n)
_, D = instances.random_instances(B, n, seed=7 * n)
N = n * (n - 1) // 2
regret = np.maximum(rng.random((B, N)).astype(np.float32) - np.float32(0.4), 0).astype(np.float32)
Dd, rd = torch.from_numpy(D).cuda(), torch.from_numpy(regret).cuda()
tours, costs = algorithms.nearest_neighbor_batch(rd, Dd)
bt, bc, info = algorithms.guided_local_search_batch(Dd, rd.view(B, 1, N), tours, costs, K, perturbation_moves=20, max_events=1 << 16,
                                                    keep_penalties=True)
print('status', info['status'].tolist(), 'n_events', info['n_events'].tolist())
for b in range(B):
    W = gls_port.regret_matrix(regret[b], n)
    nn = gls_port.nearest_neighbor(W)
    c0 = gls_port.tour_cost(D[b], nn)
    t, c, ev, pen = gls_port.guided_local_search(D[b], W[None], nn, c0, K, 20, return_penalties=True)
    k = int(info['n_events'][b])
    g = info['events'][b, :k].cpu().numpy()
    ev = np.asarray(ev)
    m = min(len(ev), k)
    neq = np.nonzero(g[:m] != ev[:m])[0]
    print(f'b={b}: oracle events {len(ev)} gpu events {k}; first mismatch at {neq[0] if len(neq) else None}; best cost gpu {float(bc[b]):.12f} oracle {c:.12f}')
    if len(neq):
        i = neq[0]
        print('   oracle', ev[max(0, i - 2):i + 3], '\n   gpu   ', g[max(0, i - 2):i + 3])
    gp = info['state'].penalties[b].cpu().numpy()
    print('   penalties equal:', np.array_equal(gp.astype(np.float64), pen), 'sum', gp.sum(), pen.sum())
# the exact calls of tests/test_search_gpu.py::test_nn_ls_gls_batch_vs_oracle
for me in (8192, 1 << 16):
    ls_t, ls_c, ls_info = algorithms.local_search_batch(tours, costs, Dd, max_events=512)
    bt2, bc2, info2 = algorithms.guided_local_search_batch(Dd, rd.view(B, 1, N), tours, costs, K, perturbation_moves=20, max_events=me,
                                                           keep_penalties=True)
    o_t, o_c, o_cnt = gls_port.pipeline_batch(D, regret, K, 20, nthreads=4, want_counters=True)
    print('max_events', me, 'status', info2['status'].tolist(), 'tours equal', [bt2[b].tolist() == o_t[b].tolist() for b in range(B)],
          'costs', [float(bc2[b]) == o_c[b] for b in range(B)], 'ls status', ls_info['status'].tolist(), 'ls n_events', ls_info['n_events'].tolist())
