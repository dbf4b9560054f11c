import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from gnngls_b200 import algorithms, build, instances
n, B, K = 100, 2048, 10
rng = np.random.default_rng(1)
_, D = instances.random_instances(B, n, seed=3)
N = n * (n - 1) // 2
regret = np.maximum(rng.random((B, N)).astype(np.float32) - np.float32(0.4), 0)
Dd, rd = torch.as_tensor(D).cuda(), torch.as_tensor(regret).cuda()
tours, costs = algorithms.nearest_neighbor_batch(rd, Dd)
for _ in range(2):
    bt, bc, info = algorithms.guided_local_search_batch(Dd, rd.view(B, 1, N), tours, costs, K, perturbation_moves=20, keep_penalties=True)
torch.cuda.synchronize()
raw = ctypes.CDLL(build.LIB_PATH)
buf = (ctypes.c_ulonglong * 128)()
assert raw.gnngls_debug_gls_stamps(buf, 128) == 0
a = np.array(buf[:], dtype=np.float64).reshape(16, 8)
cnt = info['counters'].cpu().numpy()
print('first instance of CTAs 0..15: cycles in perturbation (slot 6) and local search (slot 7); counters of instance b:')
for r in range(16):
    ls = a[r, 5] + a[r, 7]
    print(r, 'perturb %.0f  ls %.0f  perturb share %.2f' % (a[r, 6], ls, a[r, 6] / (a[r, 6] + ls)), cnt[r].tolist())
