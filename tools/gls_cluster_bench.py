"""A/B of the search kernels with one CTA per instance (GNNGLS_CLUSTER=0) against the cluster tier, CUDA-event timed.

    python tools/gls_cluster_bench.py [n B K]...      e.g.  500 1 10  500 8 10  200 8 10  100 8 10
Prints one JSON line per (n, B, K, cluster setting): NN + local_search + GLS times and the a2a sweep time."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnngls_b200 import _ops, algorithms, instances  # noqa: E402


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    return best, out


def main():
    args = [int(x) for x in sys.argv[1:]] or [500, 1, 10, 500, 8, 10]
    settings = os.environ.get('SETTINGS', '0,auto').split(',')
    for q in range(0, len(args), 3):
        n, B, K = args[q:q + 3]
        rng = np.random.default_rng(n + B)
        _, D = instances.random_instances(B, n, seed=n)
        N = n * (n - 1) // 2
        regret = np.maximum(rng.random((B, N)).astype(np.float32) - np.float32(0.4), 0)
        Dd = torch.as_tensor(D).cuda()
        rd = torch.as_tensor(regret).cuda()
        ref = None
        for setting in settings:
            if setting == 'auto':
                os.environ.pop('GNNGLS_CLUSTER', None)
            else:
                os.environ['GNNGLS_CLUSTER'] = setting
            tours, costs = algorithms.nearest_neighbor_batch(rd, Dd)
            ms_sweep, _ = timed(lambda: _ops.moves_eval(_ops.OP_TWO_OPT, Dd, tours, None, False), reps=5)
            ms_ls, ls = timed(lambda: algorithms.local_search_batch(tours, costs, Dd))
            ms_gls, gl = timed(lambda: algorithms.guided_local_search_batch(Dd, rd.view(B, 1, N), tours, costs, K,
                                                                            perturbation_moves=20, keep_penalties=True))
            res = (gl[0].cpu().numpy(), gl[1].cpu().numpy())
            same = True
            if ref is None:
                ref = res
            else:
                same = bool(np.array_equal(ref[0], res[0]) and np.array_equal(ref[1].view(np.int64), res[1].view(np.int64)))
            print(json.dumps(dict(n=n, B=B, K=K, cluster=setting, two_opt_sweep_ms=round(ms_sweep, 4),
                                  local_search_ms=round(ms_ls, 3), gls_ms=round(ms_gls, 3), same_as_first=same,
                                  mean_best_cost=float(res[1].mean()))), flush=True)


if __name__ == '__main__':
    main()
