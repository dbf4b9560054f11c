#!/usr/bin/env python
"""CPU baseline of BASELINE.md 3.4, measured in the BUILD container (the only place /root/reference exists): the UNMODIFIED
reference Python -- gnngls/models.py forward (on the restated GATConv, DGL being absent), nearest_neighbor, tour_cost and
guided_local_search (gnngls/algorithms.py:135-195) with K fixed outer iterations -- per TSP100 instance, on ONE core and with all
torch threads, beside the C port of the same search.  Writes profiles/r2_reference_python_cpu.json.

    python tools/ref_python_baseline.py [instances]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnngls_b200 import instances          # noqa: E402
from oracle import gls_port, model_port, ref_shim   # noqa: E402  (diagnostic script, not product code)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n, K, pm = 100, 10, 20
    ref = ref_shim.load()
    gls_port.build()
    torch.manual_seed(0)
    model = ref.models.EdgePropertyPredictionModel(1, 128, 1, 3, n_heads=8).eval()      # the reference's own class
    _, D = instances.random_instances(B, n)
    g = model_port.EdgeListGraph.kn_line_graph(n, 1)
    N = n * (n - 1) // 2
    out = {'n': n, 'K': K, 'perturbation_moves': pm, 'instances': B, 'host_cores': os.cpu_count()}
    for threads in (1, os.cpu_count() or 1):
        torch.set_num_threads(threads)
        t_model = t_search = t_port = 0.0
        costs, costs_port = [], []
        for b in range(B):
            x = (instances.edge_features(D[b:b + 1])[0] / np.sqrt(2.0)).astype(np.float32)
            t0 = time.perf_counter()
            with torch.no_grad():
                y = model(g, torch.from_numpy(x).reshape(-1, 1)).numpy().reshape(-1)
            t_model += time.perf_counter() - t0
            lo, hi = np.quantile(y, [0.05, 0.95])
            regret = np.maximum((y - lo) / max(hi - lo, 1e-6) * 0.3, 0).astype(np.float32)   # synthetic scaler (random-init weights)
            W = gls_port.regret_matrix(regret, n)
            G = ref_shim.make_graph(D[b], {'regret_pred': W})
            t0 = time.perf_counter()
            init = ref.algorithms.nearest_neighbor(G, 0, weight='regret_pred')                 # test.py:85
            c0 = ref.tour_cost(G, init)                                                        # test.py:90
            _, c, _ = ref_shim.gls_fixed_iters(G, init, c0, K, guides=('regret_pred',), perturbation_moves=pm)
            t_search += time.perf_counter() - t0
            costs.append(c)
            t0 = time.perf_counter()
            _, cp = gls_port.pipeline_batch(D[b:b + 1], regret[None], K, pm, nthreads=1)
            t_port += time.perf_counter() - t0
            costs_port.append(float(cp[0]))
        out[f'threads_{threads}'] = {
            'reference_python_model_s_per_instance': t_model / B, 'reference_python_search_s_per_instance': t_search / B,
            'reference_python_instances_per_s': B / (t_model + t_search),
            'c_port_search_s_per_instance': t_port / B, 'tours_costs_identical_to_c_port': costs == costs_port}
        print(threads, out[f'threads_{threads}'], flush=True)
    json.dump(out, open(os.path.join(os.path.dirname(__file__), '..', 'profiles', 'r2_reference_python_cpu.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
