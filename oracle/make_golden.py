"""ORACLE (test infrastructure): generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden
Everything written here is an OUTPUT OF THE REFERENCE'S OWN PYTHON (operators.py, algorithms.py,
__init__.py:tour_cost, models.py) on seeded inputs that are stored alongside.
"""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, model_port          # noqa: E402
from gnngls_b200 import instances                # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def random_tour(rng, n):
    return [0] + (rng.permutation(n - 1) + 1).tolist() + [0]


def matrix_kinds(rng, n):
    _, D = instances.random_instances(1, n, seed=int(rng.integers(1 << 30)))
    yield 'euclid', D[0]
    M = rng.integers(1, 6, size=(n, n)).astype(np.float64)
    M = np.triu(M, 1); M = M + M.T
    yield 'int_ties', M
    A = rng.random((n, n)); np.fill_diagonal(A, 0.0)
    yield 'asymmetric', A
    # regular polygon: many exactly/nearly equal deltas -> exercises np.isclose(0, delta)
    ang = 2 * np.pi * np.arange(n) / n
    P = np.stack([np.cos(ang), np.sin(ang)], 1)[None]
    yield 'polygon', instances.distance_matrices(P)[0]
    # tiny perturbations around a constant: deltas ~1e-9..1e-7 straddle the isclose threshold
    T = 1.0 + rng.integers(-40, 41, size=(n, n)).astype(np.float64) * 2.5e-9
    T = np.triu(T, 1); T = T + T.T
    yield 'near_zero', T


def gen_operators(ref):
    ops = ref.operators
    rng = np.random.default_rng(7)
    out = {}
    k = 0
    for n in (3, 4, 5, 6, 7, 8, 10, 13, 20, 33, 50):
        for kind, D in matrix_kinds(rng, n):
            for rep in range(2 if n <= 20 else 1):
                tour = random_tour(rng, n)
                for fi in (False, True):
                    pre = f'c{k}_'
                    out[pre + 'D'] = D
                    out[pre + 'tour'] = np.array(tour, dtype=np.int32)
                    out[pre + 'fi'] = np.array(int(fi))
                    out[pre + 'kind'] = np.array(kind)
                    d, t = ops.two_opt_a2a(tour, D, fi)
                    out[pre + 'two_opt_a2a_delta'] = np.float64(d); out[pre + 'two_opt_a2a_tour'] = np.array(t, np.int32)
                    d, t = ops.relocate_a2a(tour, D, fi)
                    out[pre + 'relocate_a2a_delta'] = np.float64(d); out[pre + 'relocate_a2a_tour'] = np.array(t, np.int32)
                    if n >= 3:
                        idx = list(range(1, n)) if n <= 8 else sorted(set([1, n - 1] + rng.integers(1, n, 4).tolist()))
                        d2, t2, d3, t3 = [], [], [], []
                        for i in idx:
                            d, t = ops.two_opt_o2a(tour, D, i, fi); d2.append(d); t2.append(t)
                            d, t = ops.relocate_o2a(tour, D, i, fi); d3.append(d); t3.append(t)
                        out[pre + 'o2a_i'] = np.array(idx, np.int32)
                        out[pre + 'two_opt_o2a_delta'] = np.array(d2, np.float64)
                        out[pre + 'two_opt_o2a_tour'] = np.array(t2, np.int32)
                        out[pre + 'relocate_o2a_delta'] = np.array(d3, np.float64)
                        out[pre + 'relocate_o2a_tour'] = np.array(t3, np.int32)
                    k += 1
    out['n_cases'] = np.array(k)
    return out


def synthetic_regret(rng, D):
    """fp32-valued, clamped-at-zero guide with many exact zeros (like test.py:83 output)."""
    n = D.shape[0]
    r = (rng.random((n, n)) - 0.45).astype(np.float32)
    r = np.triu(r, 1); r = r + r.T
    return np.maximum(r.astype(np.float64), 0.0)


def gen_search(ref):
    alg = ref.algorithms
    rng = np.random.default_rng(11)
    out = {}
    k = 0
    cfgs = [  # n, K, moves, guide list, first_improvement
        (5, 2, 5, ('weight',), False), (6, 3, 30, ('regret_pred',), False), (8, 4, 20, ('regret_pred',), False),
        (8, 2, 20, ('weight',), True), (12, 5, 20, ('weight', 'regret_pred'), False),
        (20, 5, 20, ('regret_pred',), False), (20, 5, 30, ('weight',), False), (20, 3, 20, ('regret_pred',), True),
        (20, 0, 20, ('regret_pred',), False), (35, 4, 20, ('regret_pred',), False),
        (50, 4, 20, ('regret_pred',), False), (50, 3, 20, ('weight',), False),
        (100, 3, 20, ('regret_pred',), False), (100, 2, 20, ('weight',), False),
    ]
    for n, K, pm, guides, fi in cfgs:
        for rep in range(3 if n <= 20 else 1):
            P, D = instances.random_instances(1, n, seed=int(rng.integers(1 << 30)))
            P, D = P[0], D[0]
            R = synthetic_regret(rng, D)
            G = ref_shim.make_graph(D, {'regret_pred': R})
            init = alg.nearest_neighbor(G, 0, weight=guides[0])
            init_cost = ref.tour_cost(G, init)
            Dm, _ = __import__('networkx').attr_matrix(G, 'weight')
            assert np.array_equal(Dm, D)
            ls_tour, ls_cost, ls_prog = alg.local_search(list(init), init_cost, D, fi)
            bt, bc, prog = ref_shim.gls_fixed_iters(G, init, init_cost, K, guides=guides, perturbation_moves=pm,
                                                    first_improvement=fi)
            pen, _ = __import__('networkx').attr_matrix(G, 'penalty')
            pre = f'c{k}_'
            out[pre + 'P'] = P; out[pre + 'regret'] = R
            out[pre + 'cfg'] = np.array([n, K, pm, int(fi)], np.int32)
            out[pre + 'guides'] = np.array(guides)
            out[pre + 'nn_tour'] = np.array(init, np.int32); out[pre + 'init_cost'] = np.float64(init_cost)
            out[pre + 'ls_tour'] = np.array(ls_tour, np.int32); out[pre + 'ls_cost'] = np.float64(ls_cost)
            out[pre + 'ls_events'] = np.array([p['cost'] for p in ls_prog], np.float64)
            out[pre + 'best_tour'] = np.array(bt, np.int32); out[pre + 'best_cost'] = np.float64(bc)
            out[pre + 'events'] = np.array(prog, np.float64)
            out[pre + 'penalty'] = np.asarray(pen, np.float64)
            k += 1
    out['n_cases'] = np.array(k)
    return out


def state_digest(sd):
    h = hashlib.sha256()
    for key in sorted(sd):
        h.update(key.encode()); h.update(sd[key].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def gen_model(ref):
    out = {}
    torch.manual_seed(0)
    model = ref.models.EdgePropertyPredictionModel(1, 128, 1, 3, n_heads=8)   # shipped params.json shape
    model_port.randomize_bn_stats(model, seed=1)
    model.eval()
    sd = model.state_dict()
    out['keys'] = np.array(list(sd.keys()))
    out['shapes'] = np.array([str(tuple(v.shape)) for v in sd.values()])
    out['digest'] = np.array(state_digest(sd))
    rng = np.random.default_rng(3)
    k = 0
    for n, B in ((5, 1), (8, 2), (20, 2), (30, 1)):
        _, D = instances.random_instances(B, n, seed=int(rng.integers(1 << 30)))
        x = torch.from_numpy(instances.edge_features(D) / np.float32(np.sqrt(2.0))).reshape(-1, 1)
        g = model_port.EdgeListGraph.kn_line_graph(n, batch=B)
        with torch.no_grad():
            y32 = model(g, x).numpy()
            model.double()
            y64 = model(g, x.double()).numpy()
            model.float()
        pre = f'c{k}_'
        out[pre + 'nB'] = np.array([n, B], np.int32)
        out[pre + 'x'] = x.numpy(); out[pre + 'y32'] = y32; out[pre + 'y64'] = y64
        k += 1
    out['n_cases'] = np.array(k)
    return out


def main():
    ref = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, 'operators.npz'), **gen_operators(ref))
    np.savez_compressed(os.path.join(OUT, 'search.npz'), **gen_search(ref))
    np.savez_compressed(os.path.join(OUT, 'model.npz'), **gen_model(ref))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
