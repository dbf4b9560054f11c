"""GPU parity (through the C ABI) of the search half against (1) the golden vectors produced by the
reference's own Python and (2) the CPU oracle (oracle/gls_port.c) on seeded random inputs.
Everything here is bit-exact: tours, chosen moves, fp64 deltas and costs."""
import numpy as np
import pytest
import torch

from gnngls_b200 import _ops, algorithms, instances, operators
from oracle import gls_port
from tests import _golden

pytestmark = pytest.mark.gpu

OPS, _ = _golden.load('operators')
SEARCH, _ = _golden.load('search')


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def random_tours(rng, B, n):
    t = np.zeros((B, n + 1), dtype=np.int32)
    for b in range(B):
        t[b, 1:n] = rng.permutation(n - 1) + 1
    return t


# ------------------------------------------------------------------ drop-in signatures vs golden
@pytest.mark.parametrize('name', ['two_opt_a2a', 'relocate_a2a'])
def test_a2a_dropin_matches_reference_golden(name):
    fn = getattr(operators, name)
    for c in OPS:
        tour = c.tour.tolist()
        delta, new = fn(tour, c.D, bool(c.fi))
        assert _golden.bits(delta) == _golden.bits(c[name + '_delta']), (name, str(c.kind), len(tour))
        assert new == c[name + '_tour'].tolist()
        if float(c[name + '_delta']) == 0.0:
            assert new is tour and delta == 0          # reference returns `0, tour`


@pytest.mark.parametrize('name', ['two_opt_o2a', 'relocate_o2a'])
def test_o2a_dropin_matches_reference_golden(name):
    fn = getattr(operators, name)
    for c in OPS[::3]:
        if 'o2a_i' not in c:
            continue
        for q, i in enumerate(c.o2a_i.tolist()):
            delta, new = fn(c.tour.tolist(), c.D, i, bool(c.fi))
            assert _golden.bits(delta) == _golden.bits(c[name + '_delta'][q])
            assert new == c[name + '_tour'][q].tolist()
    with pytest.raises(AssertionError):
        fn(OPS[20].tour.tolist(), OPS[20].D, 0)


def _search_inputs(c):
    D = instances.distance_matrices(c.P[None])[0]
    n, K, pm, fi = c.cfg.tolist()
    guides = [str(g) for g in c.guides]
    mats = np.stack([D if g == 'weight' else c.regret for g in guides])
    return D, mats, guides, n, K, pm, bool(fi)


def test_local_search_dropin_matches_reference_golden():
    for c in SEARCH:
        D, mats, guides, n, K, pm, fi = _search_inputs(c)
        tour, cost, prog = algorithms.local_search(c.nn_tour.tolist(), float(c.init_cost), D, fi)
        assert tour == c.ls_tour.tolist()
        assert _golden.bits(cost) == _golden.bits(c.ls_cost)
        assert np.array_equal(_golden.bits([p['cost'] for p in prog]), _golden.bits(c.ls_events))


def test_guided_local_search_dropin_matches_reference_golden():
    from oracle import ref_shim        # only make_graph (pure networkx); the reference itself is not needed
    for c in SEARCH:
        D, mats, guides, n, K, pm, fi = _search_inputs(c)
        G = ref_shim.make_graph(D, {'regret_pred': c.regret})
        assert algorithms.nearest_neighbor(G, 0, weight=guides[0]) == c.nn_tour.tolist()
        bt, bc, prog = algorithms.guided_local_search(G, c.nn_tour.tolist(), float(c.init_cost), None, weight='weight',
                                                      guides=guides, perturbation_moves=pm, first_improvement=fi, n_iters=K)
        assert bt == c.best_tour.tolist(), c.cfg
        assert _golden.bits(bc) == _golden.bits(c.best_cost)
        assert np.array_equal(_golden.bits([p['cost'] for p in prog]), _golden.bits(c.events))
        import networkx as nx
        pen, _ = nx.attr_matrix(G, 'penalty')
        assert np.array_equal(np.asarray(pen), c.penalty)      # side effect on G (algorithms.py:138,161)


def test_gls_chunked_resume_equals_single_launch():
    c = [c for c in SEARCH if c.cfg[0] == 50][0]
    D, mats, guides, n, K, pm, fi = _search_inputs(c)
    Dd, g = dev(D)[None], dev(mats)[None]
    t0, c0 = dev(c.nn_tour)[None], dev([float(c.init_cost)])
    st = _ops.GlsState(Dd, g, _ops.GUIDE_MATRIX_F64, t0, c0, keep_penalties=True)
    ev = []
    for k in [0] + [1] * K:           # the t_lim path: one outer iteration per launch
        info = _ops.gls_run(st, k, pm, fi, 4096)
        ev += info['events'][0, :int(info['n_events'][0])].tolist()
    assert st.best_tours[0].tolist() == c.best_tour.tolist()
    assert _golden.bits(float(st.best_costs[0])) == _golden.bits(c.best_cost)
    assert np.array_equal(_golden.bits(ev), _golden.bits(c.events))


def test_guided_local_search_time_limit_path():
    import time
    from oracle import ref_shim
    c = SEARCH[5]
    D, mats, guides, n, K, pm, fi = _search_inputs(c)
    G = ref_shim.make_graph(D, {'regret_pred': c.regret})
    t = time.time()
    bt, bc, prog = algorithms.guided_local_search(G, c.nn_tour.tolist(), float(c.init_cost), t + 0.3, guides=guides,
                                                  perturbation_moves=pm)
    assert time.time() - t >= 0.3 and bt[0] == 0 and bt[-1] == 0 and sorted(bt[:-1]) == list(range(n))
    assert bc <= float(c.ls_cost) and all('time' in p and 'cost' in p for p in prog)


# ------------------------------------------------------------------ batched kernels vs CPU oracle
@pytest.mark.parametrize('n,B', [(3, 4), (4, 8), (5, 16), (9, 64), (20, 256), (50, 128), (100, 64), (101, 16),
                                 (160, 8), (200, 6), (500, 2), (1000, 1)])
@pytest.mark.parametrize('fi', [False, True])
def test_moves_eval_batch_vs_oracle(n, B, fi):
    rng = np.random.default_rng(1000 * n + B + int(fi))
    kind = n % 3
    if kind == 0:
        D = rng.integers(1, 5, size=(B, n, n)).astype(np.float64)        # many ties
    elif kind == 1:
        D = rng.random((B, n, n))                                        # asymmetric
    else:
        _, D = instances.random_instances(B, n, seed=n)
    tours = random_tours(rng, B, n)
    pos = rng.integers(1, n, size=B).astype(np.int32)
    Dd, td, pd = dev(D), dev(tours), dev(pos)
    for op, a2a_ref, o2a_ref in ((_ops.OP_TWO_OPT, gls_port.two_opt_a2a, gls_port.two_opt_o2a),
                                 (_ops.OP_RELOCATE, gls_port.relocate_a2a, gls_port.relocate_o2a)):
        delta, move, new = _ops.moves_eval(op, Dd, td, None, fi)
        delta2, move2, new2 = _ops.moves_eval(op, Dd, td, pd, fi)
        delta, move, new = delta.cpu().numpy(), move.cpu().numpy(), new.cpu().numpy()
        delta2, move2, new2 = delta2.cpu().numpy(), move2.cpu().numpy(), new2.cpu().numpy()
        for b in range(B):
            d, t, mv = a2a_ref(tours[b], D[b], fi)
            assert _golden.bits(delta[b]) == _golden.bits(d), (n, b, op)
            assert new[b].tolist() == t and tuple(move[b]) == (mv if mv else (-1, -1))
            d, t, mv = o2a_ref(tours[b], D[b], int(pos[b]), fi)
            assert _golden.bits(delta2[b]) == _golden.bits(d), (n, b, op, 'o2a')
            assert new2[b].tolist() == t and tuple(move2[b]) == (mv if mv else (-1, -1))


def test_moves_eval_shared_matrix():
    rng = np.random.default_rng(5)
    n, B = 60, 40
    _, D = instances.random_instances(1, n, seed=3)
    tours = random_tours(rng, B, n)
    delta, move, new = _ops.moves_eval(_ops.OP_TWO_OPT, dev(D[0]), dev(tours), None, False)
    for b in range(B):
        d, t, _ = gls_port.two_opt_a2a(tours[b], D[0])
        assert _golden.bits(float(delta[b])) == _golden.bits(d) and new[b].tolist() == t


@pytest.mark.parametrize('n,B,K', [(4, 8, 3), (10, 32, 6), (20, 64, 10), (50, 32, 6), (100, 24, 4), (130, 4, 2),
                                   (170, 2, 1), (500, 2, 1)])
def test_nn_ls_gls_batch_vs_oracle(n, B, K):
    """Whole search pipeline (test.py:85-95) on a batch: NN init on an fp32 regret edge vector,
    tour_cost, local_search, GLS — bit-exact vs the CPU oracle; n=170 exercises the global-memory tier, n=500 is
    BASELINE.json's TSP500 configuration."""
    rng = np.random.default_rng(n)
    _, D = instances.random_instances(B, n, seed=7 * n)
    N = n * (n - 1) // 2
    regret = np.maximum(rng.random((B, N)).astype(np.float32) - np.float32(0.4), 0).astype(np.float32)
    Dd, rd = dev(D), dev(regret)
    tours, costs = algorithms.nearest_neighbor_batch(rd, Dd)
    ls_t, ls_c, ls_info = algorithms.local_search_batch(tours, costs, Dd, max_events=2048)
    bt, bc, info = algorithms.guided_local_search_batch(Dd, rd.view(B, 1, N), tours, costs, K, perturbation_moves=20,
                                                        max_events=8192, keep_penalties=True)
    assert int(info['status'].max()) == 0
    o_t, o_c, o_cnt = gls_port.pipeline_batch(D, regret, K, 20, nthreads=4, want_counters=True)
    for b in range(B):
        W = gls_port.regret_matrix(regret[b], n)
        nn = gls_port.nearest_neighbor(W)
        assert tours[b].tolist() == nn
        c0 = gls_port.tour_cost(D[b], nn)
        assert _golden.bits(float(costs[b])) == _golden.bits(c0)
        t1, c1, ev1 = gls_port.local_search(nn, c0, D[b])
        assert ls_t[b].tolist() == t1 and _golden.bits(float(ls_c[b])) == _golden.bits(c1)
        k = int(ls_info['n_events'][b])
        assert np.array_equal(_golden.bits(ls_info['events'][b, :k].cpu().numpy()), _golden.bits(ev1))
        assert bt[b].tolist() == o_t[b].tolist(), (n, b)
        assert _golden.bits(float(bc[b])) == _golden.bits(o_c[b])
        if b < 4:
            _, _, ev, pen = gls_port.guided_local_search(D[b], W[None], nn, c0, K, 20, return_penalties=True)
            k = int(info['n_events'][b])
            assert np.array_equal(_golden.bits(info['events'][b, :k].cpu().numpy()), _golden.bits(ev))
            assert np.array_equal(info['state'].penalties[b].cpu().numpy().astype(np.float64), pen)
    # counters: sweeps, o2a scans and accepted moves agree with the oracle's own counts
    cnt = info['counters'].cpu().numpy()
    assert np.array_equal(cnt[:, 0] + cnt[:, 1], o_cnt[:, 0]) and np.array_equal(cnt[:, 2], o_cnt[:, 1])


@pytest.mark.parametrize('n,B,K,csize', [(48, 3, 4, 2), (64, 5, 3, 8), (100, 3, 3, 16), (100, 40, 2, 4), (101, 7, 2, 16),
                                         (170, 2, 1, 2), (200, 9, 2, 16), (500, 1, 1, 16)])
@pytest.mark.parametrize('fi', [False, True])
def test_cluster_tier_bit_identical(n, B, K, csize, fi, monkeypatch):
    """One thread-block cluster per instance (north star (4)) against one CTA per instance and
    against the CPU oracle: sweeps, local_search and GLS -- tours, costs, event logs, penalties and counters bit-exact
    whatever the cluster size.  B=40 with clusters of 4 makes the clusters loop over instances."""
    rng = np.random.default_rng(31 * n + B)
    _, D = instances.random_instances(B, n, seed=11 * n + 1)
    if n == 101:
        D = np.round(D * 8.0) / 8.0                                   # many exact ties: the (delta, rank) order decides
    if n == 64:
        # (barely) asymmetric: relocate's column term cannot come from the row cache.  Only a few entries, by a few ulps: the
        # reference's delta formulas assume symmetry, so local_search need not terminate on a really asymmetric matrix
        for b in range(B):
            D[b, 3, 17] = np.nextafter(D[b, 3, 17], 2.0)
            D[b, 40, 41] = np.nextafter(np.nextafter(D[b, 40, 41], 0.0), 0.0)
    N = n * (n - 1) // 2
    regret = np.maximum(rng.random((B, N)).astype(np.float32) - np.float32(0.4), 0).astype(np.float32)
    Dd, rd = dev(D), dev(regret)
    rt = dev(random_tours(rng, B, n))

    def run():
        out = {}
        for op in (_ops.OP_TWO_OPT, _ops.OP_RELOCATE):
            d, m, t = _ops.moves_eval(op, Dd, rt, None, fi)
            out['mv%d' % op] = (d.cpu().numpy(), m.cpu().numpy(), t.cpu().numpy())
        tours, costs = algorithms.nearest_neighbor_batch(rd, Dd)
        out['nn'] = (tours.cpu().numpy(), costs.cpu().numpy())
        lt, lc, li = algorithms.local_search_batch(tours, costs, Dd, first_improvement=fi, max_events=4096)
        out['ls'] = (lt.cpu().numpy(), lc.cpu().numpy(), li['events'].cpu().numpy(), li['n_events'].cpu().numpy())
        bt, bc, info = algorithms.guided_local_search_batch(Dd, rd.view(B, 1, N), tours, costs, K, perturbation_moves=20,
                                                            first_improvement=fi, max_events=16384, keep_penalties=True)
        assert int(info['status'].max()) == 0
        out['gls'] = (bt.cpu().numpy(), bc.cpu().numpy(), info['events'].cpu().numpy(), info['n_events'].cpu().numpy(),
                      info['state'].penalties.cpu().numpy(), info['counters'].cpu().numpy())
        return out

    monkeypatch.setenv('GNNGLS_CLUSTER', '0')
    solo = run()
    monkeypatch.setenv('GNNGLS_CLUSTER', str(csize))
    for rows in ('0', '1'):                                           # without / with the per-warp row cache in shared memory
        monkeypatch.setenv('GNNGLS_ROWCACHE', rows)
        clus = run()
        for key in solo:
            for x, y in zip(solo[key], clus[key]):
                if x.dtype.kind == 'f':
                    assert np.array_equal(_golden.bits(x), _golden.bits(y)), (key, n, B, csize, rows)
                else:
                    assert np.array_equal(x, y), (key, n, B, csize, rows)
    if not fi:
        o_t, o_c = gls_port.pipeline_batch(D, regret, K, 20, nthreads=4)[:2]
        assert np.array_equal(clus['gls'][0], np.asarray(o_t))
        assert np.array_equal(_golden.bits(clus['gls'][1]), _golden.bits(np.asarray(o_c)))


def test_gls_matrix_guides_alternating_and_first_improvement():
    n, B, K = 30, 16, 5
    rng = np.random.default_rng(9)
    _, D = instances.random_instances(B, n, seed=99)
    R = np.stack([np.maximum(np.triu(rng.random((n, n)) - 0.5, 1), 0) for _ in range(B)])
    R = R + R.transpose(0, 2, 1)
    guides = np.stack([D, R], 1)
    Dd = dev(D)
    tours, costs = algorithms.nearest_neighbor_batch(Dd, Dd)
    for fi in (False, True):
        bt, bc, info = algorithms.guided_local_search_batch(Dd, dev(guides), tours, costs, K, 30, first_improvement=fi)
        for b in range(B):
            t, c, _ = gls_port.guided_local_search(D[b], guides[b], tours[b].tolist(), float(costs[b]), K, 30, fi)
            assert bt[b].tolist() == t and _golden.bits(float(bc[b])) == _golden.bits(c)


def test_full_size_properties_tsp100():
    """BASELINE config scale (n=100): size-independent properties on a larger batch."""
    n, B, K = 100, 512, 3
    _, D = instances.random_instances(B, n, seed=1)
    Dd = dev(D)
    tours, costs = algorithms.nearest_neighbor_batch(Dd, Dd)
    bt, bc, info = algorithms.guided_local_search_batch(Dd, Dd.view(B, 1, n, n), tours, costs, K, 20)
    bt_h = bt.cpu().numpy()
    assert (bt_h[:, 0] == 0).all() and (bt_h[:, -1] == 0).all()
    assert (np.sort(bt_h[:, :-1], axis=1) == np.arange(n)[None]).all()                # valid permutations
    assert (bc <= costs).all()                                                         # best never worse than init
    resum = _ops.tour_cost(Dd, bt)
    assert torch.allclose(resum, bc, rtol=1e-12, atol=0)                               # accumulated == re-summed cost
    again = algorithms.guided_local_search_batch(Dd, Dd.view(B, 1, n, n), tours, costs, K, 20)
    assert torch.equal(again[0], bt) and torch.equal(again[1], bc)                     # deterministic
    for b in (0, 17, 511):
        t, c, _ = gls_port.guided_local_search(D[b], D[b][None], tours[b].tolist(), float(costs[b]), K, 20)
        assert bt[b].tolist() == t and _golden.bits(float(bc[b])) == _golden.bits(c)


def test_error_reporting():
    with pytest.raises(RuntimeError, match='unsupported|range'):
        _ops.moves_eval(0, torch.zeros(1, 2, 2, dtype=torch.float64, device='cuda'),
                        torch.zeros(1, 3, dtype=torch.int32, device='cuda'))
