"""Pins the CPU oracle (oracle/gls_port.c, oracle/model_port.py) against golden vectors that were
produced by the reference's own Python (oracle/make_golden.py).  Bit-exact for the search half."""
import numpy as np
import pytest
import torch

from oracle import gls_port, model_port
from gnngls_b200 import instances
from tests import _golden

OPS, _ = _golden.load('operators')
SEARCH, _ = _golden.load('search')
MODEL, MODEL_TOP = _golden.load('model')


def test_golden_present():
    assert len(OPS) > 100 and len(SEARCH) >= 14 and len(MODEL) >= 3


@pytest.mark.parametrize('name', ['two_opt_a2a', 'relocate_a2a'])
def test_a2a_operators_bit_exact(name):
    fn = getattr(gls_port, name)
    for c in OPS:
        delta, tour, _ = fn(c.tour, c.D, bool(c.fi))
        assert _golden.bits(delta) == _golden.bits(c[name + '_delta']), (name, str(c.kind))
        assert tour == c[name + '_tour'].tolist()


@pytest.mark.parametrize('name', ['two_opt_o2a', 'relocate_o2a'])
def test_o2a_operators_bit_exact(name):
    fn = getattr(gls_port, name)
    checked = 0
    for c in OPS:
        if 'o2a_i' not in c:
            continue
        for q, i in enumerate(c.o2a_i.tolist()):
            delta, tour, _ = fn(c.tour, c.D, i, bool(c.fi))
            assert _golden.bits(delta) == _golden.bits(c[name + '_delta'][q])
            assert tour == c[name + '_tour'][q].tolist()
            checked += 1
    assert checked > 300


def test_o2a_asserts_on_depot_index():
    c = OPS[10]
    with pytest.raises(AssertionError):
        gls_port.two_opt_o2a(c.tour, c.D, 0)
    with pytest.raises(AssertionError):
        gls_port.relocate_o2a(c.tour, c.D, len(c.tour) - 1)


def _inputs(c):
    D = instances.distance_matrices(c.P[None])[0]
    n, K, pm, fi = c.cfg.tolist()
    guides = [str(g) for g in c.guides]
    mats = np.stack([D if g == 'weight' else c.regret for g in guides])
    return D, mats, n, K, pm, bool(fi)


def test_nearest_neighbor_and_tour_cost():
    for c in SEARCH:
        D, mats, *_ = _inputs(c)
        assert gls_port.nearest_neighbor(mats[0]) == c.nn_tour.tolist()
        assert _golden.bits(gls_port.tour_cost(D, c.nn_tour)) == _golden.bits(c.init_cost)


def test_local_search_trajectory_bit_exact():
    for c in SEARCH:
        D, mats, n, K, pm, fi = _inputs(c)
        tour, cost, ev = gls_port.local_search(c.nn_tour, float(c.init_cost), D, fi)
        assert tour == c.ls_tour.tolist()
        assert _golden.bits(cost) == _golden.bits(c.ls_cost)
        assert np.array_equal(_golden.bits(ev), _golden.bits(c.ls_events))


def test_guided_local_search_trajectory_bit_exact():
    for c in SEARCH:
        D, mats, n, K, pm, fi = _inputs(c)
        bt, bc, ev, pen = gls_port.guided_local_search(D, mats, c.nn_tour, float(c.init_cost), K, pm, fi,
                                                       return_penalties=True)
        assert bt == c.best_tour.tolist(), c.cfg
        assert _golden.bits(bc) == _golden.bits(c.best_cost)
        assert np.array_equal(_golden.bits(ev), _golden.bits(c.events))
        assert np.array_equal(pen, c.penalty)


def test_pipeline_batch_matches_single_calls():
    cs = [c for c in SEARCH if c.cfg[0] == 20 and [str(g) for g in c.guides] == ['regret_pred'] and c.cfg[3] == 0
          and c.cfg[1] == 5]
    assert len(cs) == 3
    D = np.stack([instances.distance_matrices(c.P[None])[0] for c in cs])
    iu = np.triu_indices(20, 1)
    reg = np.stack([c.regret[iu].astype(np.float32) for c in cs])
    tours, costs = gls_port.pipeline_batch(D, reg, 5, 20, nthreads=2)
    for b, c in enumerate(cs):
        assert tours[b].tolist() == c.best_tour.tolist()
        assert _golden.bits(costs[b]) == _golden.bits(c.best_cost)


# ---------------------------------------------------------------- model half
def _port_model():
    torch.manual_seed(0)
    m = model_port.EdgeModelPort(1, 128, 1, 3, n_heads=8)
    model_port.randomize_bn_stats(m, seed=1)
    return m.eval()


def test_model_port_state_dict_matches_reference_layout():
    import hashlib
    m = _port_model()
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in MODEL_TOP['keys']]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in MODEL_TOP['shapes']]
    assert len(sd) == 140 and sum(v.numel() for k, v in sd.items() if 'num_batches' not in k and 'running' not in k) == 1191297
    h = hashlib.sha256()
    for key in sorted(sd):
        h.update(key.encode()); h.update(sd[key].contiguous().numpy().tobytes())
    assert h.hexdigest() == str(MODEL_TOP['digest'])


def test_model_port_reproduces_reference_outputs():
    m = _port_model()
    for c in MODEL:
        n, B = c.nB.tolist()
        g = model_port.EdgeListGraph.kn_line_graph(n, batch=B)
        with torch.no_grad():
            y32 = m(g, torch.from_numpy(c.x)).numpy()
            y64 = m.double()(g, torch.from_numpy(c.x).double()).numpy()
            m.float()
        assert np.allclose(y64, c.y64, rtol=0, atol=1e-12)
        assert np.allclose(y32, c.y32, rtol=0, atol=1e-4)      # fp32 summation-order noise only
        # fp32 evaluation error vs the fp64 yardstick: the basis for the GPU tolerance
        assert np.abs(y32 - c.y64).max() < 5e-5


def test_gatconv_two_independent_restatements_agree():
    torch.manual_seed(5)
    n = 7
    g = model_port.EdgeListGraph.kn_line_graph(n)
    N = g.number_of_nodes()
    conv = model_port.GATConvPort(128, 16, 8).double()
    h = torch.randn(N, 128, dtype=torch.float64)
    adj = torch.zeros(N, N, dtype=torch.bool)
    adj[g.dst, g.src] = True
    assert adj.sum(1).eq(2 * (n - 2)).all() and not adj.diagonal().any()
    with torch.no_grad():
        a = conv(g, h).reshape(N, -1)
        b = model_port.dense_gat_reference(h, conv.fc.weight, conv.attn_l, conv.attn_r, adj)
    assert torch.allclose(a, b, rtol=0, atol=1e-12)


def test_kn_line_graph_matches_networkx():
    import itertools
    import networkx as nx
    n = 6
    G = nx.complete_graph(n)
    lG = nx.line_graph(G)
    nodes = sorted(tuple(sorted(e)) for e in lG.nodes)
    rank = {e: i for i, e in enumerate(nodes)}
    ref = set()
    for a, b in lG.edges:
        u, v = rank[tuple(sorted(a))], rank[tuple(sorted(b))]
        ref.add((u, v)); ref.add((v, u))
    s, d = model_port.kn_line_graph_edges(n)
    assert set(zip(s.tolist(), d.tolist())) == ref
    assert [tuple(e) for e in model_port.kn_edge_list(n).tolist()] == nodes
    for (i, j) in itertools.combinations(range(n), 2):
        assert nodes[model_port.kn_rank(i, j, n)] == (i, j)
