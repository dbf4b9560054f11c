"""gnngls_b200 — B200-native drop-in for the inference hot path of proroklab/gnngls.

    import gnngls_b200 as gnngls
    from gnngls_b200 import models, operators, algorithms

mirrors ``gnngls``, ``gnngls.models``, ``gnngls.operators`` and ``gnngls.algorithms`` of the
reference for the path *edge-regret prediction -> nearest-neighbour init -> guided local search*.
All compute runs in hand-written sm_100a CUDA behind ``libgnngls_b200.so`` (include/gnngls_b200.h);
there is no CPU fallback.
"""
import numpy as np

__version__ = '0.1.0'


def edge_matrix(G, attr, dtype=np.float64):
    """Dense symmetric [n,n] matrix of an edge attribute of a networkx graph whose nodes are
    0..n-1 (what ``nx.attr_matrix(G, attr)`` returns at /root/reference/gnngls/algorithms.py:140)."""
    n = G.number_of_nodes()
    M = np.zeros((n, n), dtype=dtype)
    for u, v, w in G.edges(data=attr, default=0):
        M[u, v] = w
        M[v, u] = w
    return M


def tour_cost(G, tour, weight='weight'):
    """/root/reference/gnngls/__init__.py:17-21 — sequential fp64 sum along the tour, on device."""
    import torch
    from . import _ops
    D = torch.from_numpy(edge_matrix(G, weight)).cuda()[None]
    t = torch.tensor([list(tour)], dtype=torch.int32, device='cuda')
    return float(_ops.tour_cost(D, t)[0])


def tour_to_edge_attribute(G, tour):
    """__init__.py:9-14 (host-side bookkeeping, not on the compute path)."""
    on = set()
    for a, b in zip(tour[:-1], tour[1:]):
        on.add((a, b)); on.add((b, a))
    return {e: (e in on) for e in G.edges}


def is_equivalent_tour(tour_a, tour_b):
    """__init__.py:24-29."""
    return tour_a == tour_b or tour_a == tour_b[::-1]


def is_valid_tour(G, tour):
    """__init__.py:32-44."""
    if tour[0] != 0 or tour[-1] != 0:
        return False
    counts = {}
    for v in tour:
        counts[v] = counts.get(v, 0) + 1
    return all(counts.get(v, 0) == (2 if v == 0 else 1) for v in G.nodes)


def optimal_cost(G, weight='weight'):
    """__init__.py:55-60 (needs the 'in_solution' labels of the reference's datasets)."""
    c = 0
    for e in G.edges:
        if G.edges[e]['in_solution']:
            c += G.edges[e][weight]
    return c
