"""ORACLE (test infrastructure): import the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by the CPU
tests that are skipped when the reference tree is absent.

The reference package imports solver/plot libraries at module top
(/root/reference/gnngls/__init__.py:1-6) and ``dgl`` (models.py:1, datasets.py:5)
which are not installed; empty stub modules make the import succeed.  ``dgl.nn``
gets ``GATConv = oracle.model_port.GATConvPort`` and
``utils.Sequential = torch.nn.Sequential`` so models.py runs unmodified.
"""
import importlib
import itertools
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get('GNNGLS_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'gnngls', 'operators.py'))


_cache = {}


def load():
    """Return the reference ``gnngls`` package (with .operators/.algorithms/.models)."""
    if 'pkg' in _cache:
        return _cache['pkg']
    if not available():
        raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
    import torch.nn as nn
    from . import model_port

    def stub(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    for name in ('concorde', 'lkh', 'tsplib95', 'matplotlib'):
        if name not in sys.modules:
            stub(name)
    if 'concorde.tsp' not in sys.modules:
        sys.modules['concorde'].tsp = stub('concorde.tsp')
    if 'matplotlib.colors' not in sys.modules:
        sys.modules['matplotlib'].colors = stub('matplotlib.colors')
    if 'dgl' not in sys.modules:
        dgl = stub('dgl')
        dgl.nn = stub('dgl.nn')
        dgl.nn.utils = stub('dgl.nn.utils')
        dgl.nn.GATConv = model_port.GATConvPort
        dgl.nn.utils.Sequential = nn.Sequential
    # import under a private name so it never shadows anything called ``gnngls``
    spec = importlib.util.spec_from_file_location(
        'gnngls_reference', os.path.join(REFERENCE_ROOT, 'gnngls', '__init__.py'),
        submodule_search_locations=[os.path.join(REFERENCE_ROOT, 'gnngls')])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules['gnngls_reference'] = pkg
    spec.loader.exec_module(pkg)
    pkg.operators = importlib.import_module('gnngls_reference.operators')
    pkg.algorithms = importlib.import_module('gnngls_reference.algorithms')
    pkg.models = importlib.import_module('gnngls_reference.models')
    _cache['pkg'] = pkg
    return pkg


def make_graph(D, extra=None):
    """networkx K_n built like scripts/generate_instances.py:27-33 from a dense matrix."""
    import networkx as nx
    n = D.shape[0]
    G = nx.Graph()
    for v in range(n):
        G.add_node(v)
    for i, j in itertools.combinations(range(n), 2):
        attrs = {'weight': float(D[i, j])}
        for k, M in (extra or {}).items():
            attrs[k] = M[i, j]
        G.add_edge(i, j, **attrs)
    return G


class _FakeClock:
    def __init__(self):
        self.ticks = 0

    def time(self):
        return float(self.ticks)


def gls_fixed_iters(G, init_tour, init_cost, n_iters, weight='weight', guides=('weight',),
                    perturbation_moves=30, first_improvement=False):
    """Run the unmodified ``guided_local_search`` (algorithms.py:135-195) for exactly
    ``n_iters`` outer iterations: ``algorithms.time`` is swapped for a clock that
    advances by one per ``local_search`` call, and t_lim = n_iters + 0.5."""
    ref = load()
    alg = ref.algorithms
    clock = _FakeClock()
    real_time, real_ls = alg.time, alg.local_search

    def counted_ls(*a, **k):
        out = real_ls(*a, **k)
        clock.ticks += 1
        return out

    alg.time, alg.local_search = clock, counted_ls
    try:
        best_tour, best_cost, progress = alg.guided_local_search(
            G, list(init_tour), init_cost, n_iters + 0.5, weight=weight, guides=list(guides),
            perturbation_moves=perturbation_moves, first_improvement=first_improvement)
    finally:
        alg.time, alg.local_search = real_time, real_ls
    return best_tour, best_cost, [float(p['cost']) for p in progress]
