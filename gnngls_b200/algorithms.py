"""Drop-in for the search half of /root/reference/gnngls/algorithms.py on B200.

``nearest_neighbor``, ``local_search`` and ``guided_local_search`` keep the reference's arguments
and return values; the work runs in the persistent per-instance kernels of csrc/search.cu and
yields bit-identical tours and costs.  The wall-clock limit ``t_lim`` is honoured exactly as in the
reference — it is tested between outer iterations (algorithms.py:146) — by running the device loop
one outer iteration per launch; ``n_iters=`` (extension) runs a fixed number of outer iterations in
a single launch instead.  ``*_batch`` functions are the throughput API on device tensors.

Out of scope (not on the inference path, SURVEY.md section 2 #3b): probabilistic_nearest_neighbour,
best_probabilistic_nearest_neighbour, cheapest_insertion, insertion.
"""
import time

import numpy as np
import torch

from . import _ops, edge_matrix

DEFAULT_MAX_EVENTS = 4096


# ---- batched device API -------------------------------------------------------------------------
def nearest_neighbor_batch(guide, D=None, depot=0):
    """guide: [B,n,n] fp64 matrix or [B,N] fp32 edge vector (line-graph node order).
    Returns (tours [B,n+1] int32, costs [B] fp64 under D or None)."""
    kind = _ops.GUIDE_MATRIX_F64 if guide.dtype == torch.float64 else _ops.GUIDE_EDGEVEC_F32
    return _ops.nn_init(guide, kind, D, depot)


def local_search_batch(init_tours, init_costs, D, first_improvement=False, max_events=0):
    """Returns (tours, costs, info) with info = dict(events, n_events, status, counters)."""
    tours, costs = init_tours.clone(), init_costs.clone()
    info = _ops.local_search(D, tours, costs, first_improvement, max_events, want_counters=True)
    return tours, costs, info


def guided_local_search_batch(D, guides, init_tours, init_costs, n_iters, perturbation_moves=30,
                              first_improvement=False, max_events=0, keep_penalties=False):
    """guides: [B,n_guides,n,n] fp64 or [B,n_guides,N] fp32.  Returns (best_tours, best_costs, info);
    info additionally carries the resumable ``state``."""
    kind = _ops.GUIDE_MATRIX_F64 if guides.dtype == torch.float64 else _ops.GUIDE_EDGEVEC_F32
    # a penalties buffer lets the kernel use its L2-resident and cluster tiers (csrc/search.cu), as pipeline.RegretGLS.solve does
    state = _ops.GlsState(D, guides, kind, init_tours, init_costs, keep_penalties=keep_penalties or D.shape[-1] >= 64)
    info = _ops.gls_run(state, n_iters, perturbation_moves, first_improvement, max_events, want_counters=True)
    info['state'] = state
    return state.best_tours, state.best_costs, info


# ---- reference signatures ------------------------------------------------------------------------
def _check_status(status):
    s = int(status.max()) if status.numel() else 0
    if s & _ops.INST_STALLED:
        raise RuntimeError('guided_local_search made no progress (the reference would loop forever on this input)')
    if s & _ops.INST_PENALTY_OVERFLOW:
        raise RuntimeError('penalty counter overflow')


def _progress(events, n_events, stamp):
    k = int(n_events)
    if k > events.shape[0]:
        raise RuntimeError(f'search_progress overflow: {k} events > max_events={events.shape[0]}')
    return [{'time': stamp, 'cost': c} for c in events[:k].tolist()]


def nearest_neighbor(G, depot, weight='weight'):
    """algorithms.py:9-18.  Ties go to the first neighbour in ``G.neighbors`` order, which for the
    reference's instances (edges added by itertools.combinations) is ascending node id."""
    n = G.number_of_nodes()
    for i in G.nodes:
        nb = list(G.neighbors(i))
        if nb != sorted(nb) or len(nb) != n - 1:
            raise NotImplementedError('nearest_neighbor expects a complete graph with ascending adjacency order')
    W = torch.from_numpy(edge_matrix(G, weight)).cuda()[None]
    tours, _ = _ops.nn_init(W, _ops.GUIDE_MATRIX_F64, None, depot)
    return tours[0].tolist()


def local_search(init_tour, init_cost, D, first_improvement=False):
    """algorithms.py:111-132 -> (cur_tour, cur_cost, search_progress)."""
    Dd = torch.as_tensor(np.asarray(D, dtype=np.float64)).cuda().contiguous()[None]
    tours = torch.tensor([list(init_tour)], dtype=torch.int32, device='cuda')
    costs = torch.tensor([float(init_cost)], dtype=torch.float64, device='cuda')
    info = _ops.local_search(Dd, tours, costs, first_improvement, DEFAULT_MAX_EVENTS)
    torch.cuda.synchronize()
    progress = _progress(info['events'][0], info['n_events'][0], time.time())
    if not progress:
        return init_tour, init_cost, progress
    return tours[0].tolist(), float(costs[0]), progress


def guided_local_search(G, init_tour, init_cost, t_lim, weight='weight', guides=['weight'], perturbation_moves=30,
                        first_improvement=False, n_iters=None):
    """algorithms.py:135-195 -> (best_tour, best_cost, search_progress).

    Like the reference it leaves the final ``'penalty'`` attribute on the edges of ``G``."""
    n = G.number_of_nodes()
    D = torch.from_numpy(edge_matrix(G, weight)).cuda()[None]
    gm = torch.stack([torch.from_numpy(edge_matrix(G, g)) for g in guides])[None].cuda().contiguous()
    tours = torch.tensor([list(init_tour)], dtype=torch.int32, device='cuda')
    costs = torch.tensor([float(init_cost)], dtype=torch.float64, device='cuda')
    state = _ops.GlsState(D, gm, _ops.GUIDE_MATRIX_F64, tours, costs, keep_penalties=True)
    progress = []

    def run(k):
        info = _ops.gls_run(state, k, perturbation_moves, first_improvement, DEFAULT_MAX_EVENTS)
        torch.cuda.synchronize()
        _check_status(info['status'])
        progress.extend(_progress(info['events'][0], info['n_events'][0], time.time()))

    if n_iters is not None:
        run(int(n_iters))
    else:
        run(0)                                   # :142 initial local search
        while time.time() < t_lim:               # :146
            run(1)
    pen = state.penalties[0].cpu().numpy()
    for u, v in G.edges:                          # :138,161 side effect on the caller's graph
        G.edges[u, v]['penalty'] = float(pen[u, v]) if pen[u, v] else 0
    return state.best_tours[0].tolist(), float(state.best_costs[0]), progress
