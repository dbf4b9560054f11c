"""ORACLE (test infrastructure): ctypes binding of oracle/gls_port.c.

Python-facing helpers mirror the reference's signatures (operators.py, algorithms.py) so
tests read like calls into the reference.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libgls_port.so')
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_fp = ctypes.POINTER(ctypes.c_float)
_lp = ctypes.POINTER(ctypes.c_long)


def build(force=False):
    src = os.path.join(_HERE, 'gls_port.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.glsp_tour_cost.restype = ctypes.c_double
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _prep(tour, D):
    D = np.ascontiguousarray(D, dtype=np.float64)
    t = np.ascontiguousarray(tour, dtype=np.int32)
    return t, D, D.shape[0]


def tour_cost(D, tour):
    t, D, n = _prep(tour, D)
    return float(lib().glsp_tour_cost(_d(D), n, _i(t)))


def _move(fn, tour, D, i, first_improvement):
    t, D, n = _prep(tour, D)
    assert len(t) == n + 1
    out = np.empty(n + 1, dtype=np.int32)
    delta = ctypes.c_double()
    mi, mj = ctypes.c_int(), ctypes.c_int()
    if i is None:
        found = fn(_i(t), _d(D), n, int(first_improvement), ctypes.byref(delta), ctypes.byref(mi),
                   ctypes.byref(mj), _i(out))
    else:
        found = fn(_i(t), _d(D), n, int(i), int(first_improvement), ctypes.byref(delta), ctypes.byref(mi),
                   ctypes.byref(mj), _i(out))
    if found < 0:
        raise AssertionError('i out of range')
    return (delta.value if found else 0.0), out.tolist(), (mi.value, mj.value) if found else None


def two_opt_a2a(tour, D, first_improvement=False):
    return _move(lib().glsp_two_opt_a2a, tour, D, None, first_improvement)


def relocate_a2a(tour, D, first_improvement=False):
    return _move(lib().glsp_relocate_a2a, tour, D, None, first_improvement)


def two_opt_o2a(tour, D, i, first_improvement=False):
    return _move(lib().glsp_two_opt_o2a, tour, D, i, first_improvement)


def relocate_o2a(tour, D, i, first_improvement=False):
    return _move(lib().glsp_relocate_o2a, tour, D, i, first_improvement)


def local_search(init_tour, init_cost, D, first_improvement=False, max_events=4096):
    t, D, n = _prep(init_tour, D)
    t = t.copy()
    cost = ctypes.c_double(float(init_cost))
    ev = np.zeros(max_events, dtype=np.float64)
    nev = ctypes.c_int()
    lib().glsp_local_search(_i(t), ctypes.byref(cost), _d(D), n, int(first_improvement), _d(ev),
                            ctypes.byref(nev), max_events)
    return t.tolist(), cost.value, ev[:min(nev.value, max_events)].tolist()


def nearest_neighbor(W, depot=0):
    W = np.ascontiguousarray(W, dtype=np.float64)
    n = W.shape[0]
    t = np.empty(n + 1, dtype=np.int32)
    lib().glsp_nearest_neighbor(_d(W), n, int(depot), _i(t))
    return t.tolist()


def guided_local_search(D, guides, init_tour, init_cost, n_iters, perturbation_moves=30,
                        first_improvement=False, max_events=1 << 16, return_penalties=False):
    """guides: array [n_guides, n, n].  Returns (best_tour, best_cost, event_costs[, penalties])."""
    t, D, n = _prep(init_tour, D)
    guides = np.ascontiguousarray(guides, dtype=np.float64).reshape(-1, n, n)
    best = np.empty(n + 1, dtype=np.int32)
    bc = ctypes.c_double()
    ev = np.zeros(max_events, dtype=np.float64)
    nev = ctypes.c_int()
    pen = np.zeros((n, n), dtype=np.float64)
    cnt = np.zeros(3, dtype=np.int64)
    rc = lib().glsp_guided_local_search(_d(D), _d(guides), guides.shape[0], n, _i(t), ctypes.c_double(float(init_cost)),
                                        int(n_iters), int(perturbation_moves), int(first_improvement), _i(best),
                                        ctypes.byref(bc), _d(ev), ctypes.byref(nev), max_events, _d(pen),
                                        cnt.ctypes.data_as(_lp))
    if rc != 0:
        raise AssertionError('operator index out of range')
    out = (best.tolist(), bc.value, ev[:min(nev.value, max_events)].tolist())
    if return_penalties:
        out = out + (pen,)
    return out


def regret_matrix(regret_f32, n):
    r = np.ascontiguousarray(regret_f32, dtype=np.float32).reshape(-1)
    W = np.empty((n, n), dtype=np.float64)
    lib().glsp_regret_matrix(r.ctypes.data_as(_fp), n, _d(W))
    return W


def pipeline_batch(D, regret_f32, n_iters, perturbation_moves=20, nthreads=1, want_counters=False):
    """test.py:79-95 for a batch: D [B,n,n] fp64, regret_f32 [B,N] fp32 or None (guide = weight)."""
    D = np.ascontiguousarray(D, dtype=np.float64)
    B, n, _ = D.shape
    tours = np.empty((B, n + 1), dtype=np.int32)
    costs = np.empty(B, dtype=np.float64)
    cnt = np.zeros((B, 3), dtype=np.int64) if want_counters else None
    rp = None
    if regret_f32 is not None:
        regret_f32 = np.ascontiguousarray(regret_f32, dtype=np.float32).reshape(B, -1)
        rp = regret_f32.ctypes.data_as(_fp)
    lib().glsp_pipeline_batch(_d(D), rp, B, n, int(n_iters), int(perturbation_moves), int(nthreads), _i(tours),
                              _d(costs), cnt.ctypes.data_as(_lp) if want_counters else None)
    return (tours, costs, cnt) if want_counters else (tours, costs)
