"""Thin torch<->C-ABI call layer: validates tensors, passes raw device pointers and the current
CUDA stream to libgnngls_b200.so.  Every function here launches hand-written sm_100a kernels; there
is no CPU implementation behind any of them."""
import ctypes

import torch

from . import _lib

OP_TWO_OPT, OP_RELOCATE = 0, 1
GUIDE_MATRIX_F64, GUIDE_EDGEVEC_F32 = 0, 1
DENSE_TCGEN05, DENSE_SIMT, DENSE_TCGEN05_F16 = 0, 1, 2
FT_F32, FT_TF32, FT_F16 = 0, 1, 2       # gnngls_ft_dtype
INST_EVENTS_TRUNCATED, INST_PENALTY_OVERFLOW, INST_STALLED = 1, 2, 4


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _chk(t, dtype, name, shape=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f'{name} must be a CUDA tensor (gnngls_b200 has no CPU path)')
    if t.dtype != dtype:
        raise TypeError(f'{name} must be {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError(f'{name} must be contiguous')
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f'{name} must have shape {tuple(shape)}, got {tuple(t.shape)}')
    return t


def moves_eval(op, D, tours, pos=None, first_improvement=False, want_tours=True):
    """Batched operators.{two_opt,relocate}_{a2a,o2a}.  D: [B,n,n] or [n,n] fp64; tours [B,n+1] int32;
    pos [B] int32 selects the one-to-all variant.  Returns (delta [B] f64, move [B,2] i32, new_tours|None)."""
    lib = _lib.load()
    B, n1 = tours.shape
    n = n1 - 1
    _chk(tours, torch.int32, 'tours')
    if D.dim() == 2:
        _chk(D, torch.float64, 'D', (n, n)); stride = 0
    else:
        _chk(D, torch.float64, 'D', (B, n, n)); stride = n * n
    delta = torch.empty(B, dtype=torch.float64, device=tours.device)
    move = torch.empty(B, 2, dtype=torch.int32, device=tours.device)
    out = torch.empty_like(tours) if want_tours else None
    with torch.cuda.device(tours.device):
        if pos is None:
            rc = lib.gnngls_moves_eval_a2a(op, _ptr(D), stride, _ptr(tours), B, n, int(first_improvement), _ptr(delta),
                                           _ptr(move), _ptr(out), _stream())
        else:
            _chk(pos, torch.int32, 'pos', (B,))
            rc = lib.gnngls_moves_eval_o2a(op, _ptr(D), stride, _ptr(tours), _ptr(pos), B, n, int(first_improvement),
                                           _ptr(delta), _ptr(move), _ptr(out), _stream())
    _lib.check(rc)
    return delta, move, out


def local_search(D, tours, costs, first_improvement=False, max_events=0, want_counters=False):
    """In-place batched algorithms.local_search.  Returns dict(events, n_events, status, counters)."""
    lib = _lib.load()
    B, n1 = tours.shape
    n = n1 - 1
    _chk(tours, torch.int32, 'tours'); _chk(costs, torch.float64, 'costs', (B,)); _chk(D, torch.float64, 'D', (B, n, n))
    dev = tours.device
    events = torch.zeros(B, max_events, dtype=torch.float64, device=dev) if max_events > 0 else None
    n_events = torch.zeros(B, dtype=torch.int32, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    counters = torch.zeros(B, 4, dtype=torch.int64, device=dev) if want_counters else None
    with torch.cuda.device(dev):
        rc = lib.gnngls_local_search_batch(_ptr(D), _ptr(tours), _ptr(costs), B, n, int(first_improvement), _ptr(events),
                                           _ptr(n_events), max_events, _ptr(status), _ptr(counters), _stream())
    _lib.check(rc)
    return dict(events=events, n_events=n_events, status=status, counters=counters)


class GlsState:
    """Device-resident state of a batch of guided_local_search runs (resumable between calls)."""

    def __init__(self, D, guides, guide_kind, init_tours, init_costs, keep_penalties=True):
        B, n1 = init_tours.shape
        n = n1 - 1
        dev = init_tours.device
        self.B, self.n, self.device = B, n, dev
        self.D = _chk(D, torch.float64, 'D', (B, n, n))
        self.guide_kind = guide_kind
        if guide_kind == GUIDE_MATRIX_F64:
            _chk(guides, torch.float64, 'guides')
            if guides.dim() != 4 or guides.shape[0] != B or tuple(guides.shape[2:]) != (n, n):
                raise ValueError('guides must be [B, n_guides, n, n]')
        else:
            _chk(guides, torch.float32, 'guides')
            if guides.dim() != 3 or guides.shape[0] != B or guides.shape[2] != n * (n - 1) // 2:
                raise ValueError('guides must be [B, n_guides, n(n-1)/2]')
        self.guides, self.n_guides = guides, guides.shape[1]
        self.cur_tours = _chk(init_tours, torch.int32, 'init_tours').clone()
        self.cur_costs = _chk(init_costs, torch.float64, 'init_costs', (B,)).clone()
        self.best_tours = torch.empty_like(self.cur_tours)
        self.best_costs = torch.empty_like(self.cur_costs)
        self.k = torch.empty(B, dtype=torch.float64, device=dev)
        self.penalties = torch.zeros(B, n, n, dtype=torch.int32, device=dev) if keep_penalties else None
        self.iters_done = 0
        self.started = False


def gls_run(state, n_iters, perturbation_moves=30, first_improvement=False, max_events=0, want_counters=False):
    """Run `n_iters` more outer iterations of algorithms.guided_local_search on `state`
    (the first call also performs the initial local_search of algorithms.py:142)."""
    lib = _lib.load()
    dev = state.device
    B = state.B
    events = torch.zeros(B, max_events, dtype=torch.float64, device=dev) if max_events > 0 else None
    n_events = torch.zeros(B, dtype=torch.int32, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    counters = torch.zeros(B, 4, dtype=torch.int64, device=dev) if want_counters else None
    a = _lib.GlsArgs()
    a.B, a.n = B, state.n
    a.D = state.D.data_ptr()
    a.guide_kind, a.n_guides = state.guide_kind, state.n_guides
    a.guides = state.guides.data_ptr()
    a.cur_tours, a.cur_costs = state.cur_tours.data_ptr(), state.cur_costs.data_ptr()
    a.best_tours, a.best_costs = state.best_tours.data_ptr(), state.best_costs.data_ptr()
    a.k = state.k.data_ptr()
    a.penalties = state.penalties.data_ptr() if state.penalties is not None else None
    a.resume = 1 if state.started else 0
    if state.started and state.penalties is None:
        raise ValueError('resuming a GLS run needs keep_penalties=True')
    a.iter_begin, a.n_iters = state.iters_done, int(n_iters)
    a.perturbation_moves, a.first_improvement = int(perturbation_moves), int(first_improvement)
    a.events = events.data_ptr() if events is not None else None
    a.n_events = n_events.data_ptr()
    a.max_events = max_events
    a.status = status.data_ptr()
    a.counters = counters.data_ptr() if counters is not None else None
    with torch.cuda.device(dev):
        rc = lib.gnngls_gls_batch(ctypes.byref(a), _stream())
    _lib.check(rc)
    state.started = True
    state.iters_done += int(n_iters)
    return dict(events=events, n_events=n_events, status=status, counters=counters)


def nn_init(guide, guide_kind, D=None, depot=0):
    """Batched nearest_neighbor (+ tour_cost under D).  guide: [B,n,n] f64 or [B,N] f32."""
    lib = _lib.load()
    if guide_kind == GUIDE_MATRIX_F64:
        _chk(guide, torch.float64, 'guide')
        B, n = guide.shape[0], guide.shape[-1]
    else:
        _chk(guide, torch.float32, 'guide')
        if D is None:
            raise ValueError('edge-vector guides need D to infer n')
        B, n = guide.shape[0], D.shape[-1]
        if guide.numel() != B * n * (n - 1) // 2:
            raise ValueError('guide must be [B, n(n-1)/2]')
    if D is not None:
        _chk(D, torch.float64, 'D', (B, n, n))
    dev = guide.device
    tours = torch.empty(B, n + 1, dtype=torch.int32, device=dev)
    costs = torch.empty(B, dtype=torch.float64, device=dev) if D is not None else None
    with torch.cuda.device(dev):
        rc = lib.gnngls_nn_init_batch(guide_kind, _ptr(guide), _ptr(D), B, n, int(depot), _ptr(tours), _ptr(costs), _stream())
    _lib.check(rc)
    return tours, costs


def tour_cost(D, tours):
    lib = _lib.load()
    B, n1 = tours.shape
    n = n1 - 1
    _chk(D, torch.float64, 'D', (B, n, n)); _chk(tours, torch.int32, 'tours')
    out = torch.empty(B, dtype=torch.float64, device=tours.device)
    with torch.cuda.device(tours.device):
        rc = lib.gnngls_tour_cost_batch(_ptr(D), _ptr(tours), B, n, _ptr(out), _stream())
    _lib.check(rc)
    return out


def edge_features(D, scale=1.0, min_=0.0):
    lib = _lib.load()
    _chk(D, torch.float64, 'D')
    B, n = D.shape[0], D.shape[-1]
    x = torch.empty(B, n * (n - 1) // 2, dtype=torch.float32, device=D.device)
    with torch.cuda.device(D.device):
        rc = lib.gnngls_edge_features(_ptr(D), B, n, float(scale), float(min_), _ptr(x), _stream())
    _lib.check(rc)
    return x


def regret_postprocess(y, scale=1.0, min_=0.0, out=None):
    lib = _lib.load()
    _chk(y, torch.float32, 'y')
    out = torch.empty_like(y) if out is None else out
    with torch.cuda.device(y.device):
        rc = lib.gnngls_regret_postprocess(_ptr(y), y.numel(), float(scale), float(min_), _ptr(out), _stream())
    _lib.check(rc)
    return out
