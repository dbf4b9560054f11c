"""Drop-in for /root/reference/gnngls/operators.py, evaluated on the GPU.

``two_opt_a2a / relocate_a2a / two_opt_o2a / relocate_o2a`` keep the reference's arguments and
return values ``(delta, new_tour)``; the candidate scan runs in the fp64 shared-memory move
evaluator (csrc/search.cu) and picks exactly the move the reference's sequential scan picks.
``*_batch`` variants take device tensors ``tours [B,n+1] int32`` and ``D [B,n,n]`` (or one shared
``[n,n]``) fp64.
"""
import numpy as np
import torch

from . import _ops


# ---- scalar helpers of the reference API (list manipulation / one fp64 expression) -------------
def two_opt(tour, i, j):
    """operators.py:6-11."""
    if i == j:
        return tour
    if j < i:
        i, j = j, i
    return tour[:i] + tour[i:j][::-1] + tour[j:]


def two_opt_cost(tour, D, i, j):
    """operators.py:14-29."""
    if i == j:
        return 0
    if j < i:
        i, j = j, i
    a, b, c, d = tour[i], tour[i - 1], tour[j], tour[j - 1]
    return D[a, c] + D[b, d] - D[a, b] - D[c, d]


def relocate(tour, i, j):
    """operators.py:76-80."""
    new_tour = list(tour)
    node = new_tour.pop(i)
    new_tour.insert(j, node)
    return new_tour


def relocate_cost(tour, D, i, j):
    """operators.py:83-103."""
    if i == j:
        return 0
    a, b, c = tour[i - 1], tour[i], tour[i + 1]
    d, e = (tour[j], tour[j + 1]) if i < j else (tour[j - 1], tour[j])
    return -D[a, b] - D[b, c] + D[a, c] - D[d, e] + D[d, b] + D[b, e]


# ---- batched device API -------------------------------------------------------------------------
def two_opt_a2a_batch(tours, D, first_improvement=False):
    return _ops.moves_eval(_ops.OP_TWO_OPT, D, tours, None, first_improvement)


def relocate_a2a_batch(tours, D, first_improvement=False):
    return _ops.moves_eval(_ops.OP_RELOCATE, D, tours, None, first_improvement)


def two_opt_o2a_batch(tours, D, pos, first_improvement=False):
    return _ops.moves_eval(_ops.OP_TWO_OPT, D, tours, pos, first_improvement)


def relocate_o2a_batch(tours, D, pos, first_improvement=False):
    return _ops.moves_eval(_ops.OP_RELOCATE, D, tours, pos, first_improvement)


# ---- reference signatures ------------------------------------------------------------------------
def _as_device_matrix(D):
    if isinstance(D, torch.Tensor):
        return D.to(device='cuda', dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(D, dtype=np.float64))).cuda()


def _single(op, tour, D, i, first_improvement):
    Dd = _as_device_matrix(D)
    t = torch.tensor([list(tour)], dtype=torch.int32, device='cuda')
    pos = None
    if i is not None:
        assert i > 0 and i < len(tour) - 1          # operators.py:54,107
        pos = torch.tensor([i], dtype=torch.int32, device='cuda')
    delta, move, new = _ops.moves_eval(op, Dd, t, pos, first_improvement)
    if int(move[0, 0]) < 0:
        return 0, tour                                # operators.py:50,73,126,147
    return float(delta[0]), new[0].tolist()


def two_opt_a2a(tour, D, first_improvement=False):
    """operators.py:32-50."""
    return _single(_ops.OP_TWO_OPT, tour, D, None, first_improvement)


def two_opt_o2a(tour, D, i, first_improvement=False):
    """operators.py:53-73."""
    return _single(_ops.OP_TWO_OPT, tour, D, i, first_improvement)


def relocate_o2a(tour, D, i, first_improvement=False):
    """operators.py:106-126."""
    return _single(_ops.OP_RELOCATE, tour, D, i, first_improvement)


def relocate_a2a(tour, D, first_improvement=False):
    """operators.py:129-147."""
    return _single(_ops.OP_RELOCATE, tour, D, None, first_improvement)
