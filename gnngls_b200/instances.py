"""Synthetic TSP instances (host side).

Distribution follows the reference's generator (/root/reference/scripts/generate_instances.py:27-33):
coordinates uniform in [0,1)^2, complete graph, Euclidean weights, node 0 is the depot.  Distances
are evaluated elementwise as sqrt(dx*dx + dy*dy) in fp64 (one IEEE rounding per operation), so any
host reproduces them bit-for-bit.
"""
import numpy as np

DEFAULT_SEED = 20211005


def distance_matrices(P):
    """P: [B,n,2] fp64 coordinates -> D: [B,n,n] fp64, symmetric, zero diagonal."""
    P = np.asarray(P, dtype=np.float64)
    dx = P[:, :, None, 0] - P[:, None, :, 0]
    dy = P[:, :, None, 1] - P[:, None, :, 1]
    return np.sqrt(dx * dx + dy * dy)


def random_instances(B, n, seed=DEFAULT_SEED):
    """Returns (P [B,n,2], D [B,n,n]) fp64."""
    rng = np.random.default_rng(seed)
    P = rng.random((B, n, 2))
    return P, distance_matrices(P)


def edge_features(D):
    """[B,n,n] -> [B,N] fp32: the reference's only node feature of the line graph, the edge weight
    cast to float32 (/root/reference/gnngls/datasets.py:14-20), in line-graph node order (i<j)."""
    n = D.shape[-1]
    iu = np.triu_indices(n, 1)
    return D[..., iu[0], iu[1]].astype(np.float32)
