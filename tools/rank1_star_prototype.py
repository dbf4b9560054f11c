#!/usr/bin/env python
"""Prototype (CPU, fp64) of the sorted-threshold formulation of the K_n star aggregate.

For one star (vertex i of K_n) and one head, destination j aggregates the members k != j with weights
    w_jk = 2^(leaky_relu(el_k + er_j) - mx_j).
Because leaky_relu is piecewise linear, with the members sorted by el the set {k : el_k + er_j >= 0} is a suffix, and on
each side of the threshold the weight matrix is rank-1:  w_jk = C1_j * A_k  (k above)  or  C2_j * A'_k  (k below).
This script (1) checks that decomposition against the direct evaluation, (2) evaluates the block structure the GPU
kernel would see -- 16-destination tiles x 16-member k-steps, members sorted by el, destinations sorted by -er -- and
reports how many (tile, step) blocks still need the element-wise max ("mixed") for el/er taken from the real model.

    python tools/rank1_star_prototype.py [n] [instances]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from oracle import model_port  # noqa: E402  (test infrastructure; this is an analysis script, not product code)


def scores_from_model(n, B, seed=0):
    """el, er [B*N, 8] (log2 domain) of the first and the last layer of the seeded random-init model."""
    torch.manual_seed(seed)
    port = model_port.EdgeModelPort(1, 128, 1, 3, n_heads=8)
    model_port.randomize_bn_stats(port, seed=1)
    port.eval().double()
    g = model_port.EdgeListGraph.kn_line_graph(n, batch=B)
    rng = np.random.default_rng(20211005)
    P = rng.random((B, n, 2))
    D = np.sqrt(((P[:, :, None, :] - P[:, None, :, :]) ** 2).sum(-1))
    iu = np.triu_indices(n, 1)
    x = torch.as_tensor((D[:, iu[0], iu[1]] / np.sqrt(2)).reshape(-1, 1))
    out = []
    with torch.no_grad():
        h = port.embed_layer(x)
        for li, layer in enumerate(port.message_passing_layers):
            gat = layer.message_passing.module
            ft = gat.fc(h).view(-1, 8, 16)
            el = (ft * gat.attn_l).sum(-1) * np.log2(np.e)
            er = (ft * gat.attn_r).sum(-1) * np.log2(np.e)
            if li in (0, len(port.message_passing_layers) - 1):
                out.append((el.numpy(), er.numpy(), ft.numpy()))
            h = layer(g, h)
    return out


def star_members(i, n):
    """line-graph nodes {i,k} for k != i, in k order, and k itself"""
    ks = np.array([k for k in range(n) if k != i])
    lo, hi = np.minimum(i, ks), np.maximum(i, ks)
    return lo * (2 * n - lo - 1) // 2 + (hi - lo - 1), ks


def check_and_count(el, er, ft, n, stars):
    lrelu = lambda s: np.maximum(s, 0.2 * s)      # noqa: E731
    max_err, mixed, total = 0.0, 0, 0
    for (b, i) in stars:
        nodes, ks = star_members(i, n)
        N = n * (n - 1) // 2
        idx = b * N + nodes
        for hd in range(8):
            e_l, e_r, f = el[idx, hd], er[idx, hd], ft[idx, hd, :]
            # ---- direct evaluation: destination j (a member slot), sources = the other members
            s = e_l[None, :] + e_r[:, None]                      # [dest j, member k]
            L = lrelu(s)
            np.fill_diagonal(L, -np.inf)
            mx = L.max(1)
            W = np.exp2(L - mx[:, None])
            direct_num, direct_den = W @ f, W.sum(1)
            # ---- factorised: reference mx'_j = lrelu(m1 + er_j) (upper bound), A_k, A'_k, C1_j, C2_j
            m1 = e_l.max()
            ref = lrelu(m1 + e_r)
            A, A5 = np.exp2(e_l - m1), np.exp2(0.2 * (e_l - m1))
            C1, C2 = np.exp2(m1 + e_r - ref), np.exp2(0.2 * (m1 + e_r) - ref)
            order = np.argsort(e_l, kind='stable')               # members ascending in el
            thr = np.searchsorted(e_l[order], -e_r, side='left') # first sorted member with el_k + er_j >= 0
            rank = np.empty_like(order); rank[order] = np.arange(len(order))
            fs, As, A5s = f[order], A[order], A5[order]
            suffix1 = np.concatenate([np.cumsum((As[:, None] * fs)[::-1], 0)[::-1], np.zeros((1, 16))])
            prefix2 = np.concatenate([np.zeros((1, 16)), np.cumsum(A5s[:, None] * fs, 0)])
            suf1d = np.concatenate([np.cumsum(As[::-1])[::-1], [0.0]])
            pre2d = np.concatenate([[0.0], np.cumsum(A5s)])
            num = C1[:, None] * suffix1[thr] + C2[:, None] * prefix2[thr]
            den = C1 * suf1d[thr] + C2 * pre2d[thr]
            own_hi = rank >= thr                                  # the destination's own slot, to be excluded
            own_w = np.where(own_hi, C1 * A, C2 * A5)
            num -= own_w[:, None] * f
            den -= own_w
            # both are softmax numerators/denominators w.r.t. different references: compare the normalised outputs
            err = np.abs(num / den[:, None] - direct_num / direct_den[:, None]).max()
            max_err = max(max_err, err)
            # ---- block structure: destinations sorted by threshold, 16 x 16 blocks
            dorder = np.argsort(thr, kind='stable')
            m = len(order)
            for t0 in range(0, m, 16):
                tt = thr[dorder[t0:t0 + 16]]
                for s0 in range(0, m, 16):
                    total += 1
                    # block is unmixed iff every destination's threshold lies outside (s0, s0+16) ... i.e. <= s0 or >= s0+16
                    if np.any((tt > s0) & (tt < min(s0 + 16, m))):
                        mixed += 1
    return max_err, mixed, total


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(0)
    for name, (el, er, ft) in zip(('first layer', 'last layer'), scores_from_model(n, B)):
        stars = [(int(rng.integers(B)), int(rng.integers(n))) for _ in range(6)]
        err, mixed, total = check_and_count(el, er, ft, n, stars)
        print(f'n={n} {name}: max |factorised - direct| = {err:.2e} (normalised outputs, fp64); '
              f'mixed 16x16 blocks {mixed}/{total} = {100 * mixed / total:.1f} %  '
              f'(el range {el.min():.1f}..{el.max():.1f}, er range {er.min():.1f}..{er.max():.1f} log2 units)')


if __name__ == '__main__':
    main()
