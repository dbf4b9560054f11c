"""ORACLE (test infrastructure, not product): CPU/torch restatement of the model half.

* ``GATConvPort`` restates ``dgl.nn.GATConv`` as pinned by the reference
  (dgl-cu111==0.6.1, /root/reference/Pipfile.lock:316-329; call site
  /root/reference/gnngls/models.py:23).  DGL is a third-party dependency that is
  NOT in /root/reference and not installable here, so this follows the published
  algorithm:
      ft = fc(h).view(N,H,F); el = (ft*attn_l).sum(-1); er = (ft*attn_r).sum(-1)
      e_uv = leaky_relu(el_u + er_v, 0.2); a = edge_softmax(e) over in-edges of v
      out_v = sum_u a_uv * ft_u   (+ bias for DGL >= 0.7 checkpoints)
  "parity unpinned" for this op: no DGL golden output exists offline.
* ``dense_gat_reference`` is a second, independent restatement of the same
  formula (dense masked softmax) used to cross-check the first.
* ``EdgeModelPort`` restates /root/reference/gnngls/models.py:5-70 with the same
  module tree, so its ``state_dict`` has exactly the reference's 140 keys.

Works in fp32 or fp64 (``.double()``); fp64 is the accuracy yardstick.
"""
import math

import numpy as np
import torch
import torch.nn as nn


def kn_rank(i, j, n):
    """Line-graph node id of TSP edge (i<j): sorted-tuple order (datasets.py:56-60)."""
    return i * (2 * n - i - 1) // 2 + (j - i - 1)


def kn_edge_list(n):
    """int64 [N,2] array of (i,j), i<j, in line-graph node order (== ndata['e'])."""
    iu = np.triu_indices(n, 1)
    return np.stack([iu[0], iu[1]], axis=1).astype(np.int64)


def kn_line_graph_edges(n):
    """(src, dst) int64 arrays of the directed line graph of K_n, grouped by dst.

    Node (i,j) receives from every (i,k) and (k,j), k not in {i,j}: 2(n-2) in-edges.
    """
    es = kn_edge_list(n)
    N = es.shape[0]
    idx = np.zeros((n, n), dtype=np.int64)
    idx[es[:, 0], es[:, 1]] = np.arange(N)
    idx = idx + idx.T
    src = np.empty((N, 2 * (n - 2)), dtype=np.int64)
    for v, (i, j) in enumerate(es):
        ks = np.array([k for k in range(n) if k != i and k != j], dtype=np.int64)
        src[v, : n - 2] = idx[i, ks]
        src[v, n - 2:] = idx[ks, j]
    dst = np.repeat(np.arange(N, dtype=np.int64), 2 * (n - 2))
    return src.reshape(-1), dst


class EdgeListGraph:
    """Minimal stand-in for the DGLGraph surface the reference model touches."""

    regular_degree = None

    def __init__(self, src, dst, num_nodes):
        self.src = torch.as_tensor(src, dtype=torch.int64)
        self.dst = torch.as_tensor(dst, dtype=torch.int64)
        self._n = int(num_nodes)
        self.ndata = {}

    @classmethod
    def kn_line_graph(cls, n, batch=1):
        s, d = kn_line_graph_edges(n)
        N = n * (n - 1) // 2
        if batch > 1:
            offs = (np.arange(batch, dtype=np.int64) * N)[:, None]
            s = (s[None, :] + offs).reshape(-1)
            d = (d[None, :] + offs).reshape(-1)
        g = cls(s, d, N * batch)
        g.regular_degree = 2 * (n - 2)      # dst-sorted, constant in-degree: enables the dense fast path
        g.ndata['e'] = torch.as_tensor(np.tile(kn_edge_list(n), (batch, 1)))
        return g

    def number_of_nodes(self):
        return self._n

    def edges(self):
        return self.src, self.dst

    def to(self, device):
        return self


class GATConvPort(nn.Module):
    def __init__(self, in_feats, out_feats, num_heads, negative_slope=0.2, bias=False):
        super().__init__()
        self._heads, self._out = num_heads, out_feats
        self._slope = negative_slope
        self.fc = nn.Linear(in_feats, out_feats * num_heads, bias=False)
        self.attn_l = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.attn_r = nn.Parameter(torch.empty(1, num_heads, out_feats))
        if bias:
            self.bias = nn.Parameter(torch.zeros(num_heads * out_feats))
        else:
            self.bias = None    # DGL 0.6.1: no bias tensor in the state_dict
        gain = nn.init.calculate_gain('relu')
        nn.init.xavier_normal_(self.fc.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        nn.init.xavier_normal_(self.attn_r, gain=gain)

    def forward(self, graph, feat):
        N = graph.number_of_nodes()
        H, F = self._heads, self._out
        src, dst = graph.edges()
        ft = self.fc(feat).view(N, H, F)
        el = (ft * self.attn_l).sum(-1)
        er = (ft * self.attn_r).sum(-1)
        if getattr(self, '_tf32_aggregate', False):
            ft = tf32_round(ft)
        deg = getattr(graph, 'regular_degree', None)
        if deg is not None:
            # same formula on a dst-sorted constant-degree graph, without scatter ops (faster on CPU)
            e = torch.nn.functional.leaky_relu(el[src].view(N, deg, H) + er[:, None, :], self._slope)
            a = torch.softmax(e, dim=1)
            out = torch.einsum('ndh,ndhf->nhf', a, ft[src].view(N, deg, H, F))
            if self.bias is not None:
                out = out + self.bias.view(1, H, F)
            return out
        e = torch.nn.functional.leaky_relu(el[src] + er[dst], self._slope)      # [E,H]
        idx = dst[:, None].expand(-1, H)
        emax = torch.full((N, H), -math.inf, dtype=e.dtype).scatter_reduce(0, idx, e, 'amax')
        p = torch.exp(e - emax[dst])
        denom = torch.zeros((N, H), dtype=e.dtype).index_add_(0, dst, p)
        a = p / denom[dst]
        out = torch.zeros((N, H, F), dtype=ft.dtype).index_add_(0, dst, a[:, :, None] * ft[src])
        if self.bias is not None:
            out = out + self.bias.view(1, H, F)
        return out


def dense_gat_reference(h, fc_w, attn_l, attn_r, adj, slope=0.2):
    """Independent dense restatement.  adj[v,u]=True iff edge u->v.  Returns [N,H*F]."""
    H, F = attn_l.shape[-2], attn_l.shape[-1]
    N = h.shape[0]
    ft = (h @ fc_w.t()).view(N, H, F)
    el = torch.einsum('nhf,hf->nh', ft, attn_l.view(H, F))
    er = torch.einsum('nhf,hf->nh', ft, attn_r.view(H, F))
    s = er[:, None, :] + el[None, :, :]                     # [v,u,h]
    s = torch.where(s > 0, s, slope * s)
    s = s.masked_fill(~adj[:, :, None], -math.inf)
    a = torch.softmax(s, dim=1)
    return torch.einsum('vuh,uhf->vhf', a, ft).reshape(N, H * F)


class _Skip(nn.Module):
    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, x, G=None):
        y = self.module(x) if G is None else self.module(G, x).reshape(G.number_of_nodes(), -1)
        return x + y


class _AttnLayer(nn.Module):
    def __init__(self, embed_dim, n_heads, hidden_dim, gat_bias=False):
        super().__init__()
        self.message_passing = _Skip(GATConvPort(embed_dim, embed_dim // n_heads, n_heads, bias=gat_bias))
        self.feed_forward = nn.Sequential(
            nn.BatchNorm1d(embed_dim),
            _Skip(nn.Sequential(nn.Linear(embed_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, embed_dim))),
            nn.BatchNorm1d(embed_dim),
        )

    def forward(self, G, x):
        return self.feed_forward(self.message_passing(x, G=G))


class EdgeModelPort(nn.Module):
    """models.py:44-70.  NB the reference builds ``n_heads`` layers, not ``n_layers`` (models.py:60)."""

    def __init__(self, in_dim, embed_dim, out_dim, n_layers, n_heads=1, gat_bias=False):
        super().__init__()
        self.embed_dim = embed_dim
        self.embed_layer = nn.Linear(in_dim, embed_dim)
        self.message_passing_layers = nn.Sequential(
            *[_AttnLayer(embed_dim, n_heads, 512, gat_bias) for _ in range(n_heads)])
        self.decision_layer = nn.Linear(embed_dim, out_dim)

    def forward(self, G, x):
        h = self.embed_layer(x)
        for layer in self.message_passing_layers:
            h = layer(G, h)
        return self.decision_layer(h)


def randomize_bn_stats(model, seed=1):
    """Give eval-mode BatchNorm non-trivial running stats/affine so parity tests bite."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm1d):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.weight.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    return model


# ---------------------------------------------------------------------------------------------
# TF32 emulation: what an "allow_tf32" evaluation of the same model computes.  torch 1.11 (the
# reference's pin) enables TF32 matmuls by default on Ampere+, i.e. GEMM operands are rounded to a
# 10-bit mantissa and products are accumulated in fp32.  Used to calibrate the GPU tolerance.
# ---------------------------------------------------------------------------------------------
def tf32_round(t):
    """Round to TF32 (nearest, ties away from zero == PTX cvt.rna.tf32.f32); keeps dtype."""
    f = t.detach().to(torch.float32).contiguous()
    r = ((f.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    return r.to(t.dtype)


class emulate_tf32:
    """Context manager: every nn.Linear with in_features >= 128 rounds its input and weight to TF32
    (exact products, wide accumulation when the model is .double()); GATConvPort additionally rounds
    the aggregated features, like the tensor-core aggregate kernel."""

    def __init__(self, model, round_aggregate=True):
        self.model, self.round_aggregate, self._saved = model, round_aggregate, []

    def __enter__(self):
        for m in self.model.modules():
            if isinstance(m, nn.Linear) and m.in_features >= 128 and m.out_features >= 128:
                self._saved.append((m, m.forward))
                m.forward = (lambda x, m=m: torch.nn.functional.linear(tf32_round(x), tf32_round(m.weight), m.bias))
            if isinstance(m, GATConvPort):
                m._tf32_aggregate = self.round_aggregate
        return self

    def __exit__(self, *exc):
        for m, f in self._saved:
            m.forward = f
        for m in self.model.modules():
            if isinstance(m, GATConvPort):
                m._tf32_aggregate = False
