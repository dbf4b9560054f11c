"""LineGraph: the graph argument of EdgePropertyPredictionModel.forward.

The reference passes a DGLGraph built as ``dgl.from_networkx(nx.line_graph(G), node_attrs=['e'])``
(/root/reference/gnngls/datasets.py:56-60): nodes are the TSP edges (i<j) in sorted order, two
line-graph nodes are adjacent iff the TSP edges share a vertex.  The model only touches
``G.number_of_nodes()`` (models.py:12,39) plus whatever GATConv needs; callers touch ``ndata`` and
``.to(device)`` (scripts/test.py:72-81).  This class exposes that surface with two storage forms:

* ``kind == 'kn'``  — a batch of line graphs of K_n described by ``(n, batch)`` only; the kernels
  compute the adjacency arithmetically (no index memory);
* ``kind == 'csr'`` — any graph as destination-sorted CSR (int32 ``indptr``/``indices`` of the
  in-edges), e.g. from an edge list or a networkx line graph.
"""
import numpy as np
import torch


def kn_edges(n):
    """int64 [N,2]: TSP edge (i,j), i<j, of each line-graph node — the reference's ndata['e']."""
    iu = np.triu_indices(n, 1)
    return np.stack([iu[0], iu[1]], axis=1).astype(np.int64)


def kn_rank(i, j, n):
    if i > j:
        i, j = j, i
    return i * (2 * n - i - 1) // 2 + (j - i - 1)


def kn_csr(n):
    """(indptr, indices) int32 numpy arrays of the line graph of K_n (in-edges per node)."""
    es = kn_edges(n)
    N = es.shape[0]
    idx = np.zeros((n, n), dtype=np.int64)
    idx[es[:, 0], es[:, 1]] = np.arange(N)
    idx = idx + idx.T
    deg = 2 * (n - 2)
    ks = np.arange(n)
    indices = np.empty((N, deg), dtype=np.int32)
    for v, (i, j) in enumerate(es):
        keep = ks[(ks != i) & (ks != j)]
        indices[v, : n - 2] = idx[i, keep]
        indices[v, n - 2:] = idx[keep, j]
    indptr = (np.arange(N + 1, dtype=np.int64) * deg).astype(np.int32)
    return indptr, indices.reshape(-1)


class LineGraph:
    def __init__(self, kind, num_nodes, n=None, batch_size=1, indptr=None, indices=None, device='cpu'):
        self.kind = kind
        self._num_nodes = int(num_nodes)
        self.n = n
        self.batch_size = int(batch_size)
        self.indptr, self.indices = indptr, indices
        self.device = torch.device(device)
        self.ndata = {}

    # ---- constructors
    @classmethod
    def complete(cls, n, batch_size=1, device='cpu'):
        """Batch of `batch_size` line graphs of K_n (block-diagonal, like dgl.batch)."""
        if n < 3:
            raise ValueError('line graph of K_n needs n >= 3')
        g = cls('kn', batch_size * n * (n - 1) // 2, n=n, batch_size=batch_size, device=device)
        e = torch.from_numpy(kn_edges(n))
        g.ndata['e'] = (e.repeat(batch_size, 1) if batch_size > 1 else e).to(device)
        return g

    @classmethod
    def from_edges(cls, src, dst, num_nodes, device='cpu'):
        """Directed edges u->v (v aggregates from u), any order; converted to in-edge CSR."""
        src = np.asarray(src, dtype=np.int64)
        dst = np.asarray(dst, dtype=np.int64)
        order = np.argsort(dst, kind='stable')
        counts = np.bincount(dst, minlength=num_nodes)
        if counts.min() == 0:
            # dgl.nn.GATConv(allow_zero_in_degree=False) raises DGLError in this case
            raise ValueError('graph has nodes with zero in-degree')
        indptr = np.zeros(num_nodes + 1, dtype=np.int64)
        np.cumsum(counts, out=indptr[1:])
        if indptr[-1] >= 2 ** 31:
            raise ValueError('too many edges for int32 CSR')
        g = cls('csr', num_nodes, indptr=torch.from_numpy(indptr.astype(np.int32)),
                indices=torch.from_numpy(src[order].astype(np.int32)), device='cpu')
        return g.to(device)

    @classmethod
    def from_networkx_line_graph(cls, lG, device='cpu'):
        """Mirror of datasets.py:57-60: nodes relabelled in sorted order, each edge both ways."""
        nodes = sorted(lG.nodes)
        rank = {e: k for k, e in enumerate(nodes)}
        src, dst = [], []
        for a, b in lG.edges:
            src += [rank[a], rank[b]]
            dst += [rank[b], rank[a]]
        g = cls.from_edges(src, dst, len(nodes), device=device)
        g.ndata['e'] = torch.tensor(nodes, dtype=torch.int64, device=device)
        return g

    # ---- DGLGraph surface used by the reference
    def number_of_nodes(self):
        return self._num_nodes

    num_nodes = number_of_nodes

    def to(self, device):
        device = torch.device(device)
        g = LineGraph(self.kind, self._num_nodes, self.n, self.batch_size,
                      None if self.indptr is None else self.indptr.to(device),
                      None if self.indices is None else self.indices.to(device), device)
        g.ndata = {k: v.to(device) for k, v in self.ndata.items()}
        return g

    def __deepcopy__(self, memo):          # datasets.py:91 deep-copies the template graph
        g = LineGraph(self.kind, self._num_nodes, self.n, self.batch_size, self.indptr, self.indices, self.device)
        g.ndata = {k: v.clone() for k, v in self.ndata.items()}
        return g

    def csr(self):
        """(indptr, indices) on self.device; materialised on demand for 'kn' graphs."""
        if self.indptr is None:
            ip, ix = kn_csr(self.n)
            N, E = self.n * (self.n - 1) // 2, ix.shape[0]
            if self.batch_size > 1:
                ix = (ix[None, :].astype(np.int64) + (np.arange(self.batch_size, dtype=np.int64) * N)[:, None]).reshape(-1)
                ip = np.concatenate([(ip[None, :-1].astype(np.int64) + (np.arange(self.batch_size, dtype=np.int64) * E)[:, None]).reshape(-1),
                                     np.array([E * self.batch_size], dtype=np.int64)])
                if ip[-1] >= 2 ** 31:
                    raise ValueError('too many edges for int32 CSR')
            self.indptr = torch.from_numpy(ip.astype(np.int32)).to(self.device)
            self.indices = torch.from_numpy(ix.astype(np.int32)).to(self.device)
        return self.indptr, self.indices


def batch(graphs):
    """dgl.batch for LineGraphs (scripts/train.py:118): block-diagonal union."""
    graphs = list(graphs)
    if all(g.kind == 'kn' and g.n == graphs[0].n for g in graphs):
        out = LineGraph.complete(graphs[0].n, sum(g.batch_size for g in graphs), graphs[0].device)
    else:
        ips, ixs, off_n, off_e = [], [], 0, 0
        for g in graphs:
            ip, ix = g.csr()
            ips.append(ip[:-1].to(torch.int64) + off_e)
            ixs.append(ix.to(torch.int64) + off_n)
            off_n += g.number_of_nodes()
            off_e += int(ip[-1])
        ips.append(torch.tensor([off_e], dtype=torch.int64, device=ips[0].device))
        out = LineGraph('csr', off_n, indptr=torch.cat(ips).to(torch.int32), indices=torch.cat(ixs).to(torch.int32),
                        device=graphs[0].device)
    keys = set(graphs[0].ndata)
    for k in keys:
        if all(k in g.ndata for g in graphs):
            out.ndata[k] = torch.cat([g.ndata[k] for g in graphs])
    return out
