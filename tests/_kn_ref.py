"""fp64 evaluation of the GATConv aggregate + skip + BN1 on the line graph of K_n for a SAMPLE of destination
rows, with the adjacency computed arithmetically (no edge list): usable at sizes where neither the torch oracle
nor a CSR index fits.  Semantics: SURVEY.md Appendix A (dgl.nn.GATConv reached from gnngls/models.py:23);
scores are in the log2 domain as the C ABI stores them (include/gnngls_b200.h)."""
import numpy as np


def kn_node(i, j, n):
    i, j = np.minimum(i, j), np.maximum(i, j)
    return i * (2 * n - i - 1) // 2 + (j - i - 1)


def kn_pairs(n):
    iu = np.triu_indices(n, 1)
    return iu[0], iu[1]


def aggregate_rows(n, rows, ft, el, er, h, bias, sc, sh):
    """rows: global node ids (any instance of the batch).  ft [M,128], el/er [M,8], h [M,128] numpy arrays;
    bias may be None.  Returns fp64 [len(rows),128]."""
    N = n * (n - 1) // 2
    pi, pj = kn_pairs(n)
    ft, el, er, h = (np.asarray(a, dtype=np.float64) for a in (ft, el, er, h))
    out = np.empty((len(rows), 128))
    ks = np.arange(n)
    for t, v in enumerate(rows):
        b, loc = divmod(int(v), N)
        i, j = int(pi[loc]), int(pj[loc])
        keep = ks[(ks != i) & (ks != j)]
        src = np.concatenate([kn_node(np.full_like(keep, i), keep, n), kn_node(keep, np.full_like(keep, j), n)]) + b * N
        e = el[src] + er[v][None, :]                      # [deg,8]
        e = np.maximum(e, 0.2 * e)
        w = np.exp2(e - e.max(0, keepdims=True))
        w /= w.sum(0, keepdims=True)
        agg = np.einsum('dh,dhf->hf', w, ft[src].reshape(-1, 8, 16)).reshape(128)
        if bias is not None:
            agg = agg + np.asarray(bias, dtype=np.float64)
        out[t] = (h[v] + agg) * np.asarray(sc, dtype=np.float64) + np.asarray(sh, dtype=np.float64)
    return out
