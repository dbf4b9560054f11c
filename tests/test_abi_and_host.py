"""CPU-only checks: the C-ABI library builds, loads and exports every symbol the header declares;
host-side logic (graph construction, TF32 rounding, scaler semantics, state_dict compatibility)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from gnngls_b200 import _lib, graph, instances, models
from oracle import model_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'gnngls_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gnngls_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert sorted(_lib.SIGNATURES) == names
    assert lib.gnngls_abi_version() == 1
    assert lib.gnngls_sizeof_gls_args() == ctypes.sizeof(_lib.GlsArgs)


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, 'gnngls_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('no CPU or eager', ''), f


def test_ops_refuse_cpu_tensors():
    from gnngls_b200 import _ops
    with pytest.raises(TypeError):
        _ops.moves_eval(0, torch.zeros(1, 5, 5, dtype=torch.float64), torch.zeros(1, 6, dtype=torch.int32))
    m = models.EdgePropertyPredictionModel(1, 128, 1, 3, n_heads=8).eval()
    with pytest.raises(RuntimeError):
        m(graph.LineGraph.complete(5), torch.zeros(10, 1))


def test_kn_graph_matches_oracle_edges():
    for n in (3, 4, 7, 12):
        ip, ix = graph.kn_csr(n)
        s, d = model_port.kn_line_graph_edges(n)
        N = n * (n - 1) // 2
        assert ip.tolist() == [2 * (n - 2) * k for k in range(N + 1)]
        got = {(int(u), int(v)) for v in range(N) for u in ix[ip[v]:ip[v + 1]]}
        assert got == set(zip(s.tolist(), d.tolist()))
        assert np.array_equal(graph.kn_edges(n), model_port.kn_edge_list(n))


def test_batched_graph_and_from_edges():
    g = graph.LineGraph.complete(5, 3)
    assert g.number_of_nodes() == 30 and g.ndata['e'].shape == (30, 2)
    ip, ix = g.csr()
    assert ip.shape[0] == 31 and int(ip[-1]) == 30 * 6
    assert int(ix[int(ip[10]):int(ip[11])].min()) >= 10 and int(ix[int(ip[19]):int(ip[20])].max()) < 20
    s, d = model_port.kn_line_graph_edges(5)
    perm = np.random.default_rng(0).permutation(len(s))
    g2 = graph.LineGraph.from_edges(s[perm], d[perm], 10)
    ip2, ix2 = g2.csr()
    ip1, ix1 = graph.kn_csr(5)
    assert ip2.tolist() == ip1.tolist()
    for v in range(10):
        assert sorted(ix2[ip2[v]:ip2[v + 1]].tolist()) == sorted(ix1[ip1[v]:ip1[v + 1]].tolist())
    gb = graph.batch([graph.LineGraph.complete(5), g2])
    assert gb.kind == 'csr' and gb.number_of_nodes() == 20
    with pytest.raises(ValueError):
        graph.LineGraph.from_edges([0], [1], 3)


def test_from_networkx_line_graph_matches_reference_construction():
    import networkx as nx
    lG = nx.line_graph(nx.complete_graph(6))
    g = graph.LineGraph.from_networkx_line_graph(lG)
    assert g.ndata['e'].tolist() == graph.kn_edges(6).tolist()
    ip, ix = g.csr()
    ip1, ix1 = graph.kn_csr(6)
    for v in range(15):
        assert sorted(ix[ip[v]:ip[v + 1]].tolist()) == sorted(ix1[ip1[v]:ip1[v + 1]].tolist())


def test_tf32_round_matches_definition():
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -11 + 2 ** -20, -1.0 - 2 ** -11, 3.14159265, 1e-20, -7.5e10])
    r = models.tf32_round(x)
    assert (r.view(torch.int32) & 0x1FFF).eq(0).all()
    assert r[0] == 1.0 and r[1] == 1.0 + 2 ** -10 and r[3] == -1.0 - 2 ** -10     # ties away from zero
    assert ((r - x).abs() <= x.abs() * 2 ** -11).all()


def test_state_dict_compat_with_reference_layout():
    torch.manual_seed(0)
    port = model_port.EdgeModelPort(1, 128, 1, 3, n_heads=8)
    m = models.EdgePropertyPredictionModel(1, 128, 1, 3, n_heads=8)
    assert list(m.state_dict().keys()) == list(port.state_dict().keys())
    m.load_state_dict(port.state_dict(), strict=True)
    # DGL >= 0.7 checkpoints carry a GATConv bias: must load strictly too, and round-trip
    port_b = model_port.EdgeModelPort(1, 128, 1, 3, n_heads=8, gat_bias=True)
    m2 = models.EdgePropertyPredictionModel(1, 128, 1, 3, n_heads=8)
    m2.load_state_dict(port_b.state_dict(), strict=True)
    assert list(m2.state_dict().keys()) == list(port_b.state_dict().keys())
    # checkpoint file layout of scripts/train.py:60-67
    ck = {'epoch': 1, 'model_state_dict': port.state_dict(), 'optimizer_state_dict': {}, 'loss': 0.0, 'val_loss': 0.0}
    m3 = models.EdgePropertyPredictionModel(1, 128, 1, 3, n_heads=8)
    m3.load_state_dict(ck['model_state_dict'])
    assert len(m3.message_passing_layers) == 8       # n_heads layers, not n_layers (models.py:60)


def test_training_mode_and_unsupported_dims_fail_loudly():
    m = models.EdgePropertyPredictionModel(1, 64, 1, 3, n_heads=4)
    assert len(m.state_dict()) > 0


def test_scaler_semantics_match_sklearn():
    from sklearn.preprocessing import MinMaxScaler
    rng = np.random.default_rng(0)
    sc = MinMaxScaler().fit(rng.random((1000, 1)) * 1.3 + 0.01)
    x = (rng.random((5000, 1)) * 1.4).astype(np.float32)
    y = sc.transform(x)
    s, m = float(sc.scale_[0]), float(sc.min_[0])
    mine = (x.astype(np.float64) * s).astype(np.float32)
    mine = (mine.astype(np.float64) + m).astype(np.float32)
    assert np.array_equal(mine, y)          # the recipe csrc/glue.cu implements
    z = sc.inverse_transform(y.copy())
    inv = (y.astype(np.float64) - m).astype(np.float32)
    inv = (inv.astype(np.float64) / s).astype(np.float32)
    assert np.array_equal(inv, z)


def test_synthetic_instances_are_reproducible():
    P, D = instances.random_instances(3, 10)
    P2, D2 = instances.random_instances(3, 10)
    assert np.array_equal(D, D2) and np.array_equal(D, D.transpose(0, 2, 1)) and (np.diagonal(D, axis1=1, axis2=2) == 0).all()
    x = instances.edge_features(D)
    assert x.shape == (3, 45) and x.dtype == np.float32 and x[0, 0] == np.float32(D[0, 0, 1])


def test_python_constants_match_header_enums():
    """The ctypes side passes these as plain ints: they must be the values include/gnngls_b200.h declares."""
    import os
    import re
    from gnngls_b200 import _ops
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'gnngls_b200.h')).read()
    enums = {m.group(1): int(m.group(2)) for m in re.finditer(r'\b(GNNGLS_[A-Z0-9_]+)\s*=\s*(-?\d+)', hdr)}
    assert enums['GNNGLS_FT_F32'] == _ops.FT_F32 and enums['GNNGLS_FT_TF32'] == _ops.FT_TF32 and enums['GNNGLS_FT_F16'] == _ops.FT_F16
    assert enums['GNNGLS_DENSE_TCGEN05'] == _ops.DENSE_TCGEN05 and enums['GNNGLS_DENSE_SIMT'] == _ops.DENSE_SIMT
    assert enums['GNNGLS_DENSE_TCGEN05_F16'] == _ops.DENSE_TCGEN05_F16


def test_operand_dtype_switches(monkeypatch):
    from gnngls_b200 import models
    monkeypatch.delenv('GNNGLS_OP_DTYPE', raising=False)
    monkeypatch.delenv('GNNGLS_FF_DTYPE', raising=False)
    assert models._op_f16()
    monkeypatch.setenv('GNNGLS_OP_DTYPE', 'tf32')
    assert not models._op_f16()
    monkeypatch.setenv('GNNGLS_OP_DTYPE', 'f16')
    monkeypatch.setenv('GNNGLS_FF_DTYPE', 'TF32')
    assert not models._op_f16()


def test_no_divergent_uniform_register_moves_in_sass():
    """A cache-policy operand (`L2::cache_hint`) travels in a uniform register.  ptxas reloads it with a PREDICATED R2UR when
    the hinted instruction sits in a branch only some lanes take -- an illegal instruction at run time (seen on B200 with a
    lane-subset cp.async).  The built library must not contain one."""
    import shutil
    import subprocess
    from gnngls_b200 import build
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not available')
    lib = build.build()
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    # (R2UR.BROADCAST under an elect-style predicate, as in the tcgen05 GEMM kernels, is the legal form)
    bad = [l for l in sass.splitlines() if 'R2UR ' in l and '@' in l.split('R2UR')[0]]
    assert not bad, bad[:5]


def test_cluster_tier_uses_cluster_barriers_and_distributed_shared_memory():
    """The cluster tier of the search kernels must really be a thread-block-cluster program: cluster barriers,
    generic stores through mapa-translated (distributed shared memory) addresses and cp.async staging in the SASS of
    gls_cluster_kernel, and none of the cluster barriers in the one-CTA kernel."""
    import shutil
    import subprocess
    from gnngls_b200 import build
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not available')
    lib = build.build()
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    funcs = {}
    name = None
    for line in sass.splitlines():
        if 'Function :' in line:
            name = line.split('Function :')[1].strip()
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    def body(key, exclude=()):
        hits = [k for k in funcs if key in k and not any(e in k for e in exclude)]
        assert hits, key
        return '\n'.join('\n'.join(funcs[k]) for k in hits)
    cl = body('gls_cluster_kernel')
    assert 'UCGABAR_ARV' in cl and 'UCGABAR_WAIT' in cl and 'LDGSTS' in cl and 'ST.E' in cl
    solo = body('gls_kernel', exclude=('cluster',))
    assert 'UCGABAR' not in solo


def test_int16_record_scale_arithmetic():
    """The K_n kernel's partial records (csrc/gat_kn_tc.cu): (v, den, M) == (v c, den c, M - log2 c) for any c > 0, and the c the
    epilogue derives from the exponent of max|v| puts every rounded numerator inside int16 with 15 significant bits.  Restated in
    numpy fp32 with the kernel's constants; the merge formula must give the same result from the scaled record (up to the int16
    quantum) as from the unscaled one."""
    rng = np.random.default_rng(0)
    f32 = np.float32
    for trial in range(2000):
        mag = f32(2.0) ** f32(rng.integers(-20, 24))
        v = (rng.standard_normal(16).astype(f32) * mag).astype(f32)
        den, M = f32(rng.uniform(0.5, 99.0)), f32(rng.uniform(-30.0, 30.0))
        mx = np.abs(v).max()
        E = int(np.clip(int(np.array(mx, dtype=f32).view(np.int32)) >> 23, 40, 230))
        scale = f32(np.array((268 - E) << 23, dtype=np.int32).view(f32) * f32(0.999969482421875))
        q = np.rint(v.astype(np.float64) * np.float64(scale))                     # fma(v, c, 1.5 * 2^23) keeps rint(v c) in its low bits
        assert np.abs(q).max() <= 32767 and (mx == 0 or np.abs(q).max() >= 16383), (mx, scale)
        den_s, M_s = f32(den * scale), f32(M + f32(E - 141) + f32(4.4028e-5))
        # merge with a second, unscaled partial (algorithms of merge_rows): out = (v1 s1 + v2 s2) / (den1 s1 + den2 s2), s = 2^(M - max M)
        v2 = (rng.standard_normal(16) * float(mag)).astype(np.float64)
        den2, M2 = rng.uniform(0.5, 99.0), float(M) + rng.uniform(-3.0, 3.0)

        def merge(va, da, Ma):
            mxm = max(Ma, M2)
            s1, s2 = 2.0 ** (Ma - mxm), 2.0 ** (M2 - mxm)
            return (va * s1 + v2 * s2) / (da * s1 + den2 * s2)
        ref = merge(v.astype(np.float64), float(den), float(M))
        got = merge(q, float(den_s), float(M_s))
        tol = 2.0 ** -15 * float(mx) / float(den) * 1.5 + 1e-6 * np.abs(ref).max()    # half an int16 quantum of the head's largest numerator
        assert np.abs(got - ref).max() <= tol + 1e-30, (trial, np.abs(got - ref).max(), tol)
