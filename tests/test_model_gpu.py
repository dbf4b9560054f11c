"""GPU parity of EdgePropertyPredictionModel (through the C ABI) against the reference's own
models.py outputs (golden, fp64 yardstick) and the torch oracle.

Tolerances on the raw model output (|y| ~ 30 for the seeded random-init weights), yardstick = the
reference's own models.py evaluated in fp64 (golden `y64`):
  fp32 path  (SIMT dense + CSR or K_n aggregate, both fp32 arithmetic): max|y - y64| <= 1e-5 * max|y64| (CSR),
             2e-5 * max|y64| (K_n sorted-prefix kernel); the fp32 torch oracle itself is within 2e-6 * max|y64| of fp64;
  TF32 paths (tcgen05 dense contractions, fp16/TF32 operands): max|y - y64| <= 2 * E_tf32 + 1e-4 * max|y64|,
             where E_tf32 is the error of the SAME oracle evaluated with TF32-rounded GEMM operands
             (what torch 1.11's default allow_tf32 computes on Ampere+; oracle.model_port.emulate_tf32),
             measured per case at test time (E_tf32 ~ 1.2e-2 .. 2.2e-2 here, i.e. ~5e-4 relative).
"""
import numpy as np
import pytest
import torch

from gnngls_b200 import _ops, graph, instances, models
from oracle import model_port
from tests import _golden

pytestmark = pytest.mark.gpu

MODEL, TOP = _golden.load('model')
REL_FP32 = 1e-5


def tf32_budget(port, n, B, x, y64):
    """2 x (error of the TF32-emulating oracle vs fp64) + 1e-4 * max|y64|."""
    g = model_port.EdgeListGraph.kn_line_graph(n, batch=B)
    port = port.double()
    with torch.no_grad(), model_port.emulate_tf32(port):
        yt = port(g, torch.as_tensor(x).double()).numpy()
    port.float()
    return 2 * np.abs(yt - y64).max() + 1e-4 * np.abs(y64).max()


def make_models(gat_bias=False):
    torch.manual_seed(0)
    port = model_port.EdgeModelPort(1, 128, 1, 3, n_heads=8, gat_bias=gat_bias)
    model_port.randomize_bn_stats(port, seed=1)
    if gat_bias:
        for l in port.message_passing_layers:
            torch.nn.init.normal_(l.message_passing.module.bias, std=0.1)
    port.eval()
    m = models.EdgePropertyPredictionModel(1, 128, 1, 3, n_heads=8)
    m.load_state_dict(port.state_dict(), strict=True)
    return port, m.cuda().eval()


def run(m, n, B, x, dense, gat):
    m.dense_impl, m.gat_impl = dense, gat
    G = graph.LineGraph.complete(n, B, 'cuda')
    with torch.no_grad():
        return m(G, torch.as_tensor(x).cuda()).cpu().numpy()


@pytest.mark.parametrize('dense,gat', [('simt', 'csr'), ('simt', 'kn'), ('tcgen05', 'csr'), ('tcgen05', 'kn')])
def test_model_matches_reference_golden(dense, gat):
    port, m = make_models()
    for c in MODEL:
        n, B = c.nB.tolist()
        y = run(m, n, B, c.x, dense, gat)
        err = np.abs(y - c.y64).max()
        if dense == 'simt':
            tol = (REL_FP32 if gat == 'csr' else 2 * REL_FP32) * np.abs(c.y64).max()
        else:
            tol = tf32_budget(port, n, B, c.x, c.y64)
        print(f'{dense}+{gat} n={n} B={B}: max|y-y64|={err:.3e} tol={tol:.3e} (|y|max={np.abs(c.y64).max():.3f})')
        assert y.shape == c.y64.shape and np.isfinite(y).all()
        assert err <= tol, (dense, gat, n, err)


def test_paths_agree_and_batching_is_transparent():
    port, m = make_models()
    n, B = 12, 5
    _, D = instances.random_instances(B, n, seed=4)
    x = (instances.edge_features(D) / np.float32(np.sqrt(2))).reshape(-1, 1)
    ref = run(m, n, B, x, 'simt', 'csr')
    N = n * (n - 1) // 2
    for b in range(B):          # batched == per-instance (eval-mode BN is per-node)
        yb = run(m, n, 1, x[b * N:(b + 1) * N], 'simt', 'csr')
        assert np.array_equal(yb, ref[b * N:(b + 1) * N])
    with torch.no_grad():
        y64 = port.double()(model_port.EdgeListGraph.kn_line_graph(n, B), torch.as_tensor(x).double()).numpy()
    port.float()
    assert np.abs(ref - y64).max() < REL_FP32 * np.abs(y64).max()
    tol = tf32_budget(port, n, B, x, y64)
    assert np.abs(run(m, n, B, x, 'simt', 'kn') - y64).max() < 2 * REL_FP32 * np.abs(y64).max()
    assert np.abs(run(m, n, B, x, 'tcgen05', 'kn') - y64).max() < tol
    yk = run(m, n, B, x, 'tcgen05', 'kn')
    for b in range(B):          # TF32 path: batching is transparent too (bitwise)
        assert np.array_equal(run(m, n, 1, x[b * N:(b + 1) * N], 'tcgen05', 'kn'), yk[b * N:(b + 1) * N])
    # arbitrary-CSR entry point (edge list in random order) == K_n graph
    s, d = model_port.kn_line_graph_edges(n)
    perm = np.random.default_rng(0).permutation(len(s))
    g = graph.LineGraph.from_edges(s[perm], d[perm], N, device='cuda')
    m.dense_impl, m.gat_impl = 'simt', 'auto'
    with torch.no_grad():
        y1 = m(g, torch.as_tensor(x[:N]).cuda()).cpu().numpy()
    assert np.abs(y1 - ref[:N]).max() < REL_FP32 * np.abs(y64).max()


def test_gat_bias_checkpoints():
    port, m = make_models(gat_bias=True)
    n, B = 9, 2
    _, D = instances.random_instances(B, n, seed=5)
    x = (instances.edge_features(D) / np.float32(np.sqrt(2))).reshape(-1, 1)
    with torch.no_grad():
        y64 = port.double()(model_port.EdgeListGraph.kn_line_graph(n, B), torch.as_tensor(x).double()).numpy()
    port.float()
    assert np.abs(run(m, n, B, x, 'simt', 'csr') - y64).max() < REL_FP32 * np.abs(y64).max()
    assert np.abs(run(m, n, B, x, 'tcgen05', 'kn') - y64).max() < tf32_budget(port, n, B, x, y64)


def _tf32(t):
    return models.tf32_round(t.float().contiguous()).double()


def test_dense_kernels_against_tf32_rounded_oracle():
    """fc / FF blocks alone: the tensor-core result must match an fp64 evaluation whose operands are
    rounded to TF32 exactly as the kernels round them (tight), and the SIMT result the fp32 math."""
    from gnngls_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(1)
    for M in (1, 127, 128, 129, 1000, 4950 * 3 + 17):
        h = torch.randn(M, 128)
        W = torch.randn(128, 128) * 0.1
        al, ar = torch.randn(128) * 0.3, torch.randn(128) * 0.3
        W1, b1 = torch.randn(512, 128) * 0.1, torch.randn(512) * 0.1
        W2, b2 = torch.randn(128, 512) * 0.05, torch.randn(128) * 0.1
        sc, sh = torch.rand(128) + 0.5, torch.randn(128) * 0.1
        p = _ops._ptr
        for impl in (_ops.DENSE_SIMT, _ops.DENSE_TCGEN05):
            tc = impl == _ops.DENSE_TCGEN05
            hh = models.tf32_round(h) if tc else h
            Wd = models.tf32_round(W) if tc else W
            W1d, W2d = (models.tf32_round(W1), models.tf32_round(W2)) if tc else (W1, W2)
            hd, Wc, alc, arc = hh.cuda(), Wd.cuda(), al.cuda(), ar.cuda()      # keep device copies alive
            W1c, b1c, W2c, b2c, scc, shc = W1d.cuda(), b1.cuda(), W2d.cuda(), b2.cuda(), sc.cuda(), sh.cuda()
            el = torch.empty(M, 8, device='cuda'); er = torch.empty(M, 8, device='cuda')
            ft64 = hh.double() @ Wd.double().t()
            el64 = (ft64.view(M, 8, 16) * al.double().view(1, 8, 16)).sum(-1)
            er64 = (ft64.view(M, 8, 16) * ar.double().view(1, 8, 16)).sum(-1)
            LOG2E = 1.4426950408889634      # scores are stored in the log2 domain (include/gnngls_b200.h)
            for ft_dtype in (_ops.FT_F32, _ops.FT_TF32, _ops.FT_F16):
                ft = torch.empty(M, 128, device='cuda', dtype=torch.float16 if ft_dtype == _ops.FT_F16 else torch.float32)
                el.fill_(float('nan')); er.fill_(float('nan'))
                _lib.check(lib.gnngls_fc_forward(impl, p(hd), M, p(Wc), p(alc), p(arc), p(ft), ft_dtype, p(el), p(er),
                                                 _ops._stream()))
                # 10-bit-mantissa storage: |ft| <~ 8 here -> half an ulp is 2^-9 * 4
                tol = 1e-4 if ft_dtype == _ops.FT_F32 else 5e-3
                assert (ft.cpu().double() - ft64).abs().max() < tol, (M, impl, ft_dtype)
                if ft_dtype == _ops.FT_TF32:
                    assert torch.equal(ft.cpu(), models.tf32_round(ft.cpu()))
                assert (el.cpu().double() - LOG2E * el64).abs().max() < 2e-4 and (er.cpu().double() - LOG2E * er64).abs().max() < 2e-4
            nbytes = lib.gnngls_ff_workspace_bytes(impl, M)
            ws = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
            out = torch.empty(M, 128, device='cuda')
            out_r = torch.empty(M, 128, device='cuda')
            _lib.check(lib.gnngls_ff_forward(impl, p(hd), p(hd), M, p(W1c), p(b1c), p(W2c), p(b2c), p(scc), p(shc), p(out),
                                             p(out_r), _ops.FT_TF32, p(ws), nbytes, _ops._stream()))
            torch.cuda.synchronize()
            hid = torch.relu(hh.double() @ W1d.double().t() + b1.double())
            if tc:
                hid = _tf32(hid)
            o64 = (hh.double() + hid @ W2d.double().t() + b2.double()) * sc.double() + sh.double()
            assert (out.cpu().double() - o64).abs().max() < (1e-3 if tc else 2e-4), (M, impl)
            assert torch.equal(out_r.cpu(), models.tf32_round(out.cpu()))
            if tc:      # no pre-rounded operand copy: the kernel rounds h while staging, the skip stays unrounded fp32
                hraw = h.cuda()
                _lib.check(lib.gnngls_ff_forward(impl, p(hraw), None, M, p(W1c), p(b1c), p(W2c), p(b2c), p(scc), p(shc), p(out),
                                                 None, _ops.FT_TF32, p(ws), nbytes, _ops._stream()))
                o64b = (h.double() + hid @ W2d.double().t() + b2.double()) * sc.double() + sh.double()
                assert (out.cpu().double() - o64b).abs().max() < 1e-3, (M, 'round-in-kernel')
                # kind::f16 variant: fp16 weights, operands packed to fp16 in the kernel, fp32 skip
                W1h, W2h = W1.half().cuda(), W2.half().cuda()
                out.fill_(float('nan'))
                out_h = torch.empty(M, 128, device='cuda', dtype=torch.float16)
                _lib.check(lib.gnngls_ff_forward(_ops.DENSE_TCGEN05_F16, p(hraw), None, M, p(W1h), p(b1c), p(W2h), p(b2c), p(scc),
                                                 p(shc), p(out), p(out_h), _ops.FT_F16, p(ws), 0, _ops._stream()))
                torch.cuda.synchronize()
                assert torch.equal(out_h.cpu(), out.cpu().half())           # fp16 operand copy for the next fc
                h16, W116, W216 = h.half().double(), W1.half().double(), W2.half().double()
                hid16 = torch.relu(h16 @ W116.t() + b1.double()).half().double()
                o64h = (h.double() + hid16 @ W216.t() + b2.double()) * sc.double() + sh.double()
                assert (out.cpu().double() - o64h).abs().max() < 1e-3, (M, 'f16 feed-forward')
                # kind::f16 fc: fp16 input copy and fp16 weights
                hh16, Wh16 = h.half().cuda(), W.half().cuda()
                ft16 = torch.empty(M, 128, device='cuda', dtype=torch.float16)
                el.fill_(float('nan')); er.fill_(float('nan'))
                _lib.check(lib.gnngls_fc_forward(_ops.DENSE_TCGEN05_F16, p(hh16), M, p(Wh16), p(alc), p(arc), p(ft16), _ops.FT_F16,
                                                 p(el), p(er), _ops._stream()))
                torch.cuda.synchronize()
                f64 = h.half().double() @ W.half().double().t()
                e64 = (f64.view(M, 8, 16) * al.double().view(1, 8, 16)).sum(-1)
                r64 = (f64.view(M, 8, 16) * ar.double().view(1, 8, 16)).sum(-1)
                assert (ft16.cpu().double() - f64).abs().max() < 5e-3, (M, 'f16 fc')
                assert (el.cpu().double() - LOG2E * e64).abs().max() < 2e-4 and (er.cpu().double() - LOG2E * r64).abs().max() < 2e-4


def test_glue_kernels_bit_exact():
    from sklearn.preprocessing import MinMaxScaler
    rng = np.random.default_rng(2)
    B, n = 7, 23
    _, D = instances.random_instances(B, n, seed=8)
    fs = MinMaxScaler().fit(np.array([[0.003], [1.37]]))
    rs = MinMaxScaler().fit(np.array([[0.0], [0.31]]))
    x = _ops.edge_features(torch.as_tensor(D).cuda(), float(fs.scale_[0]), float(fs.min_[0])).cpu().numpy()
    ref = fs.transform(instances.edge_features(D).reshape(-1, 1)).reshape(B, -1)
    assert x.dtype == np.float32 and np.array_equal(x, ref)
    y = (rng.random((B, n * (n - 1) // 2)).astype(np.float32) - np.float32(0.3))
    r = _ops.regret_postprocess(torch.as_tensor(y).cuda(), float(rs.scale_[0]), float(rs.min_[0])).cpu().numpy()
    ref = np.maximum(rs.inverse_transform(y.reshape(-1, 1).copy()), 0).reshape(B, -1)
    assert np.array_equal(r, ref.astype(np.float32))


def test_pipeline_tours_bit_exact_given_gpu_regrets():
    """End to end (scripts/test.py:72-95): GPU regrets -> the CPU oracle's NN+GLS must produce exactly the GPU tours."""
    from gnngls_b200 import pipeline
    from oracle import gls_port
    _, m = make_models()
    n, B, K = 20, 24, 5
    _, D = instances.random_instances(B, n, seed=6)
    solver = pipeline.RegretGLS(m, micro_batch=7)
    solver.calibrate_synthetic_regret_scaler(torch.as_tensor(D).cuda())
    res = solver.solve(torch.as_tensor(D).cuda(), n_iters=K, perturbation_moves=20, keep_regret=True)
    regret = res.regret.cpu().numpy()
    assert regret.min() >= 0 and 0.01 < (regret == 0).mean() < 0.2 and regret.max() < 1.0
    o_t, o_c = gls_port.pipeline_batch(D, regret, K, 20, nthreads=4)
    assert np.array_equal(res.best_tours.cpu().numpy(), o_t)
    assert np.array_equal(_golden.bits(res.best_costs.cpu().numpy()), _golden.bits(o_c))
    t_h, c_h = solver.solve_host(D, chunk=10, n_iters=K, perturbation_moves=20)
    assert np.array_equal(t_h, o_t) and np.array_equal(_golden.bits(c_h), _golden.bits(o_c))


@pytest.mark.parametrize('n,B', [(3, 4), (4, 3), (7, 2), (33, 2), (100, 1)])
def test_model_sizes_vs_fp64_oracle(n, B):
    """Smallest legal graphs (n=3: every node has 2 in-edges), odd sizes that exercise every padding path of the
    K_n kernel (n not a multiple of 4/8/16) and one full-size TSP100 instance."""
    port, m = make_models()
    _, D = instances.random_instances(B, n, seed=100 + n)
    x = (instances.edge_features(D) / np.float32(np.sqrt(2))).reshape(-1, 1)
    with torch.no_grad():
        y64 = port.double()(model_port.EdgeListGraph.kn_line_graph(n, B), torch.as_tensor(x).double()).numpy()
    port.float()
    tol = tf32_budget(port, n, B, x, y64)
    for dense, gat in (('tcgen05', 'kn'), ('tcgen05', 'csr')):
        err = np.abs(run(m, n, B, x, dense, gat) - y64).max()
        print(f'n={n} B={B} {dense}+{gat}: err={err:.3e} tol={tol:.3e}')
        assert err <= tol
    assert np.abs(run(m, n, B, x, 'simt', 'csr') - y64).max() <= REL_FP32 * np.abs(y64).max()


def test_public_layer_and_gatconv_forward():
    """AttentionLayer.forward(G, x) and GATConv.forward(G, feat) keep the reference call signatures (models.py:38,12)."""
    port, m = make_models()
    n, B = 9, 2
    G = graph.LineGraph.complete(n, B, 'cuda')
    g = model_port.EdgeListGraph.kn_line_graph(n, B)
    torch.manual_seed(3)
    h = torch.randn(G.number_of_nodes(), 128)
    layer, player = m.message_passing_layers[2], port.message_passing_layers[2]
    with torch.no_grad():
        want = player.double()(g, h.double()).float()
        got = layer(G, h.cuda()).cpu()
        gat_want = player.message_passing.module(g, h.double()).float()
        gat_got = layer.message_passing.module(G, h.cuda()).cpu()
    port.float()
    assert got.shape == want.shape and (got - want).abs().max() < 2e-2 * want.abs().max()
    assert gat_got.shape == gat_want.shape == (G.number_of_nodes(), 8, 16)
    assert (gat_got - gat_want).abs().max() < 2e-2 * gat_want.abs().max()


@pytest.mark.parametrize('n,B', [(3, 2), (4, 3), (5, 1), (8, 2), (9, 2), (15, 1), (16, 2), (17, 1), (24, 1), (25, 2),
                                 (32, 1), (33, 1), (40, 1), (57, 1), (100, 1)])
def test_aggregate_kernels_all_storage_formats(n, B):
    """gnngls_gat_aggregate_kn / _csr with fp32, TF32 and fp16 feature storage against an fp64 evaluation of the same
    softmax-aggregate on inputs that are exactly representable in fp16 (so only the attention weights round).
    Sizes straddle the chunk boundaries of the K_n kernel's scans and both of its sort capacities (64, 128 slots)."""
    from gnngls_b200 import _lib
    lib = _lib.load()
    p = _ops._ptr
    g = torch.Generator().manual_seed(100 * n + B)
    N = n * (n - 1) // 2
    M = B * N
    ft = (torch.randn(M, 128, generator=g) * 2).half().float()
    el, er = torch.randn(M, 8, generator=g) * 3, torch.randn(M, 8, generator=g) * 3       # log2-domain scores
    # a few dominant sources (far above every other member of their stars): the K_n kernel then takes its
    # exact-row path for the destination that is itself the arg-max member
    hot = torch.randint(0, M, (max(1, M // 7),), generator=g)
    el[hot, torch.randint(0, 8, (len(hot),), generator=g)] += 25.0
    el[hot[0]] += 60.0
    h = torch.randn(M, 128, generator=g)
    bias = torch.randn(128, generator=g) * 0.1
    sc, sh = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1
    # fp64 reference (dst-sorted, constant in-degree)
    s, _ = model_port.kn_line_graph_edges(n)
    deg = 2 * (n - 2)
    src = torch.as_tensor(s).view(N, deg)
    src = (src[None] + (torch.arange(B) * N)[:, None, None]).reshape(M, deg)
    e = el.double()[src] + er.double()[:, None, :]
    e = torch.maximum(e, 0.2 * e)
    a = torch.softmax(e * np.log(2.0), dim=1)                                              # 2^e normalised
    agg = torch.einsum('mdh,mdhf->mhf', a, ft.double()[src].view(M, deg, 8, 16)).reshape(M, 128)
    ref = (h.double() + agg + bias.double()) * sc.double() + sh.double()
    G = graph.LineGraph.complete(n, B, 'cuda')
    indptr, indices = G.csr()
    elc, erc, hc, bc, scc, shc = el.cuda(), er.cuda(), h.cuda(), bias.cuda(), sc.cuda(), sh.cuda()
    nbytes = lib.gnngls_gat_kn_workspace_bytes(B, n)
    wk = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    scale = float(ft.abs().max())
    for ft_dtype in (_ops.FT_F32, _ops.FT_TF32, _ops.FT_F16):
        ftc = ft.cuda().half() if ft_dtype == _ops.FT_F16 else ft.cuda()
        for kind in ('csr', 'kn'):
            out = torch.full((M, 128), float('nan'), device='cuda')
            if kind == 'csr':
                _lib.check(lib.gnngls_gat_aggregate_csr(p(indptr), p(indices), M, p(ftc), ft_dtype, p(elc), p(erc), p(hc),
                                                        p(bc), p(scc), p(shc), p(out), None, _ops._stream()))
                tol = 2e-5 * scale                    # fp32 arithmetic throughout
            else:
                _lib.check(lib.gnngls_gat_aggregate_kn(B, n, p(ftc), ft_dtype, p(elc), p(erc), p(hc), p(bc), p(scc),
                                                       p(shc), p(out), None, p(wk), nbytes, _ops._stream()))
                # fp32 / TF32 storage (and n > 128): sorted-prefix kernel, fp32 arithmetic throughout.  fp16 storage, n <= 128: the
                # tcgen05 kernel, whose operand rows (weight x feature) are rounded to fp16 like the attention weights of the
                # fp16 tensor-core path always were
                tol = (1.5e-3 if ft_dtype == _ops.FT_F16 else 3e-5) * scale
            torch.cuda.synchronize()
            err = (out.cpu().double() - ref).abs().max().item()
            assert np.isfinite(err) and err < tol, (n, B, ft_dtype, kind, err, tol)


def test_tf32_feature_storage_switch(monkeypatch):
    """GNNGLS_FT_DTYPE=tf32 / GNNGLS_FF_DTYPE=tf32 keep the all-TF32 tensor-core kernels reachable from the model."""
    port, m = make_models()
    c = MODEL[0]
    n, B = c.nB.tolist()
    y16 = run(m, n, B, c.x, 'tcgen05', 'kn')
    monkeypatch.setenv('GNNGLS_FT_DTYPE', 'tf32')
    monkeypatch.setenv('GNNGLS_OP_DTYPE', 'tf32')
    y32 = run(m, n, B, c.x, 'tcgen05', 'kn')
    assert not np.array_equal(y16, y32)
    tol = tf32_budget(port, n, B, c.x, c.y64)
    assert np.abs(y32 - c.y64).max() <= tol and np.abs(y16 - c.y64).max() <= tol


@pytest.mark.parametrize('n,ft_dtype', [(128, 'f16'), (129, 'f16'), (150, 'f32'), (200, 'f16'), (256, 'f16'), (257, 'f32'),
                                        (380, 'f16'), (500, 'f16'), (512, 'f32'), (513, 'f16'), (700, 'f16'), (1000, 'f16')])
def test_large_n_kn_kernel_vs_fp64_rows(n, ft_dtype):
    """Above the sizes the torch oracle handles in seconds (every shared-memory configuration of csrc/gat_kn.cu: 4, 2
    and 1 heads per CTA, 4..32 sort elements per lane): 400 sampled destination rows against an fp64 evaluation with
    arithmetic adjacency (tests/_kn_ref.py).  Features are fp16-representable, so only fp32 arithmetic separates them."""
    from gnngls_b200 import _lib
    from tests import _kn_ref
    lib = _lib.load()
    p = _ops._ptr
    g = torch.Generator().manual_seed(n)
    B = 2 if n <= 257 else 1
    N = n * (n - 1) // 2
    M = B * N
    ft = (torch.randn(M, 128, generator=g) * 2).half()
    el, er = torch.randn(M, 8, generator=g) * 3, torch.randn(M, 8, generator=g) * 3
    hot = torch.randint(0, M, (max(1, M // 50),), generator=g)          # dominant members: exercises the exact arg-max row
    el[hot, torch.randint(0, 8, (len(hot),), generator=g)] += 30.0
    h = torch.randn(M, 128, generator=g)
    bias = torch.randn(128, generator=g) * 0.1
    sc, sh = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1
    f16 = ft_dtype == 'f16'
    ftc = ft.cuda() if f16 else ft.float().cuda()
    elc, erc, hc, bc, scc, shc = el.cuda(), er.cuda(), h.cuda(), bias.cuda(), sc.cuda(), sh.cuda()
    nbytes = lib.gnngls_gat_kn_workspace_bytes(B, n)
    wk = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    out = torch.full((M, 128), float('nan'), device='cuda')
    _lib.check(lib.gnngls_gat_aggregate_kn(B, n, p(ftc), _ops.FT_F16 if f16 else _ops.FT_F32, p(elc), p(erc), p(hc), p(bc),
                                           p(scc), p(shc), p(out), None, p(wk), nbytes, _ops._stream()))
    torch.cuda.synchronize()
    out = out.cpu()
    assert torch.isfinite(out).all()
    rng = np.random.default_rng(n)
    rows = np.unique(np.concatenate([rng.integers(0, M, 380), hot[:20].numpy(), [0, N - 1, M - 1]]))
    ref = _kn_ref.aggregate_rows(n, rows, ft.float().numpy(), el.numpy(), er.numpy(), h.numpy(), bias.numpy(), sc.numpy(), sh.numpy())
    err = np.abs(out.numpy()[rows].astype(np.float64) - ref).max()
    # n <= 128 with fp16 storage runs the tcgen05 kernel (operand rows rounded to fp16); everything else is fp32 arithmetic
    tol = (1.5e-3 if (f16 and n <= 128) else 3e-5) * float(ft.abs().max())
    assert err < tol, (n, err, tol)


def test_operand_range_fp16_vs_tf32_mode():
    """fp16 operands share TF32's 10-bit mantissa but not its exponent: activations above 65504 saturate and activations
    below 6e-5 lose mantissa bits.  Scale the embedding so that the residual stream sits (a) far above and (b) far below
    fp16's range: `operand_dtype='tf32'` (wide exponent, fp32 feature storage, exact K_n kernel) must stay inside the TF32
    budget in both cases; the default fp16 mode must stay finite, and is expected to leave the budget -- which is the
    documented reason the mode exists (README / INTEGRATION.md)."""
    n, B = 12, 2
    N = n * (n - 1) // 2
    x = np.random.default_rng(3).random((B * N, 1)).astype(np.float32)
    for scale, label in ((3e5, 'above fp16 range'), (1e-6, 'below fp16 normal range')):
        port, m = make_models()
        with torch.no_grad():                                   # eval-mode BN with running stats is affine: the scale survives
            port.embed_layer.weight.mul_(scale); port.embed_layer.bias.mul_(scale)
            for l in port.message_passing_layers:               # keep the attention logits O(1) so the softmax stays informative
                l.message_passing.module.attn_l.div_(scale); l.message_passing.module.attn_r.div_(scale)
        m.load_state_dict(port.state_dict(), strict=True)
        with torch.no_grad():
            y64 = port.double()(model_port.EdgeListGraph.kn_line_graph(n, B), torch.as_tensor(x).double()).numpy()
        port.float()
        tol = tf32_budget(port, n, B, x, y64)
        m.operand_dtype = 'tf32'
        e_tf32 = np.abs(run(m, n, B, x, 'tcgen05', 'kn') - y64).max()
        m.operand_dtype = 'f16'
        y16 = run(m, n, B, x, 'tcgen05', 'kn')
        e_f16 = np.abs(y16 - y64).max()
        print(f'{label}: |y|max={np.abs(y64).max():.3e} tol={tol:.3e} err tf32 mode={e_tf32:.3e} fp16 mode={e_f16:.3e}')
        assert e_tf32 <= tol, (label, e_tf32, tol)
        assert np.isfinite(y16).all(), label


@pytest.mark.parametrize('n,B,mode', [(20, 100, 'f16'), (20, 100, 'tf32'), (50, 48, 'f16'), (100, 24, 'f16'), (100, 24, 'tf32')])
def test_optimality_statistically_unchanged(n, B, mode):
    """North-star correctness leg 3 (/root/reference/scripts/test.py:59-104): the same instances solved end to end on the CPU
    (oracle: fp32 torch model + C port of nearest_neighbor / guided_local_search) and on the GPU, EACH SIDE WITH ITS OWN
    predicted regrets.  Tours differ where near-ties of the (random-init, i.e. noise) guide flip; the best costs must be
    statistically indistinguishable: |mean paired difference| <= 3 standard errors, or below 0.1 % of the mean cost."""
    from gnngls_b200 import pipeline
    from oracle import gls_port
    gls_port.build()
    torch.manual_seed(0)
    port = model_port.EdgeModelPort(1, 128, 1, 3, n_heads=8).eval()
    m = models.EdgePropertyPredictionModel(1, 128, 1, 3, n_heads=8)
    m.load_state_dict(port.state_dict(), strict=True)
    m = m.cuda().eval()
    m.operand_dtype = mode
    solver = pipeline.RegretGLS(m, micro_batch=64)
    _, D = instances.random_instances(B, n, seed=77)
    _, Dcal = instances.random_instances(8, n, seed=76)
    s = solver.calibrate_synthetic_regret_scaler(torch.from_numpy(Dcal).cuda())
    K, pm = 10, 20
    res = solver.solve(torch.from_numpy(D).cuda(), n_iters=K, perturbation_moves=pm)
    g_cost = res.best_costs.cpu().numpy()
    # CPU side: features -> oracle model -> inverse scale / clamp -> C port of the search, same scalers
    x = instances.edge_features(D)
    x = ((x.astype(np.float64) * s.feat_scale).astype(np.float32).astype(np.float64) + s.feat_min).astype(np.float32)
    g1 = model_port.EdgeListGraph.kn_line_graph(n, 1)
    regret = np.empty((B, n * (n - 1) // 2), dtype=np.float32)
    with torch.no_grad():
        for b in range(B):
            y = port(g1, torch.from_numpy(x[b]).reshape(-1, 1)).numpy().reshape(-1)
            r = ((y.astype(np.float64) - s.regret_min).astype(np.float32).astype(np.float64) / s.regret_scale).astype(np.float32)
            regret[b] = np.maximum(r, 0)
    import os
    _, c_cost = gls_port.pipeline_batch(D, regret, K, pm, nthreads=os.cpu_count() or 1)
    d = g_cost - c_cost
    se = d.std(ddof=1) / np.sqrt(B)
    rel = abs(d.mean()) / c_cost.mean()
    print(f'TSP{n} x {B} [{mode}]: CPU mean {c_cost.mean():.5f} GPU mean {g_cost.mean():.5f} paired diff {d.mean():+.5f} '
          f'+- {1.96 * se:.5f} (95 %), {100 * rel:.3f} % of the mean; identical tours costs: {(d == 0).mean():.2f}')
    assert abs(d.mean()) <= 3 * se or rel < 1e-3, (n, B, mode, d.mean(), se)


def test_fused_model_forward_entry_matches_per_op_path(monkeypatch):
    """gnngls_model_forward (one C-ABI call for the whole forward) issues the same kernels on the same buffers' contents as
    the per-op Python loop: bitwise-identical outputs, in the fp16, TF32 and fp32 operand modes."""
    n, B = 17, 5
    N = n * (n - 1) // 2
    x = np.random.default_rng(11).random((B * N, 1)).astype(np.float32)
    port, m = make_models(gat_bias=True)
    for dense, mode in (('tcgen05', 'f16'), ('tcgen05', 'tf32'), ('simt', None)):
        m.operand_dtype = mode
        monkeypatch.setenv('GNNGLS_MODEL_FORWARD', 'per_op')
        y_ops = run(m, n, B, x, dense, 'kn')
        monkeypatch.setenv('GNNGLS_MODEL_FORWARD', 'fused')
        y_fused = run(m, n, B, x, dense, 'kn')
        assert np.array_equal(y_ops, y_fused), (dense, mode, np.abs(y_ops - y_fused).max())
