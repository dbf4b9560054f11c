"""Drop-in for /root/reference/gnngls/models.py on B200.

Same class names, constructor arguments, ``forward(G, x)`` signature and module tree, so a
``state_dict`` saved by the reference (scripts/train.py:60-67) loads with ``strict=True`` — including
checkpoints trained with DGL >= 0.7 whose GATConv carries an extra ``bias``.  The forward pass is
inference only (eval-mode BatchNorm) and runs entirely in hand-written sm_100a kernels through the
C ABI (include/gnngls_b200.h); per AttentionLayer (models.py:38-41):

    fc (+el/er)  ->  edge-softmax/aggregate + skip + BN1  ->  FF + skip + BN2

There is no CPU or eager-PyTorch fallback: tensors must live on a CUDA device.
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib, _ops
from ._timing import stage
from .graph import LineGraph

HIDDEN_DIM = 512        # hard-coded in the reference (models.py:60)
_SUPPORTED = dict(embed_dim=128, n_heads=8)


def _dense_impl_default():
    name = os.environ.get('GNNGLS_DENSE_IMPL', 'tcgen05').lower()
    if name not in ('tcgen05', 'simt'):
        raise ValueError('GNNGLS_DENSE_IMPL must be tcgen05 or simt')
    return name


class SkipConnection(nn.Module):
    """models.py:5-15.  Kept for the module tree; AttentionLayer fuses the add into its kernels."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, x, G=None):
        if G is not None:
            y = self.module(G, x).view(G.number_of_nodes(), -1)
        else:
            y = self.module(x)
        return x + y


class GATConv(nn.Module):
    """Parameter holder + standalone forward with dgl.nn.GATConv(in, out, heads) semantics
    (feat_drop=attn_drop=0, negative_slope=0.2, no residual, no activation)."""

    def __init__(self, in_feats, out_feats, num_heads, negative_slope=0.2):
        super().__init__()
        self._in_feats, self._out_feats, self._num_heads = in_feats, out_feats, num_heads
        if negative_slope != 0.2:
            raise NotImplementedError('kernels are built for negative_slope=0.2')
        self.fc = nn.Linear(in_feats, out_feats * num_heads, bias=False)
        self.attn_l = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.attn_r = nn.Parameter(torch.empty(1, num_heads, out_feats))
        self.bias = None      # DGL 0.6.1 (the pinned version) has no bias; created on load if present
        gain = nn.init.calculate_gain('relu')
        nn.init.xavier_normal_(self.fc.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        nn.init.xavier_normal_(self.attn_r, gain=gain)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        if prefix + 'bias' in state_dict and self.bias is None:
            self.bias = nn.Parameter(torch.zeros(self._out_feats * self._num_heads, device=self.fc.weight.device,
                                                 dtype=self.fc.weight.dtype))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def forward(self, G, feat):
        M = G.number_of_nodes()
        if not feat.is_cuda:
            raise RuntimeError('gnngls_b200 has no CPU path: move the layer, graph and features to a CUDA device')
        if feat.dim() != 2 or tuple(feat.shape) != (M, self._in_feats) or self._in_feats != 128:
            raise RuntimeError(f'GATConv expects [{M}, 128] features; got {tuple(feat.shape)}')
        with torch.cuda.device(feat.device):                  # launches go to the stream of the tensor's device
            return self._forward(G, feat, M)

    def _forward(self, G, feat, M):
        zero = torch.zeros(M, self._in_feats, dtype=torch.float32, device=feat.device)
        one = torch.ones(self._in_feats, dtype=torch.float32, device=feat.device)
        tc = _dense_impl_default() != 'simt'
        feat = feat.detach().float().contiguous()
        W = self.fc.weight.detach().float().contiguous()
        out, _ = _gat_block(G, tf32_round(feat) if tc else feat, zero, tf32_round(W) if tc else W,
                            self.attn_l.detach().reshape(-1).contiguous(), self.attn_r.detach().reshape(-1).contiguous(),
                            None if self.bias is None else self.bias.detach().contiguous(), one,
                            torch.zeros_like(one), _ops.DENSE_TCGEN05 if tc else _ops.DENSE_SIMT, 'auto', {})
        return out.clone().view(M, self._num_heads, self._out_feats)


def tf32_round(t):
    """Round fp32 to TF32 (10-bit mantissa), nearest with ties away from zero == PTX cvt.rna.tf32.f32.
    The tcgen05 kind::tf32 MMA ignores the low 13 mantissa bits of its operands; weights are rounded
    here once so that this truncation is exact (activations are rounded by the producing kernels)."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


_KN_MAX_N = 1024      # largest vertex star the K_n kernel holds in one CTA's shared memory (csrc/gat_kn.cu)


_OPERAND_OVERRIDE = []      # innermost EdgePropertyPredictionModel.operand_dtype while its forward runs


def _op_f16():
    """Tensor-core path: operand copies / weights travel as fp16 unless the model's `operand_dtype` attribute,
    GNNGLS_OP_DTYPE=tf32 (or the older GNNGLS_FF_DTYPE=tf32) asks for the all-TF32 kernels."""
    if _OPERAND_OVERRIDE and _OPERAND_OVERRIDE[-1] is not None:
        return _OPERAND_OVERRIDE[-1] != 'tf32'
    return (os.environ.get('GNNGLS_OP_DTYPE', 'f16').lower() != 'tf32'
            and os.environ.get('GNNGLS_FF_DTYPE', 'f16').lower() != 'tf32')


def _bn_affine(bn):
    """Eval-mode BatchNorm1d as y = x*scale + shift (models.py:27,35)."""
    with torch.no_grad():
        scale = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
        shift = bn.bias.float() - bn.running_mean.float() * scale
    return scale.contiguous(), shift.contiguous()


def _buf(ws, name, shape, dtype, device):
    t = ws.get(name)
    numel = 1
    for s in shape:
        numel *= s
    if t is None or t.numel() < numel or t.dtype != dtype or t.device != device:
        t = torch.empty(numel, dtype=dtype, device=device)
        ws[name] = t
    return t[:numel].view(*shape)


def _gat_block(G, h_op, skip, Wfc, al, ar, bias, bn_scale, bn_shift, dense_impl, gat_impl, ws):
    """fc + aggregate(+skip+BN1).  `h_op` is the fc operand (the TF32-rounded copy on the tensor-core
    path), `skip` the unrounded activation.  Returns (h1, h1_tf32 or None), both [M,128]."""
    lib = _lib.load()
    M, dev = h_op.shape[0], h_op.device
    st = _ops._stream()
    tc = dense_impl in (_ops.DENSE_TCGEN05, _ops.DENSE_TCGEN05_F16)
    # tensor-core path: ft travels as fp16 (same mantissa as the TF32 operand it would be rounded to, half the
    # bytes); GNNGLS_FT_DTYPE=tf32 keeps fp32 storage.  The fp32 debug path keeps ft exact.
    if not tc:
        ft_dtype = _ops.FT_F32
    elif os.environ.get('GNNGLS_FT_DTYPE', 'f16').lower() == 'tf32' or (_OPERAND_OVERRIDE and _OPERAND_OVERRIDE[-1] == 'tf32'):
        ft_dtype = _ops.FT_TF32
    else:
        ft_dtype = _ops.FT_F16
    ft = _buf(ws, 'ft', (M, 128), torch.float16 if ft_dtype == _ops.FT_F16 else torch.float32, dev)
    el = _buf(ws, 'el', (M, 8), torch.float32, dev)
    er = _buf(ws, 'er', (M, 8), torch.float32, dev)
    h1 = _buf(ws, 'h1', (M, 128), torch.float32, dev)
    h1r = None      # the fused FF kernel rounds h1 to TF32 itself while staging it into tensor memory
    p = _ops._ptr
    with stage('fc'):
        _lib.check(lib.gnngls_fc_forward(dense_impl, p(h_op), M, p(Wfc), p(al), p(ar), p(ft), ft_dtype, p(el), p(er), st))
    use_kn = gat_impl == 'kn' or (gat_impl == 'auto' and G.kind == 'kn' and G.n <= _KN_MAX_N)
    if use_kn:
        if G.kind != 'kn':
            raise ValueError("gat_impl='kn' needs a LineGraph.complete graph")
        nbytes = lib.gnngls_gat_kn_workspace_bytes(G.batch_size, G.n)
        wk = _buf(ws, 'gat_ws', (nbytes,), torch.uint8, dev)
        with stage('gat_kn'):
            _lib.check(lib.gnngls_gat_aggregate_kn(G.batch_size, G.n, p(ft), ft_dtype, p(el), p(er), p(skip), p(bias),
                                                   p(bn_scale), p(bn_shift), p(h1), p(h1r), p(wk), nbytes, st))
    else:
        indptr, indices = G.csr()
        with stage('gat_csr'):
            _lib.check(lib.gnngls_gat_aggregate_csr(p(indptr), p(indices), M, p(ft), ft_dtype, p(el), p(er), p(skip), p(bias),
                                                    p(bn_scale), p(bn_shift), p(h1), p(h1r), st))
    return h1, h1r


class AttentionLayer(nn.Module):
    """models.py:18-41."""

    def __init__(self, embed_dim, n_heads, hidden_dim):
        super().__init__()
        self.message_passing = SkipConnection(GATConv(embed_dim, embed_dim // n_heads, n_heads))
        self.feed_forward = nn.Sequential(
            nn.BatchNorm1d(embed_dim),
            SkipConnection(nn.Sequential(nn.Linear(embed_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, embed_dim))),
            nn.BatchNorm1d(embed_dim),
        )
        self._dims = (embed_dim, n_heads, hidden_dim)

    def _device_params(self):
        gat = self.message_passing.module
        ff = self.feed_forward[1].module
        s1, t1 = _bn_affine(self.feed_forward[0])
        s2, t2 = _bn_affine(self.feed_forward[2])
        f = lambda t: t.detach().float().contiguous()      # noqa: E731
        d = dict(Wfc=f(gat.fc.weight), al=f(gat.attn_l).reshape(-1), ar=f(gat.attn_r).reshape(-1),
                 gbias=None if gat.bias is None else f(gat.bias), s1=s1, t1=t1,
                 W1=f(ff[0].weight), b1=f(ff[0].bias), W2=f(ff[2].weight), b2=f(ff[2].bias), s2=s2, t2=t2)
        for k in ('Wfc', 'W1', 'W2'):
            d[k + '_tf32'] = tf32_round(d[k])
        for k in ('Wfc', 'W1', 'W2'):                       # fp16 variants (same 10-bit mantissa as TF32)
            d[k + '_f16'] = d[k].clamp(-65504.0, 65504.0).to(torch.float16).contiguous()
        return d

    def forward(self, G, x, _params=None, _ws=None, _dense_impl=None, _gat_impl='auto', _out=None, _x_tf32=None,
                _out_tf32=None):
        """Public form: forward(G, x) -> [M,128].  The underscore arguments let the model reuse buffers and
        pass the TF32-rounded operand copies between layers."""
        if self.training:
            raise NotImplementedError('gnngls_b200 implements the inference path only: call model.eval()')
        if self._dims != (128, 8, 512):
            raise NotImplementedError('kernels are specialised for embed_dim=128, n_heads=8, hidden_dim=512')
        if not x.is_cuda:
            raise RuntimeError('gnngls_b200 has no CPU path: move the layer, graph and features to a CUDA device')
        if x.dim() != 2 or x.shape[1] != 128:
            raise RuntimeError(f'AttentionLayer expects [nodes, 128] features; got {tuple(x.shape)}')
        if x.shape[0] != G.number_of_nodes():
            raise ValueError(f'features have {x.shape[0]} rows but the graph has {G.number_of_nodes()} nodes')
        with torch.cuda.device(x.device):                      # launches go to the stream of the tensor's device
            return self._forward(G, x, _params, _ws, _dense_impl, _gat_impl, _out, _x_tf32, _out_tf32)

    def _forward(self, G, x, _params, _ws, _dense_impl, _gat_impl, _out, _x_tf32, _out_tf32):
        lib = _lib.load()
        prm = _params if _params is not None else self._device_params()
        ws = _ws if _ws is not None else {}
        impl = _dense_impl if _dense_impl is not None else (
            _ops.DENSE_SIMT if _dense_impl_default() == 'simt' else _ops.DENSE_TCGEN05)
        tc = impl == _ops.DENSE_TCGEN05
        f16 = tc and _op_f16()
        x = x.detach().to(torch.float32).contiguous()
        M, dev = x.shape[0], x.device
        sfx = '_f16' if f16 else ('_tf32' if tc else '')
        op_impl = _ops.DENSE_TCGEN05_F16 if f16 else impl
        op_dtype = _ops.FT_F16 if f16 else _ops.FT_TF32
        if tc and _x_tf32 is None:
            _x_tf32 = x.clamp(-65504.0, 65504.0).to(torch.float16) if f16 else tf32_round(x)
        h1, h1r = _gat_block(G, _x_tf32 if tc else x, x, prm['Wfc' + sfx], prm['al'], prm['ar'], prm['gbias'],
                             prm['s1'], prm['t1'], op_impl, _gat_impl, ws)
        nbytes = lib.gnngls_ff_workspace_bytes(op_impl, M)
        wk = _buf(ws, 'ff_ws', (nbytes,), torch.uint8, dev)
        out = _out if _out is not None else torch.empty(M, 128, dtype=torch.float32, device=dev)
        p = _ops._ptr
        with stage('ff'):
            _lib.check(lib.gnngls_ff_forward(op_impl, p(h1), p(h1r), M, p(prm['W1' + sfx]), p(prm['b1']),
                                             p(prm['W2' + sfx]), p(prm['b2']), p(prm['s2']), p(prm['t2']), p(out),
                                             p(_out_tf32), op_dtype, p(wk), nbytes, _ops._stream()))
        return out


class EdgePropertyPredictionModel(nn.Module):
    """models.py:44-70.  NB: like the reference, the number of AttentionLayers is ``n_heads``
    (models.py:60 iterates ``range(n_heads)``); ``n_layers`` is accepted and ignored."""

    def __init__(self, in_dim, embed_dim, out_dim, n_layers, n_heads=1):
        super().__init__()
        self.embed_dim = embed_dim
        self.embed_layer = nn.Linear(in_dim, embed_dim)
        self.message_passing_layers = nn.Sequential(
            *(AttentionLayer(embed_dim, n_heads, HIDDEN_DIM) for _ in range(n_heads)))
        self.decision_layer = nn.Linear(embed_dim, out_dim)
        self.dense_impl = None      # None -> $GNNGLS_DENSE_IMPL or 'tcgen05'
        self.gat_impl = 'auto'      # 'auto' | 'kn' | 'csr'
        # operands of the tensor-core contractions: None -> $GNNGLS_OP_DTYPE or 'f16'; 'tf32' = the wide-exponent mode for
        # checkpoints whose activations leave fp16's range (|h| > 65504 saturates, |h| < 6e-5 loses bits); feature storage
        # becomes fp32 too and the K_n aggregate runs its exact fp32 sorted-prefix kernel
        self.operand_dtype = None
        self._cache_sig, self._cache = None, None
        self._ws = {}

    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _prepared(self):
        sig = self._signature()
        if sig != self._cache_sig:
            f = lambda t: t.detach().float().contiguous()  # noqa: E731
            self._cache = dict(We=f(self.embed_layer.weight), be=f(self.embed_layer.bias),
                               Wd=f(self.decision_layer.weight), bd=f(self.decision_layer.bias),
                               layers=[l._device_params() for l in self.message_passing_layers])
            self._cache_sig = sig
        return self._cache

    def forward(self, G, x):
        if self.training:
            raise NotImplementedError('gnngls_b200 implements the inference path only: call model.eval()')
        if not isinstance(G, LineGraph):
            raise TypeError('G must be a gnngls_b200.LineGraph (LineGraph.complete(n, batch) or .from_edges(...))')
        if not x.is_cuda:
            raise RuntimeError('gnngls_b200 has no CPU path: move the model, graph and features to a CUDA device')
        n_heads = len(self.message_passing_layers)
        if self.embed_dim != _SUPPORTED['embed_dim'] or n_heads != _SUPPORTED['n_heads']:
            raise NotImplementedError('kernels are specialised for embed_dim=128, n_heads=8 (the shipped params.json)')
        lib = _lib.load()
        prm = self._prepared()
        name = self.dense_impl or _dense_impl_default()
        impl = _ops.DENSE_SIMT if name == 'simt' else _ops.DENSE_TCGEN05
        in_dim, out_dim = self.embed_layer.in_features, self.decision_layer.out_features
        if x.dim() != 2 or x.shape[1] != in_dim:              # what nn.Linear would raise; the kernel indexes x[m * in_dim + k]
            raise RuntimeError(f'features must be [nodes, {in_dim}] (embed_layer.in_features); got {tuple(x.shape)}')
        x = x.detach().to(torch.float32).contiguous()
        M, dev = x.shape[0], x.device
        if M != G.number_of_nodes():
            raise ValueError(f'features have {M} rows but the graph has {G.number_of_nodes()} nodes')
        p = _ops._ptr
        if self.operand_dtype not in (None, 'f16', 'tf32'):
            raise ValueError("operand_dtype must be None, 'f16' or 'tf32'")
        _OPERAND_OVERRIDE.append(self.operand_dtype)
        try:
            return self._forward_impl(G, x, lib, prm, impl, M, dev, in_dim, out_dim)
        finally:
            _OPERAND_OVERRIDE.pop()

    def _forward_fused(self, G, x, lib, prm, impl, M, dev, in_dim, out_dim):
        """Whole forward behind one C-ABI call (gnngls_model_forward): K_n line graphs, no per-stage timers."""
        import ctypes
        p = _ops._ptr
        tc = impl == _ops.DENSE_TCGEN05
        f16 = tc and _op_f16()
        sfx = '_f16' if f16 else ('_tf32' if tc else '')
        op_impl = _ops.DENSE_TCGEN05_F16 if f16 else impl
        if not tc:
            ft_dtype = _ops.FT_F32
        elif os.environ.get('GNNGLS_FT_DTYPE', 'f16').lower() == 'tf32' or (_OPERAND_OVERRIDE and _OPERAND_OVERRIDE[-1] == 'tf32'):
            ft_dtype = _ops.FT_TF32
        else:
            ft_dtype = _ops.FT_F16
        key = ('fused_layers', sfx)
        if key not in prm:
            arr = (_lib.LayerParams * len(prm['layers']))()
            for a, lp in zip(arr, prm['layers']):
                a.Wfc, a.attn_l, a.attn_r = p(lp['Wfc' + sfx]), p(lp['al']), p(lp['ar'])
                a.gat_bias = p(lp['gbias'])
                a.bn1_scale, a.bn1_shift = p(lp['s1']), p(lp['t1'])
                a.W1, a.b1, a.W2, a.b2 = p(lp['W1' + sfx]), p(lp['b1']), p(lp['W2' + sfx]), p(lp['b2'])
                a.bn2_scale, a.bn2_shift = p(lp['s2']), p(lp['t2'])
            prm[key] = arr
        with torch.cuda.device(dev):
            nbytes = lib.gnngls_model_forward_workspace_bytes(G.batch_size, G.n, op_impl)
            wk = _buf(self._ws, 'model_ws', (nbytes,), torch.uint8, dev)
            y = torch.empty(M, out_dim, dtype=torch.float32, device=dev)
            args = _lib.ModelArgs(G.batch_size, G.n, in_dim, out_dim, len(prm['layers']), op_impl, ft_dtype, 0, p(x), p(prm['We']),
                                  p(prm['be']), p(prm['Wd']), p(prm['bd']), prm[key], p(y))
            _lib.check(lib.gnngls_model_forward(ctypes.byref(args), p(wk), nbytes, _ops._stream()))
        return y

    def _forward_impl(self, G, x, lib, prm, impl, M, dev, in_dim, out_dim):
        from . import _timing
        if (G.kind == 'kn' and self.gat_impl in ('auto', 'kn') and G.n <= _KN_MAX_N and _timing._active is None
                and os.environ.get('GNNGLS_MODEL_FORWARD', 'fused') != 'per_op'):
            return self._forward_fused(G, x, lib, prm, impl, M, dev, in_dim, out_dim)
        p = _ops._ptr
        with torch.cuda.device(dev):
            tc = impl == _ops.DENSE_TCGEN05
            f16 = tc and _op_f16()
            op_t = torch.float16 if f16 else torch.float32     # operand copies of the activations for the fc GEMMs
            ha = _buf(self._ws, 'ha', (M, 128), torch.float32, dev)
            hb = _buf(self._ws, 'hb', (M, 128), torch.float32, dev)
            har = _buf(self._ws, 'ha_op', (M, 128), op_t, dev) if tc else None
            hbr = _buf(self._ws, 'hb_op', (M, 128), op_t, dev) if tc else None
            with stage('embed'):
                _lib.check(lib.gnngls_embed_forward(p(x), M, in_dim, p(prm['We']), p(prm['be']), p(ha), p(har),
                                                    _ops.FT_F16 if f16 else _ops.FT_TF32, _ops._stream()))
            cur, nxt, cur_r, nxt_r = ha, hb, har, hbr
            for layer, lp in zip(self.message_passing_layers, prm['layers']):
                layer(G, cur, _params=lp, _ws=self._ws, _dense_impl=impl, _gat_impl=self.gat_impl, _out=nxt,
                      _x_tf32=cur_r, _out_tf32=nxt_r)
                cur, nxt, cur_r, nxt_r = nxt, cur, nxt_r, cur_r
            y = torch.empty(M, out_dim, dtype=torch.float32, device=dev)
            with stage('decision'):
                _lib.check(lib.gnngls_decision_forward(p(cur), M, out_dim, p(prm['Wd']), p(prm['bd']), p(y),
                                                       _ops._stream()))
        return y
