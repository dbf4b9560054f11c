"""ctypes loader for libgnngls_b200.so.  There is NO fallback: if the CUDA library cannot be
loaded every op of this package raises."""
import ctypes
import os

from . import build as _build

_p = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_sz = ctypes.c_size_t
_d = ctypes.c_double


class GlsArgs(ctypes.Structure):
    """Mirror of struct gnngls_gls_args (include/gnngls_b200.h)."""
    _fields_ = [
        ('B', ctypes.c_int32), ('n', ctypes.c_int32),
        ('D', _p),
        ('guide_kind', ctypes.c_int32), ('n_guides', ctypes.c_int32),
        ('guides', _p),
        ('cur_tours', _p), ('cur_costs', _p), ('best_tours', _p), ('best_costs', _p),
        ('k', _p), ('penalties', _p),
        ('resume', ctypes.c_int32), ('iter_begin', ctypes.c_int32), ('n_iters', ctypes.c_int32),
        ('perturbation_moves', ctypes.c_int32), ('first_improvement', ctypes.c_int32),
        ('events', _p), ('n_events', _p), ('max_events', ctypes.c_int32),
        ('status', _p), ('counters', _p),
    ]


class LayerParams(ctypes.Structure):
    """Mirror of struct gnngls_layer_params (include/gnngls_b200.h)."""
    _fields_ = [('Wfc', _p), ('attn_l', _p), ('attn_r', _p), ('gat_bias', _p), ('bn1_scale', _p), ('bn1_shift', _p),
                ('W1', _p), ('b1', _p), ('W2', _p), ('b2', _p), ('bn2_scale', _p), ('bn2_shift', _p)]


class ModelArgs(ctypes.Structure):
    """Mirror of struct gnngls_model_args (include/gnngls_b200.h)."""
    _fields_ = [('B', ctypes.c_int32), ('n', ctypes.c_int32), ('in_dim', ctypes.c_int32), ('out_dim', ctypes.c_int32),
                ('n_layers', ctypes.c_int32), ('dense_impl', ctypes.c_int32), ('ft_dtype', ctypes.c_int32),
                ('reserved', ctypes.c_int32), ('x', _p), ('We', _p), ('be', _p), ('Wd', _p), ('bd', _p),
                ('layers', ctypes.POINTER(LayerParams)), ('y', _p)]


# name -> (restype, argtypes); every symbol declared in include/gnngls_b200.h
SIGNATURES = {
    'gnngls_abi_version': (_i, []),
    'gnngls_last_error_string': (ctypes.c_char_p, []),
    'gnngls_moves_eval_a2a': (_i, [_i, _p, _i64, _p, _i, _i, _i, _p, _p, _p, _p]),
    'gnngls_moves_eval_o2a': (_i, [_i, _p, _i64, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    'gnngls_local_search_batch': (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _i, _p, _p, _p]),
    'gnngls_gls_batch': (_i, [ctypes.POINTER(GlsArgs), _p]),
    'gnngls_sizeof_gls_args': (_sz, []),
    'gnngls_nn_init_batch': (_i, [_i, _p, _p, _i, _i, _i, _p, _p, _p]),
    'gnngls_tour_cost_batch': (_i, [_p, _p, _i, _i, _p, _p]),
    'gnngls_edge_features': (_i, [_p, _i, _i, _d, _d, _p, _p]),
    'gnngls_embed_forward': (_i, [_p, _i64, _i, _p, _p, _p, _p, _i, _p]),
    'gnngls_fc_forward': (_i, [_i, _p, _i64, _p, _p, _p, _p, _i, _p, _p, _p]),
    'gnngls_gat_aggregate_csr': (_i, [_p, _p, _i64, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    'gnngls_gat_kn_workspace_bytes': (_sz, [_i, _i]),
    'gnngls_gat_aggregate_kn': (_i, [_i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    'gnngls_ff_workspace_bytes': (_sz, [_i, _i64]),
    'gnngls_ff_forward': (_i, [_i, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    'gnngls_decision_forward': (_i, [_p, _i64, _i, _p, _p, _p, _p]),
    'gnngls_regret_postprocess': (_i, [_p, _i64, _d, _d, _p, _p]),
    'gnngls_sizeof_model_args': (_sz, []),
    'gnngls_model_forward_workspace_bytes': (_sz, [_i, _i, _i]),
    'gnngls_model_forward': (_i, [ctypes.POINTER(ModelArgs), _p, _sz, _p]),
}

# kernels launched per successful C-ABI call (bench.py reports the count as `gpu_launches`)
KERNELS_PER_CALL = {
    'gnngls_moves_eval_a2a': 1, 'gnngls_moves_eval_o2a': 1, 'gnngls_local_search_batch': 1, 'gnngls_gls_batch': 1,
    'gnngls_nn_init_batch': 1, 'gnngls_tour_cost_batch': 1, 'gnngls_edge_features': 1, 'gnngls_embed_forward': 1,
    'gnngls_fc_forward': 1, 'gnngls_gat_aggregate_csr': 1,
    'gnngls_gat_aggregate_kn': 1,
    # one fused tcgen05 kernel; the SIMT debug path (impl == 1, first argument) runs two GEMM kernels
    'gnngls_ff_forward': lambda args: 2 if args[0] == 1 else 1,
    'gnngls_decision_forward': 1, 'gnngls_regret_postprocess': 1,
    # embed + n_layers x (fc, aggregate, feed-forward) + decision
    'gnngls_model_forward': lambda args: 2 + 3 * args[0].contents.n_layers if hasattr(args[0], 'contents') else 2 + 3 * args[0]._obj.n_layers,
}


class _CountingLib:
    """Forwards to the ctypes library and counts the kernels our entry points launch."""

    def __init__(self, lib):
        self._lib = lib
        self.launches = 0

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        k = KERNELS_PER_CALL.get(name, 0)
        if k == 0:
            return fn

        def counted(*args):
            self.launches += k(args) if callable(k) else k
            return fn(*args)

        setattr(self, name, counted)
        return counted


_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first if the .so is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    # build() is incremental (mtime of every source and header against its object), so a stale library never survives an
    # edit of csrc/ or the header; the lock keeps the ranks of one torchrun from compiling into the same files at once.
    # Without nvcc (a box that only received the built library) the existing .so is used as it is.
    try:
        import fcntl
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path + '.lock', 'w') as lk:
            fcntl.flock(lk, fcntl.LOCK_EX)
            try:
                _build.build()
            finally:
                fcntl.flock(lk, fcntl.LOCK_UN)
    except Exception as e:  # noqa: BLE001
        if not os.path.exists(path):
            raise RuntimeError(
                f'libgnngls_b200.so is missing and could not be built ({e}); gnngls_b200 has no CPU fallback. '
                'Run `python -m gnngls_b200.build`.') from e
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:
        raise RuntimeError(f'cannot load {path}: {e}; gnngls_b200 has no CPU fallback') from e
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.gnngls_sizeof_gls_args() != ctypes.sizeof(GlsArgs):
        raise RuntimeError('GlsArgs ctypes mirror does not match struct gnngls_gls_args')
    if lib.gnngls_sizeof_model_args() != ctypes.sizeof(ModelArgs):
        raise RuntimeError('ModelArgs ctypes mirror does not match struct gnngls_model_args')
    if lib.gnngls_abi_version() != 1:
        raise RuntimeError('libgnngls_b200.so ABI version mismatch')
    _lib = _CountingLib(lib)
    return _lib


def check(rc):
    if rc != 0:
        msg = load().gnngls_last_error_string().decode('utf-8', 'replace')
        raise RuntimeError(f'gnngls_b200 error {rc}: {msg}')
