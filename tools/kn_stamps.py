#!/usr/bin/env python
"""Per-phase cycle breakdown of gat_kn_tc_kernel (library built with GNNGLS_KN_STAMPS=1):
    GNNGLS_KN_STAMPS=1 python -m gnngls_b200.build --force && python tools/kn_stamps.py [n] [B]"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from gnngls_b200 import _lib, _ops  # noqa: E402

NAMES = ['top A', 'scores + barrier 1', 'release/late/top-2', 'part B', 'barrier 2', 'operand row + indicator', 'barrier A + MMA issue',
         'two rows of prev', 'rest of the MMA', 'accumulators -> regs', 'barrier B', 'partial']


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    lib = _lib.load()
    p = _ops._ptr
    g = torch.Generator().manual_seed(n)
    M = B * n * (n - 1) // 2
    ft = (torch.randn(M, 128, generator=g) * 2).half().cuda()
    el, er = (torch.randn(M, 8, generator=g) * 3).cuda(), (torch.randn(M, 8, generator=g) * 3).cuda()
    h = torch.randn(M, 128, generator=g).cuda()
    sc, sh = (torch.rand(128, generator=g) + 0.5).cuda(), (torch.randn(128, generator=g) * 0.1).cuda()
    nbytes = lib.gnngls_gat_kn_workspace_bytes(B, n)
    wk = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    out = torch.empty(M, 128, device='cuda')
    for _ in range(3):
        _lib.check(lib.gnngls_gat_aggregate_kn(B, n, p(ft), _ops.FT_F16, p(el), p(er), p(h), None, p(sc), p(sh), p(out), None, p(wk), nbytes,
                                               _ops._stream()))
    torch.cuda.synchronize()
    from gnngls_b200 import build as _b
    raw = ctypes.CDLL(_b.LIB_PATH)
    buf = (ctypes.c_ulonglong * (148 * 4 * 16))()
    rc = raw.gnngls_debug_kn_stamps(buf, 148 * 4 * 16)
    assert rc == 0, 'library was not built with GNNGLS_KN_STAMPS=1'
    a = np.frombuffer(buf, dtype=np.uint64).reshape(148, 4, 16).astype(np.float64)
    iters = -(-B * n // 148)
    tot = a[:, :, :12].sum(-1).mean()
    print(f'n={n} B={B}: {iters} iterations, {tot / iters:.0f} cycles per iteration (mean over CTAs and teams)')
    for k, name in enumerate(NAMES):
        per = a[:, :, k].mean() / iters
        teams = '  '.join(f'{a[:, t, k].mean() / iters:6.0f}' for t in range(4))
        print(f'  {name:28s} {per:8.0f} cycles/iteration  {100 * a[:, :, k].mean() / tot:5.1f} %   per team: {teams}')
    print('  late rows per iteration (warp 0 of each team):', '  '.join(f'{a[:, t, 12].mean() / iters:.3f}' for t in range(4)))


if __name__ == '__main__':
    main()
