#!/usr/bin/env python
"""Per-phase cycle breakdown of gat_kn_tc_kernel (library built with GNNGLS_KN_STAMPS=1):
    GNNGLS_KN_STAMPS=1 python -m gnngls_b200.build --force && python tools/kn_stamps.py [n] [B]"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from gnngls_b200 import _lib, _ops, build  # noqa: E402

NAMES = ['top (bookkeeping, loads)', 'scores + barrier', 'slice setup, top-2', 'barrier', 'operand row + indicator', 'wait::st + fences',
         'barrier A', 'MMA issue', 'merge rows (shadow)', 'rest of the MMA', 'accumulators -> regs', 'barrier B', 'partial', 'end of star', '  (operand row part)', '  (merge: wait for landing)']


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
    raw = ctypes.CDLL(build.LIB_PATH)
    cnt = 148 * 4 * 4 * 16
    buf = (ctypes.c_ulonglong * cnt)()
    assert raw.gnngls_debug_kn_stamps(buf, cnt) == 0, 'library was not built with GNNGLS_KN_STAMPS=1'
    a = np.frombuffer(buf, dtype=np.uint64).reshape(148 * 4, 4, 16).astype(np.float64)
    stars = B * n / (148 * 4)
    tot = a[:, :, :16].sum(-1).mean()
    print(f'n={n} B={B}: {stars:.1f} stars per CTA, {tot / stars:.0f} cycles per star (mean over CTAs and warps)')
    for k, name in enumerate(NAMES):
        per = a[:, :, k].mean() / stars
        warps = '  '.join(f'{a[:, w, k].mean() / stars:6.0f}' for w in range(4))
        print(f'  {name:28s} {per:8.0f} cycles/star  {100 * a[:, :, k].mean() / tot:5.1f} %   per warp: {warps}')


if __name__ == '__main__':
    main()
