#!/usr/bin/env python
"""Time gnngls_gat_aggregate_kn alone (no oracle): python tools/kn_bench.py [n] [B] [reps] [f16|f32]"""
import sys

import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from gnngls_b200 import _lib, _ops  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    f16 = (sys.argv[4] if len(sys.argv) > 4 else 'f16') == 'f16'
    lib = _lib.load()
    p = _ops._ptr
    g = torch.Generator().manual_seed(n)
    M = B * n * (n - 1) // 2
    ft = (torch.randn(M, 128, generator=g) * 2).half()
    el, er = torch.randn(M, 8, generator=g) * 3, torch.randn(M, 8, generator=g) * 3
    h = torch.randn(M, 128, generator=g)
    sc, sh = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1
    ftc = ft.cuda() if f16 else ft.float().cuda()
    elc, erc, hc, scc, shc = el.cuda(), er.cuda(), h.cuda(), sc.cuda(), sh.cuda()
    nbytes = lib.gnngls_gat_kn_workspace_bytes(B, n)
    wk = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    out = torch.empty(M, 128, device='cuda')

    def launch():
        _lib.check(lib.gnngls_gat_aggregate_kn(B, n, p(ftc), _ops.FT_F16 if f16 else _ops.FT_F32, p(elc), p(erc), p(hc), None,
                                               p(scc), p(shc), p(out), None, p(wk), nbytes, _ops._stream()))
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    print(f'n={n} B={B} {"f16" if f16 else "f32"}: {e0.elapsed_time(e1) / reps:.4f} ms per launch', flush=True)


if __name__ == '__main__':
    main()
