#!/usr/bin/env python
"""Config 5 of BASELINE.json: two_opt_a2a / relocate_a2a move-evaluation sweep, n = 20..1000, batch up to 65,536.

For every (n, batch): random valid tours on seeded Euclidean instances, one all-pairs sweep per instance through
the C ABI (gnngls_moves_eval_a2a), timed with CUDA events (10 launches after 3 warm-ups); a sample of the batch is
checked bit-exactly against the CPU oracle (oracle/gls_port.c).  Prints one JSON line per (op, n, batch).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnngls_b200 import _ops, instances            # noqa: E402
from oracle import gls_port                        # noqa: E402  (checker only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sizes', type=int, nargs='+', default=[20, 50, 100, 200, 500, 1000])
    ap.add_argument('--batches', type=int, nargs='+', default=[1, 64, 1024, 16384, 65536])
    ap.add_argument('--max-bytes', type=float, default=24e9, help='cap on batch*n*n*8 bytes of distance matrices')
    ap.add_argument('--check', type=int, default=8, help='instances per cell verified against the oracle')
    args = ap.parse_args()
    rng = np.random.default_rng(5)
    for n in args.sizes:
        for B in args.batches:
            shared = B * n * n * 8 > args.max_bytes                 # too large: one matrix shared by the batch
            nD = 1 if shared else B
            D = np.empty((nD, n, n))
            for b0 in range(0, nD, 1024):
                b1 = min(nD, b0 + 1024)
                D[b0:b1] = instances.distance_matrices(rng.random((b1 - b0, n, 2)))
            tours = np.zeros((B, n + 1), dtype=np.int32)
            perm = rng.random((B, n - 1)).argsort(1).astype(np.int32) + 1
            tours[:, 1:n] = perm
            Dd = torch.from_numpy(D[0] if shared else D).cuda()
            td = torch.from_numpy(tours).cuda()
            for op, name, cand, ref in ((_ops.OP_TWO_OPT, 'two_opt_a2a', (n - 2) * (n - 3) // 2, gls_port.two_opt_a2a),
                                        (_ops.OP_RELOCATE, 'relocate_a2a', (n - 2) ** 2, gls_port.relocate_a2a)):
                for _ in range(3):
                    delta, move, new = _ops.moves_eval(op, Dd, td, None, False)
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(10):
                    delta, move, new = _ops.moves_eval(op, Dd, td, None, False)
                t1.record()
                torch.cuda.synchronize()
                ms = t0.elapsed_time(t1) / 10
                ok = True
                for b in rng.choice(B, size=min(B, args.check), replace=False):
                    d, t, _ = ref(tours[b], D[0 if shared else b])
                    ok &= (np.float64(d).view(np.uint64) == delta[b].cpu().numpy().view(np.uint64)) and new[b].tolist() == t
                print(json.dumps({'op': name, 'n': n, 'batch': B, 'shared_D': bool(shared), 'ms': round(ms, 4),
                                  'candidates_per_s': cand * B / (ms / 1e3), 'bit_exact_vs_oracle': bool(ok)}), flush=True)
            del Dd, td


if __name__ == '__main__':
    main()
