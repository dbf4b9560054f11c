"""Batched end-to-end hot path: the per-instance loop of /root/reference/scripts/test.py:59-104
(features -> EdgePropertyPredictionModel -> inverse-scale/clamp -> nearest_neighbor -> tour_cost ->
guided_local_search) for a whole batch of instances resident on one GPU.
"""
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _ops
from ._timing import stage
from .graph import LineGraph


@dataclass
class Scalers:
    """MinMaxScaler parameters (scripts/preprocess_dataset.py:39-48): transform is x*scale_+min_."""
    feat_scale: float = 1.0 / float(np.sqrt(2.0))      # synthetic default: data_min=0, data_max=sqrt(2)
    feat_min: float = 0.0
    regret_scale: float = 1.0                           # synthetic default: data_min=0, data_max=1
    regret_min: float = 0.0

    def calibrated(self, y, clamp_frac=0.05, span=0.3):
        """Synthetic regret scaler for UNTRAINED (random-init) weights, whose raw outputs are an arbitrary
        near-constant: map the [clamp_frac, 1-clamp_frac] quantile range of `y` onto [0, span] so that, as with a
        trained model, a few percent of the predicted regrets clamp to exactly 0 and the rest are spread out."""
        q = torch.quantile(y.detach().float().flatten()[:1 << 22], torch.tensor([clamp_frac, 1 - clamp_frac], device=y.device))
        lo, hi = float(q[0]), float(q[1])
        return Scalers(self.feat_scale, self.feat_min, max(hi - lo, 1e-6) / span, lo)

    @classmethod
    def from_sklearn(cls, scalers):
        f, r = scalers['features'], scalers['regret']
        return cls(float(f.scale_[0]), float(f.min_[0]), float(r.scale_[0]), float(r.min_[0]))


@dataclass
class SolveResult:
    best_tours: torch.Tensor          # [B,n+1] int32
    best_costs: torch.Tensor          # [B] fp64
    init_costs: torch.Tensor          # [B] fp64 (nearest-neighbour tours)
    regret: torch.Tensor = None       # [B,N] fp32 (when keep_regret)
    counters: torch.Tensor = None     # [B,4] int64: 2-opt sweeps, relocate sweeps, o2a scans, accepted moves
    status: torch.Tensor = None
    extra: dict = field(default_factory=dict)


class RegretGLS:
    """predict_regret() + solve() for batches of TSP instances given as distance matrices."""

    def __init__(self, model, scalers=None, micro_batch=32):
        self.model = model
        self.scalers = scalers or Scalers()
        self.micro_batch = int(micro_batch)
        self._graphs = {}

    def _graph(self, n, b, device):
        key = (n, b, str(device))
        if key not in self._graphs:
            self._graphs[key] = LineGraph.complete(n, b, device)
        return self._graphs[key]

    @torch.no_grad()
    def calibrate_synthetic_regret_scaler(self, D):
        """See Scalers.calibrated(); uses the raw model outputs of (at most) one micro-batch of D."""
        n = D.shape[-1]
        b = min(D.shape[0], self.micro_batch)
        x = _ops.edge_features(D[:b].contiguous(), self.scalers.feat_scale, self.scalers.feat_min)
        y = self.model(self._graph(n, b, D.device), x.reshape(-1, 1))
        self.scalers = self.scalers.calibrated(y)
        return self.scalers

    @torch.no_grad()
    def predict_regret(self, D):
        """D: [B,n,n] fp64 CUDA -> regret_pred [B,N] fp32 (inverse-scaled, clamped at 0; test.py:72-83)."""
        B, n = D.shape[0], D.shape[-1]
        N = n * (n - 1) // 2
        s = self.scalers
        with stage('features'):
            x = _ops.edge_features(D, s.feat_scale, s.feat_min)
        regret = torch.empty(B, N, dtype=torch.float32, device=D.device)
        for b0 in range(0, B, self.micro_batch):
            b1 = min(B, b0 + self.micro_batch)
            G = self._graph(n, b1 - b0, D.device)
            y = self.model(G, x[b0:b1].reshape(-1, 1))
            with stage('post'):
                _ops.regret_postprocess(y.view(b1 - b0, N), s.regret_scale, s.regret_min, out=regret[b0:b1])
        return regret

    @torch.no_grad()
    def solve(self, D, n_iters=10, perturbation_moves=20, guides=('regret_pred',), keep_regret=False, max_events=0):
        """Full path for D [B,n,n] fp64 on the GPU.  guides: tuple of 'regret_pred' / 'weight'."""
        B, n = D.shape[0], D.shape[-1]
        regret = self.predict_regret(D) if 'regret_pred' in guides else None
        if tuple(guides) == ('regret_pred',):
            gl, kind = regret.view(B, 1, -1), _ops.GUIDE_EDGEVEC_F32
            with stage('nn_init'):
                init_tours, init_costs = _ops.nn_init(regret, kind, D, 0)
        else:
            mats = []
            for g in guides:
                if g == 'weight':
                    mats.append(D)
                else:
                    mats.append(regret_matrix(regret, n))
            gl, kind = torch.stack(mats, 1).contiguous(), _ops.GUIDE_MATRIX_F64
            # test.py:85-88: the initial tour follows regret_pred whenever it is among the guides, else the weight
            nn_guide = mats[list(guides).index('regret_pred')] if 'regret_pred' in guides else D
            init_tours, init_costs = _ops.nn_init(nn_guide.contiguous(), kind, D, 0)
        # a penalties buffer lets the GLS kernel pick its L2-resident tier (faster for n >= 64, see csrc/search.cu)
        state = _ops.GlsState(D, gl.contiguous(), kind, init_tours, init_costs, keep_penalties=n >= 64)
        with stage('gls'):
            info = _ops.gls_run(state, n_iters, perturbation_moves, False, max_events, want_counters=True)
        return SolveResult(state.best_tours, state.best_costs, init_costs, regret if keep_regret else None,
                           info['counters'], info['status'], {'events': info['events'], 'n_events': info['n_events']})

    @torch.no_grad()
    def solve_host(self, D_host, chunk=2048, **kw):
        """Public end-to-end entry point on HOST buffers: D_host [B,n,n] fp64 numpy (or CPU tensor).
        Copies inputs host->device chunk by chunk through pinned memory, runs solve(), and returns
        (best_tours int32 [B,n+1], best_costs fp64 [B]) as numpy arrays."""
        D_host = torch.as_tensor(D_host)
        B, n = D_host.shape[0], D_host.shape[-1]
        tours = torch.empty(B, n + 1, dtype=torch.int32).pin_memory()
        costs = torch.empty(B, dtype=torch.float64).pin_memory()
        dev = next(self.model.parameters()).device
        main = torch.cuda.current_stream(dev)
        if getattr(self, '_copy_stream', None) is None or self._copy_stream.device != dev:
            self._copy_stream = torch.cuda.Stream(dev)
        copy_stream = self._copy_stream
        starts = list(range(0, B, chunk))

        def upload(b0):
            """Host->device copy of one chunk on the copy stream (double-buffered: the next chunk's distance
            matrices travel over PCIe while the current chunk is being solved)."""
            src = D_host[b0:min(B, b0 + chunk)]
            if not src.is_pinned():
                src = src.contiguous().pin_memory()
            with torch.cuda.stream(copy_stream):
                Dd = src.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return Dd, ev

        nxt = upload(starts[0]) if starts else None
        for idx, b0 in enumerate(starts):
            b1 = min(B, b0 + chunk)
            Dd, ev = nxt
            if idx + 1 < len(starts):
                nxt = upload(starts[idx + 1])
            main.wait_event(ev)
            Dd.record_stream(main)                # allocated on the copy stream, consumed on the compute stream
            res = self.solve(Dd, **kw)
            tours[b0:b1].copy_(res.best_tours, non_blocking=True)
            costs[b0:b1].copy_(res.best_costs, non_blocking=True)
        torch.cuda.synchronize(dev)
        return tours.numpy(), costs.numpy()


def regret_matrix(regret, n):
    """[B,N] fp32 -> symmetric [B,n,n] fp64 (what test.py:81-83 stores on the graph edges)."""
    B = regret.shape[0]
    iu = torch.triu_indices(n, n, 1, device=regret.device)
    W = torch.zeros(B, n, n, dtype=torch.float64, device=regret.device)
    W[:, iu[0], iu[1]] = regret.double()
    return W + W.transpose(1, 2)
