#!/usr/bin/env python
"""Benchmark of the gnngls hot path on B200:  TSP100 regret_pred + GLS instances/sec.

    python bench.py --gpus N --steps K --warmup W            # our arm (hand-written sm_100a CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port of the reference path

A step = one pass of the hot path (features -> EdgePropertyPredictionModel -> inverse-scale/clamp ->
nearest_neighbor -> tour_cost -> guided_local_search with a fixed number of outer iterations) over
this rank's shard of synthetic instances.  Instances shard across ranks with no collective on the
data path; the final gather of tours/costs over NCCL is inside the timed region.  One JSON line is
printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'TSP100 regret_pred+GLS instances/sec'
UNIT = 'instances/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--n', type=int, default=100, help='cities per instance')
    ap.add_argument('--global-instances', type=int, default=100000,
                    help='instances of the whole job, sharded over the ranks (strong scaling): the 100k-instance TSP100 '
                         'config of BASELINE.json')
    ap.add_argument('--instances-per-gpu', type=int, default=0,
                    help='fixed shard per GPU instead (weak scaling); 0 = --global-instances / world size')
    ap.add_argument('--gls-iters', type=int, default=10, help='GLS outer iterations K (fixed count, SURVEY 8(d))')
    ap.add_argument('--perturbation-moves', type=int, default=20)
    ap.add_argument('--micro-batch', type=int, default=256)
    ap.add_argument('--chunk', type=int, default=2048, help='instances per host->device chunk in the e2e path')
    ap.add_argument('--cpu-sample', type=int, default=16, help='instances in the bounded CPU-baseline sample (all cores)')
    ap.add_argument('--cpu-sample-1core', type=int, default=2, help='instances timed on ONE host core')
    ap.add_argument('--ref-sample', type=int, default=16, help='instances per step of the reference arm')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    a = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', 1))
    a.scaling = 'weak' if a.instances_per_gpu > 0 else 'strong'
    if a.instances_per_gpu <= 0:
        a.instances_per_gpu = max(1, a.global_instances // world)
    return a


# ------------------------------------------------------------------------------------------------
def make_port_model():
    """Random-init weights in the reference checkpoint layout (LFS payloads are absent), seed 0."""
    from oracle import model_port
    torch.manual_seed(0)
    return model_port.EdgeModelPort(1, 128, 1, 3, n_heads=8).eval()


def make_model(device):
    from gnngls_b200 import models
    torch.manual_seed(0)
    m = models.EdgePropertyPredictionModel(1, 128, 1, 3, n_heads=8)      # shipped params.json shape
    ck = os.path.join(ROOT, 'models', 'tsp100', 'checkpoint_best_val.pt')
    src = 'random-init(seed 0)'
    if os.path.exists(ck) and os.path.getsize(ck) > 1 << 20:
        m.load_state_dict(torch.load(ck, map_location='cpu')['model_state_dict'])
        src = ck
    return m.to(device).eval(), src


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ------------------------------------------------------------------------------------------------
def cpu_calibrated_scalers(port, n):
    """CPU twin of RegretGLS.calibrate_synthetic_regret_scaler (one calibration instance)."""
    from gnngls_b200 import instances
    from gnngls_b200.pipeline import Scalers
    from oracle import model_port
    s = Scalers()
    _, D = instances.random_instances(1, n, seed=instances.DEFAULT_SEED - 1)
    x = instances.edge_features(D)
    x = ((x.astype(np.float64) * s.feat_scale).astype(np.float32).astype(np.float64) + s.feat_min).astype(np.float32)
    with torch.no_grad():
        y = port(model_port.EdgeListGraph.kn_line_graph(n, 1), torch.from_numpy(x[0]).reshape(-1, 1))
    return s.calibrated(y)


def cpu_reference_pass(port, D, n_iters, pm, threads, s, want_regret=False):
    """The reference's test.py:72-95 flow on CPU via the oracle: torch-CPU model (all threads) then the
    C port of nearest_neighbor + guided_local_search across `threads` host threads."""
    from gnngls_b200 import instances
    from oracle import gls_port, model_port
    B, n = D.shape[0], D.shape[-1]
    N = n * (n - 1) // 2
    x = instances.edge_features(D)
    x = ((x.astype(np.float64) * s.feat_scale).astype(np.float32).astype(np.float64) + s.feat_min).astype(np.float32)
    g = model_port.EdgeListGraph.kn_line_graph(n, 1)
    regret = np.empty((B, N), dtype=np.float32)
    with torch.no_grad():
        for b in range(B):                                      # one instance at a time, like test.py:59
            y = port(g, torch.from_numpy(x[b]).reshape(-1, 1)).numpy().reshape(-1)
            r = ((y.astype(np.float64) - s.regret_min).astype(np.float32).astype(np.float64) / s.regret_scale).astype(np.float32)
            regret[b] = np.maximum(r, 0)
    out = gls_port.pipeline_batch(D, regret, n_iters, pm, nthreads=threads)
    return out + (regret,) if want_regret else out


def run_reference(args, rank, world):
    if rank != 0:
        return
    from gnngls_b200 import instances
    from oracle import gls_port
    gls_port.build()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    port = make_port_model()
    scalers = cpu_calibrated_scalers(port, args.n)
    _, D = instances.random_instances(args.ref_sample, args.n)
    for _ in range(args.warmup):
        cpu_reference_pass(port, D, args.gls_iters, args.perturbation_moves, threads, scalers)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_pass(port, D, args.gls_iters, args.perturbation_moves, threads, scalers)
    dt = time.perf_counter() - t0
    value = args.ref_sample * args.steps / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': args.scaling,
        'vs_baseline': None, 'dtype': 'fp32 (model) / fp64 (search)', 'data': 'synthetic',
        'config': workload_config(args, world),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': f'{args.ref_sample} TSP{args.n} instances per step: oracle torch-CPU model (restated '
                                   f'GATConv, {threads} threads) + C port of nearest_neighbor/GLS, K={args.gls_iters}'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        'workload': f'TSP{args.n}: EdgePropertyPredictionModel(1,128,1,3,n_heads=8) regret_pred + nearest_neighbor + '
                    f'guided_local_search(guides=[regret_pred], perturbation_moves={args.perturbation_moves}, '
                    f'K={args.gls_iters} fixed outer iterations)',
        'n': args.n, 'instances_per_gpu': args.instances_per_gpu, 'global_instances': args.instances_per_gpu * world,
        'gls_outer_iters': args.gls_iters, 'perturbation_moves': args.perturbation_moves,
        'micro_batch': args.micro_batch, 'parallelism': f'instance-sharded x{world}, final NCCL gather only',
        'l2': 'per-step inputs+activations exceed the 126 MB L2 (no explicit flush)',
    }


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    import torch.distributed as dist
    from gnngls_b200 import _lib, _timing, instances, pipeline
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; gnngls_b200 has no CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    model, weights = make_model(dev)
    solver = pipeline.RegretGLS(model, micro_batch=args.micro_batch)
    S, n = args.instances_per_gpu, args.n
    D_host = torch.empty(S, n, n, dtype=torch.float64).pin_memory()
    rng = np.random.default_rng(instances.DEFAULT_SEED + rank)
    for b0 in range(0, S, 2048):                     # chunked: the temporaries of distance_matrices are 3x the chunk
        b1 = min(S, b0 + 2048)
        D_host[b0:b1] = torch.from_numpy(instances.distance_matrices(rng.random((b1 - b0, n, 2))))
    D_np = D_host.numpy()
    D_dev = D_host.to(dev)
    if weights.startswith('random-init'):
        # untrained weights give near-constant raw outputs: calibrate the synthetic regret scaler once so the
        # guide is non-degenerate (a few % exact zeros, rest spread), identically on every rank
        _, D_cal = instances.random_instances(min(S, args.micro_batch), n, seed=instances.DEFAULT_SEED - 1)
        solver.calibrate_synthetic_regret_scaler(torch.from_numpy(D_cal).to(dev))
    kw = dict(n_iters=args.gls_iters, perturbation_moves=args.perturbation_moves)
    from gnngls_b200 import distributed as gd

    def step_resident():
        res = solver.solve(D_dev, **kw)
        # the only collective: final gather of tours/costs over NCCL/NVLink (no-op for one rank)
        gd.gather_results(res.best_tours, res.best_costs, world * S)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        res = step_resident()
    barrier()
    # ---------------- timed region: inputs resident in HBM
    launches0 = lib.launches
    timer = _timing.StageTimer()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks, timer:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            res = step_resident()
        ev1.record()
        barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    launches = lib.launches - launches0
    value = world * S * args.steps / (ms / 1e3)
    stages = timer.totals_ms()
    counters = res.counters.sum(0).tolist()        # last step: sweeps / o2a scans / moves on this rank

    # ---------------- e2e: public API on HOST buffers, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        solver.solve_host(D_host[: min(S, args.chunk)], chunk=args.chunk, **kw)          # warm pinned staging
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            tours_h, costs_h = solver.solve_host(D_host, chunk=args.chunk, **kw)
        t1.record()
        barrier()
        ms2 = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e = {'value': world * S * args.steps / (float(ms2) / 1e3), 'unit': UNIT,
               'h2d_bytes_per_step': int(D_host.numel() * 8), 'd2h_bytes_per_step': int(S * (n + 1) * 4 + S * 8),
               'api': 'gnngls_b200.pipeline.RegretGLS.solve_host (pinned host D -> tours/costs on host)'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel: the K_n GAT aggregate (one launch per layer and micro-batch)
    # achieved = ALGORITHMIC bytes / measured kernel time, with the algorithmic bytes of DESIGN.md section 5: every node of the
    # line graph once -- fp16 ft 256 B + el/er 64 B + skip row h 512 B read, h1 512 B written = 1,344 B per node and layer.
    # (SURVEY 8(d)'s 544 B per EDGE counts the 2(n-2)-fold logical gather that the star formulation serves from shared /
    # tensor memory; it is reported as `logical_gather_gbs`, never as a fraction of the HBM peak.)
    N_nodes, E = n * (n - 1) // 2, n * (n - 1) * (n - 2)
    gat_ms, gat_calls = stages.get('gat_kn', stages.get('gat_csr', (0.0, 0)))
    per_call_instances = min(args.micro_batch, S)
    node_bytes = 256 + 2 * 32 + 512 + 512
    peak, peak_src = peaks()
    roof = None
    if gat_calls:
        inst_layers = S * 8 * args.steps
        achieved = N_nodes * node_bytes * inst_layers / (gat_ms / 1e3) / 1e9
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, 'profiles', 'gat_kn_traffic.json')          # written by tools/ncu_traffic.py from an ncu --set full capture
        if os.path.exists(tp):
            t = json.load(open(tp))
            if str(t.get('kernel', '')).startswith('gat_kn_tc_kernel') and t.get('n') == n:
                traffic = t['dram_bytes_per_instance_layer'] * per_call_instances
                traffic_src = t.get('source')
        tc_kernel = n <= 128        # csrc/gat_kn.cu: larger stars run the exact sorted-prefix kernel
        roof = {'kernel': 'gat_kn_tc_kernel (K_n edge-softmax/aggregate + skip + BN1 on tcgen05)' if tc_kernel else
                          'gat_kn_scan_kernel (K_n edge-softmax/aggregate + skip + BN1, exact sorted-prefix sums, n > 128)',
                'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': N_nodes * node_bytes * per_call_instances,
                'avg_launch_ms': gat_ms / gat_calls, 'launches_timed': gat_calls,
                'logical_gather_gbs': (E * 544 + N_nodes * 544) * inst_layers / (gat_ms / 1e3) / 1e9,
                'limiter': ('latency / instruction issue, not HBM (profiles/r2_kn_tc.md)' if tc_kernel else
                            'sort + prefix-sum latency per star, not HBM (profiles/r2_kernels_n500.md)') +
                           ': the HBM fraction says how far the kernel is from the only roofline that bounds its compulsory traffic'}
    # the other model kernels against the same HBM peak (compulsory bytes per node and layer / measured stage time)
    other = {}
    for name, nbytes, nlaunch in (('fc', 256 + 256 + 64, 8), ('ff', 512 + 512 + 256, 8)):
        ms_k, calls = stages.get(name, (0.0, 0))
        if calls:
            gbs = N_nodes * nbytes * S * nlaunch * args.steps / (ms_k / 1e3) / 1e9
            other[name] = {'bound': 'hbm', 'achieved': gbs, 'frac': gbs / peak, 'unit': 'GB/s', 'bytes_per_node': nbytes}
    if roof is not None:
        roof['other_kernels'] = other
    total_stage = sum(v[0] for v in stages.values())
    stage_ms = {k: round(v[0] / args.steps, 3) for k, v in stages.items()}
    cand_2opt, cand_rel = (n - 2) * (n - 3) // 2, (n - 2) ** 2
    gls_ms = stages.get('gls', (0.0, 0))[0] / max(args.steps, 1)
    moves = {'a2a_candidates_per_s': (counters[0] * cand_2opt + counters[1] * cand_rel) / (gls_ms / 1e3) if gls_ms else None,
             'two_opt_sweeps': counters[0], 'relocate_sweeps': counters[1], 'o2a_scans': counters[2],
             'accepted_moves': counters[3], 'note': 'per step on rank 0; rate = a2a candidates / GLS kernel time'}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import gls_port
        gls_port.build()
        threads = os.cpu_count() or 1
        port = make_port_model()
        sample = min(args.cpu_sample, S)
        torch.set_num_threads(threads)
        t0 = time.perf_counter()
        o_t, o_c = cpu_reference_pass(port, D_np[:sample], args.gls_iters, args.perturbation_moves, threads, solver.scalers)
        dt = time.perf_counter() - t0
        s1 = min(args.cpu_sample_1core, sample)
        torch.set_num_threads(1)
        t0 = time.perf_counter()
        cpu_reference_pass(port, D_np[:s1], args.gls_iters, args.perturbation_moves, 1, solver.scalers)
        dt1 = time.perf_counter() - t0
        torch.set_num_threads(threads)
        g_c = res.best_costs[:sample].cpu().numpy()
        d = g_c - o_c                                   # paired: same instances, each side with its own predicted regrets
        half = 1.96 * float(d.std(ddof=1)) / np.sqrt(sample) if sample > 1 else float('nan')
        cpu = {'value': sample / dt, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': f'first {sample} instances of the same workload, {dt:.1f} s: oracle torch-CPU model ({threads} threads) + C port '
                         f'of nearest_neighbor/GLS (K={args.gls_iters})',
               'single_core': {'value': s1 / dt1, 'unit': UNIT, 'cores': 1, 'sample': f'first {s1} instances, {dt1:.1f} s'},
               'best_cost_mean_cpu': float(o_c.mean()), 'best_cost_mean_gpu_same_instances': float(g_c.mean()),
               'best_cost_paired_diff': {'mean_gpu_minus_cpu': float(d.mean()), 'ci95_halfwidth': half,
                                         'relative_to_cpu_mean': float(d.mean() / o_c.mean()), 'pairs': sample,
                                         'note': 'random-init weights: the guide is noise, near-ties flip between the fp32 CPU '
                                                 'model and the tensor-core model; identical regrets give identical tours '
                                                 '(tests/test_model_gpu.py)'}}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'fp16 operands (10-bit mantissa, as TF32) with fp32 accumulate for the GNN contractions, fp32 elsewhere + fp64 (search)', 'data': 'synthetic',
        'config': dict(workload_config(args, world), weights=weights),
        'clocks': clocks.summary(), 'e2e': e2e, 'gpu_launches': launches, 'roofline': roof, 'cpu_baseline': cpu,
        'stage_ms_per_step': stage_ms, 'stage_coverage': round(total_stage / ms, 3) if ms else None,
        'search': moves, 'mean_best_cost': float(res.best_costs.mean()), 'mean_init_cost': float(res.init_costs.mean()),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
