#!/usr/bin/env python
# coding: utf-8
"""B200 drop-in for the reference's evaluation driver (/root/reference/scripts/test.py): same positional
arguments and options, same pickled pandas.DataFrame of search progress.

    ./test.py <dataset>/test.txt <checkpoint.pt> <run dir> regret_pred --use_gpu
    ./test.py ... --batched --n_iters 10          # extension: whole test set in one batched GPU pass

Default mode follows the reference protocol instance by instance (wall-clock --time_limit that includes feature
construction and the GNN forward, test.py:64,92).  --batched solves all instances at once with a fixed number of
GLS outer iterations; search events then carry the batch completion time.
"""
import argparse
import datetime
import json
import pathlib
import sys
import time
import uuid

import numpy as np
import pandas as pd
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import gnngls_b200 as gnngls                                        # noqa: E402
from gnngls_b200 import algorithms, datasets, models, pipeline     # noqa: E402


def main(argv=None):
    parser = argparse.ArgumentParser(description='Test model')
    parser.add_argument('data_path', type=pathlib.Path)
    parser.add_argument('model_path', type=pathlib.Path)
    parser.add_argument('run_dir', type=pathlib.Path)
    parser.add_argument('guides', type=str, nargs='+')
    parser.add_argument('--time_limit', type=float, default=10.)
    parser.add_argument('--perturbation_moves', type=int, default=20)
    parser.add_argument('--use_gpu', action='store_true')
    parser.add_argument('--batched', action='store_true', help='extension: batched GPU pass with --n_iters outer iterations')
    parser.add_argument('--n_iters', type=int, default=10)
    args = parser.parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit('gnngls_b200 has no CPU path: a CUDA device is required')
    device = torch.device('cuda')

    params = json.load(open(args.model_path.parent / 'params.json'))
    test_set = datasets.TSPDataset(args.data_path, feat_drop_idx=params.get('efeat_drop_idx', []))
    model = None
    if 'regret_pred' in args.guides:
        _, feat_dim = test_set[0].ndata['features'].shape
        model = models.EdgePropertyPredictionModel(feat_dim, params['embed_dim'], 1, params['n_layers'],
                                                   n_heads=params['n_heads']).to(device)
        checkpoint = torch.load(args.model_path, map_location=device)
        model.load_state_dict(checkpoint['model_state_dict'])
        model.eval()

    search_progress = []
    gaps = []
    if args.batched:
        Gs = [datasets.load_instance(test_set.root_dir / name) for name in test_set.instances]
        t0 = time.time()
        D = torch.from_numpy(np.stack([gnngls.edge_matrix(G, 'weight') for G in Gs])).to(device)
        if model is not None:
            solver = pipeline.RegretGLS(model, pipeline.Scalers.from_sklearn(test_set.scalers))
            res = solver.solve(D, n_iters=args.n_iters, perturbation_moves=args.perturbation_moves, guides=tuple(args.guides),
                               max_events=4096)
        else:
            solver = pipeline.RegretGLS(None)
            res = solver.solve(D, n_iters=args.n_iters, perturbation_moves=args.perturbation_moves, guides=tuple(args.guides),
                               max_events=4096)
        torch.cuda.synchronize()
        t1 = time.time()
        events, n_events = res.extra['events'].cpu().numpy(), res.extra['n_events'].cpu().numpy()
        for b, (name, G) in enumerate(zip(test_set.instances, Gs)):
            opt_cost = gnngls.optimal_cost(G, weight='weight')
            search_progress.append({'instance': name, 'time': t0, 'opt_cost': opt_cost})
            for c in events[b, :n_events[b]]:
                search_progress.append({'instance': name, 'time': t1, 'cost': float(c), 'opt_cost': opt_cost})
            gaps.append((float(res.best_costs[b]) / opt_cost - 1) * 100)
    else:
        for instance in test_set.instances:
            G = datasets.load_instance(test_set.root_dir / instance)
            opt_cost = gnngls.optimal_cost(G, weight='weight')
            t = time.time()
            search_progress.append({'instance': instance, 'time': t, 'opt_cost': opt_cost})
            if model is not None:
                H = test_set.get_scaled_features(G).to(device)
                with torch.no_grad():
                    y_pred = model(H, H.ndata['features'])
                regret_pred = test_set.scalers['regret'].inverse_transform(y_pred.cpu().numpy())
                for e, r in zip(H.ndata['e'].cpu().numpy(), regret_pred):
                    G.edges[tuple(e)]['regret_pred'] = np.maximum(r.item(), 0)
                init_tour = algorithms.nearest_neighbor(G, 0, weight='regret_pred')
            else:
                init_tour = algorithms.nearest_neighbor(G, 0, weight='weight')
            init_cost = gnngls.tour_cost(G, init_tour)
            best_tour, best_cost, progress = algorithms.guided_local_search(
                G, init_tour, init_cost, t + args.time_limit, weight='weight', guides=args.guides,
                perturbation_moves=args.perturbation_moves, first_improvement=False)
            for row in progress:
                row.update({'instance': instance, 'opt_cost': opt_cost})
                search_progress.append(row)
            gaps.append((best_cost / opt_cost - 1) * 100)
    print('Avg Gap: {:.4f}'.format(float(np.mean(gaps))))

    df = pd.DataFrame.from_records(search_progress)
    df['best_cost'] = df.groupby('instance')['cost'].cummin()
    df['gap'] = (df['best_cost'] / df['opt_cost'] - 1) * 100
    df['dt'] = df['time'] - df.groupby('instance')['time'].transform('min')
    run_name = f"{datetime.datetime.now().strftime('%b%d_%H-%M-%S')}_{uuid.uuid4().hex}.pkl"
    args.run_dir.mkdir(parents=True, exist_ok=True)
    df.to_pickle(args.run_dir / run_name)
    return args.run_dir / run_name


if __name__ == '__main__':
    main()
