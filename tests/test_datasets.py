"""datasets.TSPDataset drop-in (host logic, CPU) and the scripts/test.py CLI (GPU)."""
import pickle

import numpy as np
import pandas as pd
import pytest
import torch

from gnngls_b200 import datasets, graph
from tests import _synthetic_dataset as synth


def test_tspdataset_matches_reference_contract(tmp_path):
    data, names, scalers = synth.make_dataset(tmp_path, n=7, count=3)
    ds = datasets.TSPDataset(data / 'test.txt')
    assert len(ds) == 3 and ds.G.number_of_nodes() == 21
    assert ds.G.ndata['e'].tolist() == graph.kn_edges(7).tolist()
    G = datasets.load_instance(data / names[1])
    H = ds.get_scaled_features(G)
    # the reference loop (datasets.py:77-89): node i of the line graph <-> edge ndata['e'][i]
    feats = np.vstack([G.edges[tuple(e)]['features'] for e in ds.G.ndata['e'].numpy()])
    regret = np.vstack([G.edges[tuple(e)]['regret'] for e in ds.G.ndata['e'].numpy()])
    assert H.ndata['features'].dtype == torch.float32
    assert np.array_equal(H.ndata['features'].numpy(), scalers['features'].transform(feats).astype(np.float32))
    assert np.array_equal(H.ndata['regret'].numpy(), scalers['regret'].transform(regret).astype(np.float32))
    assert np.array_equal(H.ndata['in_solution'].numpy(), regret.astype(np.float32))      # datasets.py:94 stores the raw regret
    assert H is not ds.G and 'features' not in ds.G.ndata                                  # template untouched (deepcopy)
    assert torch.equal(ds[1].ndata['features'], H.ndata['features'])
    with open(data / 'scalers.pkl', 'wb') as f:                                            # legacy {'edges': ...} layout
        pickle.dump({'edges': scalers}, f)
    assert datasets.TSPDataset(data / 'test.txt').scalers is not None


def test_features_fall_back_to_weight(tmp_path):
    data, names, scalers = synth.make_dataset(tmp_path, n=5, count=1)
    G = datasets.load_instance(data / names[0])
    for e in G.edges:
        del G.edges[e]['features'], G.edges[e]['regret']
    H = datasets.TSPDataset(data / 'test.txt').get_scaled_features(G)
    assert H.ndata['features'].shape == (10, 1) and 'regret' not in H.ndata
    datasets.set_features(G)
    assert G.edges[0, 1]['features'].dtype == np.float32


@pytest.mark.gpu
@pytest.mark.parametrize('batched', [False, True])
def test_cli_writes_reference_style_run(tmp_path, batched):
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location('b200_test_cli', os.path.join(os.path.dirname(__file__), '..', 'scripts', 'test.py'))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    data, names, _ = synth.make_dataset(tmp_path, n=8, count=4)
    ckpt, _ = synth.make_checkpoint(tmp_path)
    argv = [str(data / 'test.txt'), str(ckpt), str(tmp_path / 'runs'), 'regret_pred', '--use_gpu', '--perturbation_moves', '5']
    argv += ['--batched', '--n_iters', '3'] if batched else ['--time_limit', '0.05']
    out = cli.main(argv)
    df = pd.read_pickle(out)
    assert {'instance', 'time', 'opt_cost', 'cost', 'best_cost', 'gap', 'dt'} <= set(df.columns)
    assert set(df['instance']) == set(names)
    final = df.dropna(subset=['best_cost']).groupby('instance')['best_cost'].last()
    for name in names:
        G = datasets.load_instance(data / name)
        lengths = sum(G.edges[e]['weight'] for e in G.edges if G.edges[e]['in_solution'])
        assert final[name] <= 2.5 * lengths and final[name] > 0
    if batched:
        # the guide must really steer the search: recompute the regrets the CLI used (same checkpoint, same kernels: bitwise
        # reproducible), run the CPU oracle's nearest_neighbor + guided_local_search on them (test.py:79-95) and require the CLI's
        # best costs to be the oracle's, bit for bit
        import json
        import gnngls_b200 as gnngls
        from gnngls_b200 import models, pipeline
        from oracle import gls_port
        params = json.load(open(ckpt.parent / 'params.json'))
        test_set = datasets.TSPDataset(data / 'test.txt')
        model = models.EdgePropertyPredictionModel(1, params['embed_dim'], 1, params['n_layers'], n_heads=params['n_heads']).cuda()
        model.load_state_dict(torch.load(ckpt, map_location='cuda')['model_state_dict'])
        model.eval()
        D = np.stack([gnngls.edge_matrix(datasets.load_instance(data / name), 'weight') for name in test_set.instances])
        solver = pipeline.RegretGLS(model, pipeline.Scalers.from_sklearn(test_set.scalers))
        regret = solver.predict_regret(torch.from_numpy(D).cuda()).cpu().numpy()
        _, o_costs = gls_port.pipeline_batch(D, regret, 3, 5)[:2]
        for b, name in enumerate(test_set.instances):
            assert np.float64(final[name]).tobytes() == np.float64(o_costs[b]).tobytes(), (name, final[name], o_costs[b])
