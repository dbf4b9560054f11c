"""Builds a tiny dataset directory in the reference's on-disk format (pickled networkx graphs with weight / features /
regret / in_solution edge attributes, test.txt, scalers.pkl, params.json + checkpoint) for the datasets / CLI tests."""
import itertools
import json
import pickle

import networkx as nx
import numpy as np
import torch
from sklearn.preprocessing import MinMaxScaler


def make_dataset(root, n=8, count=5, seed=0):
    rng = np.random.default_rng(seed)
    data = root / 'data'
    data.mkdir()
    names = []
    scalers = {'features': MinMaxScaler(), 'regret': MinMaxScaler()}
    for k in range(count):
        P = rng.random((n, 2))
        G = nx.Graph()
        for v in range(n):
            G.add_node(v, pos=P[v])
        for i, j in itertools.combinations(range(n), 2):
            w = float(np.sqrt(((P[i] - P[j]) ** 2).sum()))
            G.add_edge(i, j, weight=w, features=np.array([w], dtype=np.float32), regret=float(rng.random() * 0.3), in_solution=False)
        order = np.argsort(np.arctan2(P[:, 1] - 0.5, P[:, 0] - 0.5))        # a reasonable "solution" tour: angular order
        tour = order.tolist()
        for a, b in zip(tour, tour[1:] + tour[:1]):
            G.edges[a, b]['in_solution'] = True
            G.edges[a, b]['regret'] = 0.0
        name = f'inst{k}.pkl'
        with open(data / name, 'wb') as f:
            pickle.dump(G, f)
        names.append(name)
        for key in scalers:
            scalers[key].partial_fit(np.vstack([G.edges[e][key] for e in G.edges]))
    (data / 'test.txt').write_text('\n'.join(names) + '\n')
    with open(data / 'scalers.pkl', 'wb') as f:
        pickle.dump(scalers, f)
    return data, names, scalers


def make_checkpoint(root, seed=0):
    """Reference-layout checkpoint + params.json (scripts/train.py:60-67,165-168) with seeded random weights."""
    from oracle import model_port
    mdir = root / 'model'
    mdir.mkdir()
    torch.manual_seed(seed)
    port = model_port.EdgeModelPort(1, 128, 1, 3, n_heads=8)
    torch.save({'epoch': 0, 'model_state_dict': port.state_dict(), 'optimizer_state_dict': {}, 'loss': 0.0, 'val_loss': 0.0},
               mdir / 'checkpoint_best_val.pt')
    (mdir / 'params.json').write_text(json.dumps({'embed_dim': 128, 'n_layers': 3, 'n_heads': 8, 'lr_init': 1e-3}))
    return mdir / 'checkpoint_best_val.pt', port
