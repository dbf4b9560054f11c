"""Inference-side drop-in for /root/reference/gnngls/datasets.py.

``TSPDataset`` keeps the reference's constructor and ``get_scaled_features(G)`` contract — one template line graph
built from the first instance, per-instance MinMax-scaled node data in ``ndata`` — but the template is a
``gnngls_b200.LineGraph`` (adjacency computed arithmetically on the GPU, datasets.py:56-60 builds it with
networkx + DGL) and the O(N) Python loop of datasets.py:77-83 is a vectorised gather.  Instances are the
reference's pickled ``networkx.Graph`` files; ``nx.read_gpickle`` no longer exists in networkx >= 3, so they are
read with ``pickle.load``.  ``set_labels`` (LKH) is out of scope; ``set_features`` is kept.
"""
import copy
import pathlib
import pickle

import numpy as np
import torch
import torch.utils.data

from .graph import LineGraph, kn_edges


def load_instance(path):
    """A pickled networkx.Graph as written by scripts/generate_instances.py:55 (nx.write_gpickle == pickle.dump)."""
    with open(path, 'rb') as f:
        return pickle.load(f)


def set_features(G):
    """datasets.py:14-20: the single edge feature is the float32 edge weight."""
    for e in G.edges:
        G.edges[e]['features'] = np.array([G.edges[e]['weight']], dtype=np.float32)


def edge_attribute_vector(G, attr, n=None, default=None):
    """Values of an edge attribute in line-graph node order (sorted tuples i<j) as an [N, d] array."""
    n = G.number_of_nodes() if n is None else n
    es = kn_edges(n)
    rows = []
    for i, j in es:
        d = G.edges[int(i), int(j)]
        if attr in d:
            rows.append(np.atleast_1d(d[attr]))
        elif default is not None:
            rows.append(np.atleast_1d(default(d)))
        else:
            raise KeyError(f"edge ({i},{j}) has no attribute '{attr}'")
    return np.vstack(rows)


class TSPDataset(torch.utils.data.Dataset):
    def __init__(self, instances_file, scalers_file=None, feat_drop_idx=[]):
        if not isinstance(instances_file, pathlib.Path):
            instances_file = pathlib.Path(instances_file)
        self.root_dir = instances_file.parent
        self.instances = [line.strip() for line in open(instances_file) if line.strip()]
        if scalers_file is None:
            scalers_file = self.root_dir / 'scalers.pkl'
        with open(scalers_file, 'rb') as f:
            scalers = pickle.load(f)
        self.scalers = scalers['edges'] if 'edges' in scalers else scalers      # backward compatibility, datasets.py:48-51
        self.feat_drop_idx = list(feat_drop_idx)
        # only works for homogeneous datasets (datasets.py:55): one template graph for all instances
        G = load_instance(self.root_dir / self.instances[0])
        self.n = G.number_of_nodes()
        self.G = LineGraph.complete(self.n)

    def __len__(self):
        return len(self.instances)

    def __getitem__(self, i):
        if torch.is_tensor(i):
            i = i.tolist()
        return self.get_scaled_features(load_instance(self.root_dir / self.instances[i]))

    def get_scaled_features(self, G):
        """datasets.py:73-95.  Returns a copy of the template LineGraph with ndata 'features' (scaled, fp32),
        'regret' (scaled) and 'in_solution' (the reference stores the raw regret there, datasets.py:94)."""
        if G.number_of_nodes() != self.n:
            raise ValueError('TSPDataset only supports homogeneous datasets (all instances the same size)')
        features = edge_attribute_vector(G, 'features', self.n,
                                         default=lambda d: np.array([d['weight']], dtype=np.float32)).astype(np.float32)
        H = copy.deepcopy(self.G)
        ft = self.scalers['features'].transform(features)
        ft = np.delete(ft, self.feat_drop_idx, axis=1)
        H.ndata['features'] = torch.tensor(ft, dtype=torch.float32)
        try:
            regret = edge_attribute_vector(G, 'regret', self.n).astype(np.float64)
            H.ndata['regret'] = torch.tensor(self.scalers['regret'].transform(regret), dtype=torch.float32)
            H.ndata['in_solution'] = torch.tensor(regret, dtype=torch.float32)
        except KeyError:
            pass        # unlabeled instances are fine for inference
        return H
