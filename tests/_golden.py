"""Helpers to iterate the golden fixtures written by oracle/make_golden.py."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class Case(dict):
    __getattr__ = dict.__getitem__


def load(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    n = int(z['n_cases'])
    cases = []
    for k in range(n):
        pre = f'c{k}_'
        cases.append(Case({key[len(pre):]: z[key] for key in z.files if key.startswith(pre)}))
    top = {key: z[key] for key in z.files if not key.startswith('c') or not key[1:2].isdigit()}
    return cases, top


def bits(x):
    return np.asarray(x, dtype=np.float64).view(np.uint64)
