"""Shared loaders for the golden fixtures (tests only)."""
import json
import os

import numpy as np
import scipy.sparse as sparse

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_kats():
    with open(os.path.join(GOLDEN, 'reference_kats.json')) as h:
        return json.load(h)


def load_pipeline_golden():
    g = np.load(os.path.join(GOLDEN, 'ref_pipeline.npz'))
    dist_max, n_reps, dist_min = [int(v) for v in g['meta']]
    chroms = sorted({k.split('_')[1] for k in g.files if k.startswith('in_')
                     and k.endswith('_bias')})
    inputs, loops = [], []
    for c in chroms:
        bias = g['in_%s_bias' % c]
        n = bias.shape[0]
        mats = [sparse.csr_matrix((g['in_%s_data_%d' % (c, r)],
                                   g['in_%s_indices_%d' % (c, r)],
                                   g['in_%s_indptr_%d' % (c, r)]),
                                  shape=(n, n)) for r in range(n_reps)]
        inputs.append((mats, bias))
        lp = g['in_%s_loops' % c]
        loops.append([[tuple(p) for p in lp]])
    return dict(g=g, chroms=chroms, inputs=inputs, loops=loops,
                dist_max=dist_max, dist_min=dist_min,
                design=g['design'].astype(bool))


def load_stage_golden():
    return np.load(os.path.join(GOLDEN, 'ref_stages.npz'))
