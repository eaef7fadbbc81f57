"""debug: equalize / NLL cost per pixel as a function of the distance (feeds the
weights of dist.lpt_layout).  chr1..chr4 of the mouse workload on one GPU."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hic3defdr_b200 import engine, ops, staging
cfg = bench.WORKLOADS['mouse10kb']
design = np.array([[1, 0], [1, 0], [0, 1], [0, 1]], dtype=bool)
ins = []
for c in ('chr1', 'chr2', 'chr3', 'chr4'):
    n = cfg['chroms'][c]
    mats, bias = bench.gen_chrom_device(n, 4, 200, 20261018 + 1000 + 100 * list(cfg['chroms']).index(c), 300.0)
    ins.append((staging.csr_to_device(mats, n), bias))
states = engine.prepare_many(ins, design, dist_max=200)
x, f, dist_cat, seg, offs = engine.pool_by_distance(states, 200, n_reps=4)
edges = [4, 8, 12, 20, 30, 50, 75, 100, 125, 150, 175, 201]
for a, b in zip(edges[:-1], edges[1:]):
    sub = np.concatenate([[0], np.cumsum(np.diff(seg)[a:b])]).astype(np.int64)
    lo, hi = int(seg[a]), int(seg[b])
    xs, fs = x[:, lo:hi].contiguous(), f[:, lo:hi].contiguous()
    for rep in range(2):
        dpd, st = ops.estimate_dispersion(xs, fs, sub, design)
    npx = hi - lo
    print('d %3d-%3d px %8d  eq %.3f ns/px-eq  nll %.3f ns/px-eval  outer %.2f  evals/outer %.1f  total %.2f ns/px' % (
        a, b, npx, 1e3 * st['equalize_us'] / st['pixel_equalizations'],
        1e3 * st['nll_us'] / (st['nll_evaluations'] / (2 * (b - a)) * npx * 2),
        st['outer_iterations'] / (2 * (b - a)), st['nll_evaluations'] / st['outer_iterations'],
        1e3 * (st['equalize_us'] + st['nll_us']) / npx))
