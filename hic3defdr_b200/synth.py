"""
Deterministic synthetic Hi-C inputs in the reference's on-disk input formats
(``<rep>/<chrom>_raw.npz`` scipy.sparse CSR, ``<rep>/<chrom>_kr.bias`` text).

Sampling model (it follows the reference's own
simulation model, hic3defdr/util/simulation.py:70-204): for every band cell
(i, j=i+d), 0 <= d <= D+pad, mean = A/(1+d), dispersion 0.01 + 1e-4*d*(10kb/res),
per-replicate log-normal bias with 1 % of bins failing ``bias_thresh``, depth
factors 0.8 + 0.1*r, negative-binomial counts, zeros dropped.
"""
import os

import numpy as np
import scipy.sparse as sparse

MM10_10KB = {  # ceil(mm10 chromosome length / 10 kb)
    'chr1': 19548, 'chr2': 18212, 'chr3': 16004, 'chr4': 15651, 'chr5': 15184,
    'chr6': 14974, 'chr7': 14545, 'chr8': 12941, 'chr9': 12460, 'chr10': 13070,
    'chr11': 12209, 'chr12': 12013, 'chr13': 12043, 'chr14': 12491,
    'chr15': 10405, 'chr16': 9821, 'chr17': 9499, 'chr18': 9071, 'chr19': 6144,
    'chrX': 17104,
}

_HG38_BP = {  # GRCh38 primary assembly chromosome lengths (bp)
    'chr1': 248956422, 'chr2': 242193529, 'chr3': 198295559,
    'chr4': 190214555, 'chr5': 181538259, 'chr6': 170805979,
    'chr7': 159345973, 'chr8': 145138636, 'chr9': 138394717,
    'chr10': 133797422, 'chr11': 135086622, 'chr12': 133275309,
    'chr13': 114364328, 'chr14': 107043718, 'chr15': 101991189,
    'chr16': 90338345, 'chr17': 83257441, 'chr18': 80373285,
    'chr19': 58617616, 'chr20': 64444167, 'chr21': 46709983,
    'chr22': 50818468, 'chrX': 156040895,
}
HG38_5KB = {c: -(-n // 5000) for c, n in _HG38_BP.items()}
HG38_CHR1_1KB = {'chr1': -(-_HG38_BP['chr1'] // 1000)}

BASE_SEED = 20261018


def band_coords(n, width):
    """(row, col) of every upper-triangular band cell with col-row <= width,
    in (row, col) order."""
    d = np.arange(width + 1)
    row = np.repeat(np.arange(n), width + 1)
    col = row + np.tile(d, n)
    keep = col < n
    return row[keep].astype(np.int64), col[keep].astype(np.int64)


def make_chrom(n, n_reps, dist_max, seed, amp=300.0, res_scale=1.0, pad=5,
               loops=False, bad_frac=0.01, dtype=np.int64,
               return_classes=False):
    """Returns (list of CSR matrices, bias array (n, n_reps), loop clusters
    [, ground-truth class of every cluster: 'constit', 'A' or 'B'])."""
    row, col = band_coords(n, dist_max + pad)
    d = (col - row).astype(float)
    mu = amp / (1.0 + d)
    phi = 0.01 + 1e-4 * d * res_scale
    clusters = None
    effect = None
    classes = None
    if loops:
        lrng = np.random.default_rng(seed + 77)
        effect = np.ones((2, len(row)))
        clusters = []
        cell = {}
        # sparse dict only for loop cells
        n_loops = max(1, n // 25)
        li = lrng.integers(2, max(3, n - dist_max - 3), size=n_loops)
        ld = lrng.integers(10, min(150, dist_max - 2) + 1, size=n_loops)
        cls = lrng.choice(3, size=n_loops, p=[0.6, 0.2, 0.2])
        classes = np.array(['constit', 'A', 'B'], dtype='U7')[cls]
        key = row * (dist_max + pad + 1) + (col - row)
        order = np.argsort(key)
        skey = key[order]
        for i0, d0, c0 in zip(li, ld, cls):
            px = [(int(i0 + a), int(i0 + d0 + b)) for a in (-1, 0, 1)
                  for b in (-1, 0, 1)]
            clusters.append(px)
            if c0 == 0:
                continue
            for (a, b) in px:
                if b - a < 0 or b - a > dist_max + pad or b >= n or a < 0:
                    continue
                k = a * (dist_max + pad + 1) + (b - a)
                pos = np.searchsorted(skey, k)
                if pos < len(skey) and skey[pos] == k:
                    effect[c0 - 1, order[pos]] = 1.5
        del cell
    mats = []
    bias = np.zeros((n, n_reps))
    for r in range(n_reps):
        rng = np.random.default_rng(seed + r)
        b = rng.lognormal(0.0, 0.2, size=n)
        bad = rng.random(n) < bad_frac
        b[bad] = 0.05
        bias[:, r] = b
        depth = 0.8 + 0.1 * r if n_reps <= 4 else 0.7 + 0.08 * r
        m = mu * b[row] * b[col] * depth
        if effect is not None:
            m = m * effect[0 if r < n_reps // 2 else 1]
        shape = 1.0 / phi
        lam = rng.gamma(shape, m / shape)
        x = rng.poisson(lam)
        nz = x > 0
        mats.append(sparse.csr_matrix(
            (x[nz].astype(dtype), (row[nz], col[nz])), shape=(n, n)))
    if return_classes:
        return mats, bias, clusters, classes
    return mats, bias, clusters


def _write_chrom(task):
    """one chromosome of ``write_dataset`` (a top-level function: it also runs
    in forked worker processes); returns whether cluster files were written."""
    (root, rep_names, ci, chrom, n, n_reps, dist_max, config, amp, res_scale,
     loops, dtype) = task
    seed = BASE_SEED + 1000 * config + 100 * ci
    mats, bias, clusters, classes = make_chrom(
        n, n_reps, dist_max, seed, amp=amp, res_scale=res_scale,
        loops=loops, dtype=dtype, return_classes=True)
    for r, rep in enumerate(rep_names):
        sparse.save_npz(os.path.join(root, rep, '%s_raw.npz' % chrom), mats[r])
        np.savetxt(os.path.join(root, rep, '%s_kr.bias' % chrom), bias[:, r])
    if clusters is None:
        return False
    import json
    with open(os.path.join(root, 'clusters', 'loops_%s.json' % chrom),
              'w') as h:
        json.dump([[list(p) for p in c] for c in clusters], h)
    # ground truth in the format of the reference's simulate()
    # (analysis/simulation.py:141: one label per cluster)
    np.savetxt(os.path.join(root, 'clusters', 'labels_%s.txt' % chrom),
               classes, fmt='%s')
    return True


def write_dataset(root, chrom_sizes, n_reps=4, dist_max=200, config=1,
                  amp=300.0, res_scale=1.0, loops=False, dtype=np.int64,
                  n_jobs=1, generate=True):
    """Writes a dataset under ``root`` and returns the kwargs for HiC3DeFDR
    (raw_npz_patterns, bias_patterns, chroms, design).  ``n_jobs`` > 1
    generates the chromosomes in that many forked processes (same files: the
    random streams are per chromosome and replicate); fork before the process
    has initialised CUDA.  ``generate=False`` only describes a dataset that an
    earlier call with the same arguments wrote."""
    import pandas as pd
    rep_names = ['A%d' % (i + 1) for i in range(n_reps // 2)] + \
        ['B%d' % (i + 1) for i in range(n_reps - n_reps // 2)]
    for rep in rep_names:
        os.makedirs(os.path.join(root, rep), exist_ok=True)
    if loops:
        os.makedirs(os.path.join(root, 'clusters'), exist_ok=True)
    tasks = [(root, rep_names, ci, chrom, n, n_reps, dist_max, config, amp,
              res_scale, loops, dtype)
             for ci, (chrom, n) in enumerate(chrom_sizes.items())]
    if not generate:
        wrote = [loops]
    elif n_jobs > 1 and len(tasks) > 1:
        import multiprocessing
        # largest chromosomes first: the pool drains evenly
        order = sorted(range(len(tasks)), key=lambda i: -tasks[i][4])
        with multiprocessing.get_context('fork').Pool(
                min(n_jobs, len(tasks))) as pool:
            wrote = pool.map(_write_chrom, [tasks[i] for i in order],
                             chunksize=1)
    else:
        wrote = [_write_chrom(t) for t in tasks]
    loop_patterns = None
    if any(wrote):
        loop_patterns = {
            'A': os.path.join(root, 'clusters', 'loops_<chrom>.json'),
            'B': os.path.join(root, 'clusters', 'loops_<chrom>.json')}
    design = pd.DataFrame(
        {'A': [r.startswith('A') for r in rep_names],
         'B': [r.startswith('B') for r in rep_names]}, index=rep_names)
    return dict(
        raw_npz_patterns=[os.path.join(root, rep, '<chrom>_raw.npz')
                          for rep in rep_names],
        bias_patterns=[os.path.join(root, rep, '<chrom>_kr.bias')
                       for rep in rep_names],
        chroms=list(chrom_sizes.keys()),
        design=design,
        loop_patterns=loop_patterns,
    )
