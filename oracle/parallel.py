"""
TEST / BASELINE INFRASTRUCTURE -- the oracle pipeline fanned out over host
cores the way the reference does it: a process pool over chromosomes for
prepare_data and lrt (hic3defdr/util/parallelization.py:17-39, used at
analysis/analysis.py:68-74, 249-254) and over distances for estimate_disp
(parallelization.py:53-100, used at analysis/analysis.py:194-200).  Only
bench.py's ``cpu_baseline`` / ``--impl reference`` legs call this.
"""
import multiprocessing as mp
import os

import numpy as np

from oracle import pipeline as op

_G = {}


def _prepare(i):
    mats, bias = _G['inputs'][i]
    return op.prepare_chrom(mats, bias, _G['design'], **_G['prep_kw'])


def _qcml_bin(task):
    d, c = task
    sel = _G['dist'] == d
    if not sel.any():
        return np.nan
    reps = _G['design'][:, c]
    return op.qcml(_G['raw'][sel][:, reps], f=_G['f'][sel][:, reps])


def _lrt(i):
    st = _G['per'][i]
    di = st['disp_idx']
    return op.lrt(st['raw'][di], st['f'],
                  np.dot(st['disp'], _G['design'].T.astype(float)),
                  _G['design'], True)


def _pool(n):
    return mp.get_context('fork').Pool(n)


def run_to_qvalues(chrom_inputs, design, dist_min=4, dist_max=200,
                   bias_thresh=0.1, mean_thresh=1.0, n_threads=-1):
    """Same result as oracle.pipeline.run_to_qvalues (default kwargs), with
    the reference's process-level parallelism.  Returns the stage dict."""
    import warnings
    warnings.simplefilter('ignore')
    if n_threads == -1:
        n_threads = os.cpu_count() or 1
    design = np.asarray(design).astype(bool)
    _G.clear()
    _G.update(inputs=chrom_inputs, design=design,
              prep_kw=dict(dist_min=dist_min, dist_max=dist_max,
                           bias_thresh=bias_thresh, mean_thresh=mean_thresh))
    if n_threads > 1:
        with _pool(min(n_threads, len(chrom_inputs))) as pool:
            per = pool.map(_prepare, range(len(chrom_inputs)))
    else:
        per = [_prepare(i) for i in range(len(chrom_inputs))]
    raws, fs, dists = [], [], []
    for st in per:
        di = st['disp_idx']
        r, c = st['row'][di], st['col'][di]
        st['f'] = op.combined_factor(st['bias'], r, c, st['size_factors'][di])
        raws.append(st['raw'][di])
        fs.append(st['f'])
        dists.append(c - r)
    raw, f, dist = np.concatenate(raws), np.concatenate(fs), \
        np.concatenate(dists)
    _G.update(raw=raw, f=f, dist=dist)
    n_cond = design.shape[1]
    tasks = [(d, c) for c in range(n_cond) for d in range(dist_max + 1)]
    if n_threads > 1:
        with _pool(n_threads) as pool:
            vals = pool.map(_qcml_bin, tasks, chunksize=1)
    else:
        vals = [_qcml_bin(t) for t in tasks]
    per_dist = np.array(vals).reshape(n_cond, dist_max + 1).T
    disp = np.zeros((len(dist), n_cond))
    for c in range(n_cond):
        ok = np.isfinite(per_dist[:, c])
        xs, ys = np.arange(dist_max + 1)[ok], per_dist[:, c][ok]
        fit = op.weighted_trend(xs, ys, left_boundary=ys[0])
        disp[:, c] = op.eval_trend(fit, dist)
    offs = np.cumsum([0] + [len(d) for d in dists])
    for i, st in enumerate(per):
        st['disp'] = disp[offs[i]:offs[i + 1]]
    _G.update(per=per)
    if n_threads > 1:
        with _pool(min(n_threads, len(per))) as pool:
            res = pool.map(_lrt, range(len(per)))
    else:
        res = [_lrt(i) for i in range(len(per))]
    for st, (p, llr, mu0, mu1) in zip(per, res):
        st.update(pvalues=p, llr=llr, mu_hat_null=mu0, mu_hat_alt=mu1)
    q = op.bh(np.concatenate([st['pvalues'] for st in per]))
    qo = np.cumsum([0] + [len(st['pvalues']) for st in per])
    for i, st in enumerate(per):
        st['qvalues'] = q[qo[i]:qo[i + 1]]
    _G.clear()
    return dict(chroms=per, disp_per_dist=per_dist)
