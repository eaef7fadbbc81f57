"""GPU parity of the step after bh(): connected components
(h3d_connected_components) and the threshold() / classify() / collect()
methods of the drop-in class against clusters recorded from the unmodified
reference (tests/golden/ref_clusters.json, made by
tests/golden/make_golden_clusters.py).  Clusters are compared as sets of
pixel sets: the reference's order is dict / set iteration order."""
import json
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _fixture():
    with open(os.path.join(GOLDEN, 'ref_clusters.json')) as h:
        return json.load(h)


def _as_set(clusters):
    return {frozenset((int(i), int(j)) for i, j in np.asarray(c).reshape(-1, 2))
            for c in clusters}


def test_find_clusters_vs_reference():
    from hic3defdr_b200 import ops
    for case in _fixture()['find_clusters']:
        row = np.array(case['row'], dtype=np.int32)
        col = np.array(case['col'], dtype=np.int32)
        order = np.lexsort((col, row))
        got = ops.find_clusters(row[order], col[order])
        want = _as_set(case['clusters'])
        assert len(got) == len(want)
        assert _as_set(got) == want
        # labels: first pixel of the component; sizes at the representatives
        label, size = ops.connected_components(row[order], col[order])
        label, size = label.cpu().numpy(), size.cpu().numpy()
        if len(label):
            assert (label <= np.arange(len(label))).all()
            assert (label[label] == label).all()
            assert size.sum() == len(label)
            assert sorted(size[size > 0]) == sorted(len(c) for c in want)
        for min_size in (2, 5):
            got = ops.find_clusters(row[order], col[order], min_size)
            assert _as_set(got) == {c for c in want if len(c) >= min_size}


def test_unsorted_input_is_rejected():
    from hic3defdr_b200 import ops
    from hic3defdr_b200._native import H3DError
    with pytest.raises(H3DError):
        ops.connected_components(np.array([3, 1], dtype=np.int32),
                                 np.array([4, 2], dtype=np.int32))
    with pytest.raises(H3DError):            # duplicates
        ops.connected_components(np.array([1, 1], dtype=np.int32),
                                 np.array([2, 2], dtype=np.int32))


def test_large_random_set_against_scipy_label():
    """2 M pixels of a 10 kb-sized band: component count and sizes against
    scipy.ndimage.label on the dense mask (4-connectivity)."""
    from scipy import ndimage
    from hic3defdr_b200 import ops
    rng = np.random.default_rng(3)
    n, width = 12000, 200
    i = np.repeat(np.arange(n), width + 1)
    j = i + np.tile(np.arange(width + 1), n)
    keep = (j < n) & (rng.random(len(i)) < 0.55)
    i, j = i[keep].astype(np.int32), j[keep].astype(np.int32)
    label, size = ops.connected_components(i, j)
    label, size = label.cpu().numpy(), size.cpu().numpy()
    dense = np.zeros((n, width + 1), dtype=bool)        # (row, distance) layout
    dense[i, j - i] = True
    # neighbours in (row, col) space: (r, c + 1) -> (r, d + 1); (r + 1, c) ->
    # (r + 1, d - 1): label with that structure on the sheared grid
    structure = np.array([[0, 0, 0], [0, 1, 1], [1, 0, 0]], dtype=bool)
    structure = structure | structure[::-1, ::-1]
    lab, n_comp = ndimage.label(dense, structure=structure)
    assert (size > 0).sum() == n_comp
    want_sizes = np.bincount(lab[dense])[1:]
    assert sorted(size[size > 0]) == sorted(want_sizes)
    # same partition: one reference label per component and vice versa
    ref = lab[i, j - i]
    pairs = np.unique(np.stack([label, ref], axis=1), axis=0)
    assert len(pairs) == n_comp


@pytest.fixture()
def recorded_run(tmp_path):
    """an output directory holding the REFERENCE's recorded arrays of the
    golden run, so that threshold / classify are checked stage-isolated"""
    from hic3defdr_b200 import HiC3DeFDR
    g = np.load(os.path.join(GOLDEN, 'ref_pipeline.npz'))
    chroms = sorted({k.split('_')[1] for k in g.files
                     if k.startswith('qvalues_')})
    outdir = str(tmp_path / 'out')
    design = pd.DataFrame(g['design'].astype(bool), index=['A1', 'A2', 'B1', 'B2'],
                          columns=['A', 'B'])
    h = HiC3DeFDR(raw_npz_patterns=[], bias_patterns=[], chroms=chroms,
                  design=design, outdir=outdir,
                  loop_patterns={'A': 'unused', 'B': 'unused'}, res=10000)
    for c in chroms:
        for name in ('row', 'col', 'disp_idx', 'loop_idx', 'qvalues',
                     'mu_hat_alt'):
            np.save(os.path.join(outdir, '%s_%s.npy' % (name, c)),
                    g['%s_%s' % (name, c)])
    return h, outdir, chroms


def test_threshold_classify_collect_vs_reference(recorded_run):
    from hic3defdr_b200 import clusters as hc
    h, outdir, chroms = recorded_run
    fx = _fixture()['pipeline']
    fdrs, sizes = [0.5, 0.3], [1, 3]
    h.threshold(fdr=fdrs, cluster_size=sizes)
    h.classify(fdr=fdrs, cluster_size=sizes)
    n_checked = 0
    for case in fx:
        tag = '%g_%i_%s' % (case['fdr'], case['cluster_size'], case['chrom'])
        for name in ('sig', 'insig'):
            got = hc.load_clusters(os.path.join(outdir, '%s_%s.json'
                                                % (name, tag)))
            assert _as_set(got) == _as_set(case[name]), (name, tag)
            table = hc.load_cluster_table(os.path.join(
                outdir, '%s_%s.tsv' % (name, tag)))
            assert len(table) == len(case[name])
            assert sorted(table['cluster_size']) == \
                sorted(len(c) for c in case[name])
        if case['classes'] is not None:
            for cond, want in zip(('A', 'B'), case['classes']):
                got = hc.load_clusters(os.path.join(outdir, '%s_%s.json'
                                                    % (cond, tag)))
                assert _as_set(got) == _as_set(want), (cond, tag)
                n_checked += len(want)
    assert n_checked > 0
    h.collect(fdr=0.5, cluster_size=1)
    res = pd.read_csv(os.path.join(outdir, 'results_0.5_1.tsv'), sep='\t',
                      index_col=0)
    want_rows = sum(len(c['insig']) + sum(len(k) for k in c['classes'] or [])
                    for c in fx if c['fdr'] == 0.5 and c['cluster_size'] == 1)
    assert len(res) == want_rows
    assert set(res['classification']) <= {'constitutive', 'A', 'B'}
    assert list(res['us_chrom']) == sorted(res['us_chrom'],
                                           key=hc.natural_sort_key)
