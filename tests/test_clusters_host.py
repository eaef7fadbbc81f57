"""CPU: the host side of threshold / classify / collect
(hic3defdr_b200/clusters.py) against the reference's doctest vectors
(hic3defdr/util/cluster_table.py:43-56, 100-118; util/clusters.py:343-347)
and the fixture recorded from the reference (tests/golden/ref_clusters.json:
structure only -- the clustering itself runs on the GPU)."""
import json
import os

import numpy as np
import pandas as pd

from hic3defdr_b200 import clusters as hc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_clusters_to_table_doctest():
    df = hc.clusters_to_table([[(1, 2), (1, 1)], [(4, 4), (3, 4)]], 'chrX',
                              10000)
    first = df.iloc[0, :]
    assert first.name == 'chrX:10000-20000_chrX:10000-30000'
    assert first.to_dict() == dict(
        us_chrom='chrX', us_start=10000, us_end=20000, ds_chrom='chrX',
        ds_start=10000, ds_end=30000, cluster_size=2,
        cluster=[[1, 2], [1, 1]])
    assert list(df.columns) == hc.COLUMN_ORDER[1:]


def test_sort_cluster_table_doctest():
    clusters = [[(4, 4), (3, 4)], [(1, 2), (1, 1)]]
    df = pd.concat([hc.clusters_to_table(clusters, c, 10000)
                    for c in ('chrX', 'chr11', 'chr2', 'chr1')], axis=0)
    assert list(hc.sort_cluster_table(df).index) == [
        'chr1:10000-20000_chr1:10000-30000',
        'chr1:30000-50000_chr1:40000-50000',
        'chr2:10000-20000_chr2:10000-30000',
        'chr2:30000-50000_chr2:40000-50000',
        'chr11:10000-20000_chr11:10000-30000',
        'chr11:30000-50000_chr11:40000-50000',
        'chrX:10000-20000_chrX:10000-30000',
        'chrX:30000-50000_chrX:40000-50000']


def test_loop_id_doctest_and_file_round_trips(tmp_path):
    cluster = [(4, 5), (3, 4), (3, 5), (3, 6)]
    assert hc.cluster_to_loop_id(cluster, 'chrX', 10000) == \
        'chrX:30000-50000_chrX:40000-70000'
    path = str(tmp_path / 'c.json')
    hc.save_clusters([np.array(cluster), np.array([[7, 9]])], path)
    assert json.load(open(path)) == [[[4, 5], [3, 4], [3, 5], [3, 6]],
                                     [[7, 9]]]
    back = hc.load_clusters(path)
    assert [c.tolist() for c in back] == [[list(p) for p in cluster], [[7, 9]]]
    hc.save_clusters([], path)
    assert hc.load_clusters(path) == []
    table = hc.clusters_to_table([cluster], 'chr3', 5000)
    tsv = str(tmp_path / 't.tsv')
    table.to_csv(tsv, sep='\t')
    loaded = hc.load_cluster_table(tsv)
    assert loaded['cluster'].iloc[0] == [list(p) for p in cluster]
    assert loaded.index[0] == 'chr3:15000-25000_chr3:20000-35000'
    empty = hc.clusters_to_table([], 'chr3', 5000)
    assert len(empty) == 0 and list(empty.columns) == hc.COLUMN_ORDER[1:]


def test_fixture_is_a_partition_into_4_connected_sets():
    """the recorded reference clusters are what the GPU test expects them to
    be: a partition of the input pixels into 4-connected sets, no two of which
    touch"""
    with open(os.path.join(GOLDEN, 'ref_clusters.json')) as h:
        fx = json.load(h)
    assert len(fx['find_clusters']) == 6 and len(fx['pipeline']) == 8
    for case in fx['find_clusters']:
        px = set(zip(case['row'], case['col']))
        seen = {}
        for k, c in enumerate(case['clusters']):
            for p in c:
                assert tuple(p) in px and tuple(p) not in seen
                seen[tuple(p)] = k
        assert len(seen) == len(px)
        for (i, j), k in seen.items():
            for nb in ((i + 1, j), (i, j + 1)):
                if nb in seen:
                    assert seen[nb] == k
