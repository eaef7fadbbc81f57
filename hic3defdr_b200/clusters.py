"""
Host side of ``threshold()`` / ``classify()`` / ``collect()``: the reference's
cluster file formats (hic3defdr/util/clusters.py:116-193 sparse JSON,
hic3defdr/util/cluster_table.py:14-190 TSV tables).  The clustering itself
(connected components) runs on the GPU (``ops.connected_components``); this
module only converts between pixel arrays and the files.

A cluster is a (k, 2) integer array of [row, col] pairs here; the reference's
"list of set of tuple" is accepted wherever clusters are read.
"""
import json
import re

import numpy as np
import pandas as pd

COLUMN_ORDER = ['loop_id', 'us_chrom', 'us_start', 'us_end', 'ds_chrom',
                'ds_start', 'ds_end', 'cluster_size', 'cluster']


def save_clusters(clusters, outfile):
    """util/clusters.py:116-136: ``[[[i, j], ...], ...]``."""
    with open(outfile, 'w') as handle:
        json.dump([np.asarray(c, dtype=np.int64).reshape(-1, 2).tolist()
                   for c in clusters], handle)


def load_clusters(infile):
    """util/clusters.py:176-193 -> list of (k, 2) int64 arrays."""
    with open(infile, 'r') as handle:
        return [np.asarray(c, dtype=np.int64).reshape(-1, 2)
                for c in json.load(handle)]


def cluster_to_loop_id(cluster, chrom, resolution):
    """util/clusters.py:321-357: "chr:start-end_chr:start-end" of the bounding
    box of the cluster."""
    c = np.asarray(cluster, dtype=np.int64).reshape(-1, 2)
    return '%s:%s-%s_%s:%s-%s' % (
        chrom, c[:, 0].min() * resolution, (c[:, 0].max() + 1) * resolution,
        chrom, c[:, 1].min() * resolution, (c[:, 1].max() + 1) * resolution)


def natural_sort_key(s):
    """lib5c.util.primers.natural_sort_key as used at
    util/cluster_table.py:128-130: digit runs compare as integers, so that
    chr2 < chr11 < chrX (pinned by the doctest of sort_cluster_table,
    util/cluster_table.py:100-118)."""
    return [int(t) if t.isdigit() else t.lower()
            for t in re.split(r'(\d+)', str(s))]


def sort_cluster_table(cluster_table):
    """util/cluster_table.py:84-147: by upstream then downstream anchor,
    chromosomes in natural order."""
    chroms = sorted(set(cluster_table['us_chrom'].unique()) |
                    set(cluster_table['ds_chrom'].unique()),
                    key=natural_sort_key)
    idx = {c: i for i, c in enumerate(chroms)}
    t = cluster_table.copy()
    t['us_chrom_idx'] = t['us_chrom'].map(idx)
    t['ds_chrom_idx'] = t['ds_chrom'].map(idx)
    order = ['us_chrom_idx', 'us_start', 'us_end', 'ds_chrom_idx', 'ds_start',
             'ds_end']
    return t.sort_values(order, kind='stable') \
        .drop(columns=['us_chrom_idx', 'ds_chrom_idx'])


def clusters_to_table(clusters, chrom, res):
    """util/cluster_table.py:14-81: one row per cluster, indexed by loop id,
    with the bounding-box anchors, the size and the pixels."""
    rows = []
    for cluster in clusters:
        c = np.asarray(list(cluster), dtype=np.int64).reshape(-1, 2)
        rows.append({
            'loop_id': cluster_to_loop_id(c, chrom, res),
            'us_chrom': chrom,
            'us_start': int(c[:, 0].min() * res),
            'us_end': int((c[:, 0].max() + 1) * res),
            'ds_chrom': chrom,
            'ds_start': int(c[:, 1].min() * res),
            'ds_end': int((c[:, 1].max() + 1) * res),
            'cluster_size': len(c),
            'cluster': c.tolist(),
        })
    return sort_cluster_table(
        pd.DataFrame(rows, columns=COLUMN_ORDER).set_index('loop_id'))


def load_cluster_table(table_filename):
    """util/cluster_table.py:150-189."""
    df = pd.read_csv(table_filename, sep='\t', index_col=0)
    df['cluster'] = df['cluster'].apply(
        lambda s: json.loads(s.replace('(', '[').replace('{', '[')
                             .replace(')', ']').replace('}', ']')))
    return df


def pixel_keys(row, col):
    return (np.asarray(row, dtype=np.int64) << 32) | \
        np.asarray(col, dtype=np.int64)
