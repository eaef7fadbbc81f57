"""
TEST / BASELINE INFRASTRUCTURE -- times the UNMODIFIED reference
(``hic3defdr.HiC3DeFDR.run_to_qvalues``, files in -> files out, its own
process pools) on a bounded sample of the bench workload.  Only bench.py's
``cpu_baseline`` and ``--impl reference`` legs call this.

The reference tree is imported through oracle/refrun.py from /root/reference
(build container) or from oracle/_ref (the copy ``__graft_entry__.build()``
makes next to the built libraries so that it travels to the GPU box;
git-ignored).  If neither exists the oracle port (oracle/parallel.py) is timed
instead and the record says ``kind: "port"``.
"""
import os
import shutil
import tempfile
import time

import numpy as np

DIST_MAX = 200
FULL_CONFIG0 = ('chr18', 'chr19')       # BASELINE.json configs[0]


def sample_chroms(n_bins_total):
    """Bounded sample of the mouse 10 kb workload with ~n_bins_total bins:
    chr19 truncated to its first bins, then chr19 + truncated chr18, at most
    the whole of BASELINE configs[0] (chr18 + chr19)."""
    from hic3defdr_b200.synth import MM10_10KB
    n19, n18 = MM10_10KB['chr19'], MM10_10KB['chr18']
    n = int(max(DIST_MAX + 50, n_bins_total))
    if n <= n19:
        return {'chr19': n}
    if n < n19 + n18:
        return {'chr19': n19, 'chr18': max(DIST_MAX + 50, n - n19)}
    return {'chr18': n18, 'chr19': n19}


def describe(chroms):
    from hic3defdr_b200.synth import MM10_10KB
    parts = []
    for c, n in chroms.items():
        parts.append('%s (%d bins)' % (c, n) if n == MM10_10KB[c] else
                     '%s truncated to its first %d of %d bins'
                     % (c, n, MM10_10KB[c]))
    full = all(n == MM10_10KB[c] for c, n in chroms.items()) and \
        sorted(chroms) == sorted(FULL_CONFIG0)
    return ('%s of the synthetic mouse 10 kb generator, 2-vs-2 reps, dist cap '
            '%d bins, files in -> files out%s'
            % (' + '.join(parts), DIST_MAX,
               ' (= BASELINE configs[0], full size)' if full else ''))


class Dataset(object):
    """The sample written once in the reference's input formats (npz + bias
    text) under a temporary directory (RAM-backed when /dev/shm exists)."""

    def __init__(self, chroms):
        from hic3defdr_b200.synth import write_dataset
        base = '/dev/shm' if os.path.isdir('/dev/shm') and \
            os.access('/dev/shm', os.W_OK) else None
        self.root = tempfile.mkdtemp(prefix='h3d_ref_', dir=base)
        self.chroms = dict(chroms)
        self.kw = write_dataset(self.root, self.chroms, n_reps=4,
                                dist_max=DIST_MAX, config=1, amp=300.0)
        self.kw.pop('loop_patterns')
        self.n_out = 0

    def close(self):
        shutil.rmtree(self.root, ignore_errors=True)


def kind():
    from oracle import refrun
    return 'reference' if refrun.available() else 'port'


def run_once(ds, n_threads=-1):
    """One timed ``run_to_qvalues`` of the reference over the dataset.
    Returns (union pixels, seconds)."""
    from oracle import refrun
    if not refrun.available():
        return _run_port(ds, n_threads)
    Ref = refrun.reference_class()
    ds.n_out += 1
    outdir = os.path.join(ds.root, 'out%d' % ds.n_out)
    h = Ref(outdir=outdir, dist_thresh_max=DIST_MAX, **ds.kw)
    t0 = time.perf_counter()
    h.run_to_qvalues(n_threads=n_threads, verbose=False)
    dt = time.perf_counter() - t0
    n_px = sum(np.load(os.path.join(outdir, 'row_%s.npy' % c),
                       mmap_mode='r').shape[0] for c in ds.chroms)
    shutil.rmtree(outdir, ignore_errors=True)
    return int(n_px), dt


def _run_port(ds, n_threads):
    import scipy.sparse as sparse
    from oracle import parallel
    design = np.asarray(ds.kw['design'].values).astype(bool)
    t0 = time.perf_counter()
    ins = []
    for c in ds.chroms:
        mats = [sparse.load_npz(p.replace('<chrom>', c)).tocsr()
                for p in ds.kw['raw_npz_patterns']]
        bias = np.array([np.loadtxt(p.replace('<chrom>', c))
                         for p in ds.kw['bias_patterns']]).T
        ins.append((mats, bias))
    res = parallel.run_to_qvalues(ins, design, dist_max=DIST_MAX,
                                  n_threads=n_threads)
    dt = time.perf_counter() - t0
    return sum(len(st['row']) for st in res['chroms']), dt
