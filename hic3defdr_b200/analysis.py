"""
``HiC3DeFDR``: drop-in for the reference class on the ``run_to_qvalues`` path.

Same constructor kwargs, ``<outdir>/pickle`` and ``HiC3DeFDR.load``
(hic3defdr/analysis/constructor.py:62-86, analysis/core.py:15-33), same step
methods and kwargs (analysis/analysis.py:28-364), same per-chromosome ``.npy``
outputs (analysis/core.py:62-253; shapes and dtypes in DESIGN.md).  The
arithmetic of every step runs in libh3d (hand-written sm_100a CUDA,
include/h3d.h); this module only parses inputs, owns the device buffers
(torch tensors) and writes the files.  There is no CPU fallback.

Differences a caller can observe (all documented in DESIGN.md):
  * ``n_threads`` is accepted for signature compatibility; it sizes the host
    I/O thread pool, the compute is on the GPU;
  * results of a step are also kept on the device (``self._cache``) so that
    the next step of ``run_to_qvalues`` does not re-read them from disk; every
    step still works from the files alone in a fresh process;
  * with ``torch.distributed`` initialised (one process per GPU), chromosomes
    are sharded across ranks (hic3defdr_b200/dist.py); with fewer chromosomes
    than ranks (or ``H3D_SHARD=rows``) every chromosome is sharded by row
    range instead and the ranks write their slices of the same ``.npy`` files
    (outdir must be on a file system all ranks share, as on one node).
"""
import collections
import itertools
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import dill as pickle
import numpy as np
import pandas as pd
import scipy.sparse as sparse
import torch

from hic3defdr_b200 import dist as hdist
from hic3defdr_b200 import clusters as hclusters
from hic3defdr_b200 import engine, hostio, ops, staging


def eprint(*args, **kwargs):
    """stderr printing with ``skip=`` (hic3defdr/util/printing.py:5-19)."""
    if not kwargs.pop('skip', False):
        print(*args, file=sys.stderr, **kwargs)


def check_outdir(path):
    """lib5c.util.system.check_outdir as used at constructor.py:84."""
    d = os.path.dirname(path)
    if d and not os.path.exists(d):
        print('creating directory %s' % d)
        os.makedirs(d, exist_ok=True)


def load_clusters(infile):
    """hic3defdr/util/clusters.py:177-193."""
    with open(infile, 'r') as handle:
        return [set(tuple(e) for e in cluster) for cluster in json.load(handle)]


def _loadtxt(path):
    try:
        return pd.read_csv(path, header=None, sep=r'\s+', comment='#',
                           dtype=np.float64).values[:, 0]
    except Exception:
        return np.loadtxt(path)


class _AsyncWriter(object):
    """Background ``.npy`` writers (with the
    arithmetic on the GPU the wall time of a run is file I/O).  ``submit``
    returns at once; a worker thread waits for the kernels that produce the
    tensor (CUDA event), copies it to the host on its own stream and writes the
    file while the main thread goes on with the next chromosome.  ``wait``
    blocks until every file is on disk and re-raises the first failure."""

    # device tensors go through two pinned staging buffers per writer thread,
    # CHUNK bytes at a time: the copy of a chunk (PCIe rate, no page faults of
    # a fresh host allocation) overlaps the write() of the previous one, and
    # the pinned footprint stays bounded whatever the array sizes
    CHUNK = 16 << 20

    def __init__(self, n_threads, staged=None):
        self.pool = ThreadPoolExecutor(max(1, n_threads))
        self.futures = []
        self.streams = {}
        self.staging = {}
        self.lock = threading.Lock()
        self.staged = (os.environ.get('H3D_WRITER', 'staged') == 'staged') \
            if staged is None else staged

    def _thread_stream(self):
        key = threading.get_ident()
        with self.lock:
            stream = self.streams.get(key)
            if stream is None:
                stream = self.streams[key] = torch.cuda.Stream()
        return stream

    def _thread_staging(self):
        key = threading.get_ident()
        with self.lock:
            st = self.staging.get(key)
        if st is None:
            st = [(torch.empty(self.CHUNK, dtype=torch.uint8, pin_memory=True),
                   torch.cuda.Event()) for _ in range(2)]
            with self.lock:
                self.staging[key] = st
        return st

    @staticmethod
    def write_npy_chunks(path, dtype, shape, chunks):
        """``np.save(path, a)`` of a C-ordered array given as an iterable of
        buffers holding its bytes in order (same file, byte for byte)."""
        with open(path, 'wb') as handle:
            np.lib.format.write_array_header_1_0(handle, {
                'descr': np.lib.format.dtype_to_descr(np.dtype(dtype)),
                'fortran_order': False, 'shape': tuple(int(v) for v in shape)})
            for c in chunks:
                handle.write(c)

    def _device_chunks(self, flat, stream):
        """the bytes of a contiguous device tensor, CHUNK at a time, through
        the calling thread's staging buffers (views: valid until the next
        item is asked for)."""
        staging = self._thread_staging()
        n, off, k, pending = flat.numel(), 0, 0, []
        while off < n or pending:
            if off < n and len(pending) < 2:
                m = min(self.CHUNK, n - off)
                buf, ev = staging[k % 2]
                with torch.cuda.stream(stream):
                    buf[:m].copy_(flat[off:off + m], non_blocking=True)
                    ev.record(stream)
                pending.append((buf, m, ev))
                off += m
                k += 1
                continue
            buf, m, ev = pending.pop(0)
            ev.synchronize()
            yield memoryview(buf.numpy())[:m]

    def _job(self, data, path, event, device):
        if isinstance(data, torch.Tensor):
            if data.is_cuda:
                torch.cuda.set_device(device)
                stream = self._thread_stream()
                stream.wait_event(event)
                if self.staged and data.is_contiguous():
                    np_dtype = torch.empty(0, dtype=data.dtype).numpy().dtype
                    flat = data.reshape(-1).view(torch.uint8)
                    self.write_npy_chunks(path, np_dtype, data.shape,
                                          self._device_chunks(flat, stream))
                    return
                with torch.cuda.stream(stream):
                    host = data.to('cpu', non_blocking=False)
                stream.synchronize()
                data = host
            data = data.numpy()
        np.save(path, data)

    def submit(self, data, path):
        event, device = None, None
        if isinstance(data, torch.Tensor) and data.is_cuda:
            device = data.device.index
            event = torch.cuda.Event()
            event.record(torch.cuda.current_stream())
        self.futures.append(self.pool.submit(self._job, data, path, event,
                                             device))

    def wait(self):
        futures, self.futures = self.futures, []
        for f in futures:
            f.result()


class HiC3DeFDR(object):
    """See ``hic3defdr.analysis.constructor.HiC3DeFDR`` for the attributes."""

    def __init__(self, raw_npz_patterns, bias_patterns, chroms, design, outdir,
                 dist_thresh_min=4, dist_thresh_max=200, bias_thresh=0.1,
                 mean_thresh=1.0, loop_patterns=None, res=None,
                 _write_pickle=True):
        self.raw_npz_patterns = raw_npz_patterns
        self.bias_patterns = bias_patterns
        self.chroms = chroms
        if type(design) == str:
            self.design = pd.read_csv(design, index_col=0)
        else:
            self.design = design
        self.outdir = outdir
        self.dist_thresh_min = dist_thresh_min
        self.dist_thresh_max = dist_thresh_max
        self.bias_thresh = bias_thresh
        self.mean_thresh = mean_thresh
        self.loop_patterns = loop_patterns
        self.res = res
        state = self.__dict__.copy()
        del state['outdir']
        if _write_pickle:
            if hdist.rank() == 0:
                check_outdir(self.picklefile)
                # written under a temporary name and renamed: a concurrent
                # reader never sees a truncated file
                tmp = '%s.tmp%d' % (self.picklefile, os.getpid())
                with open(tmp, 'wb') as handle:
                    pickle.dump(state, handle, -1)
                os.replace(tmp, self.picklefile)
            hdist.barrier()
        self._cache = {}
        self._shards = {}
        self._writer = None
        self._defer_writes = False
        self.timings = {}

    # ---------------------------------------------------------------- core
    @property
    def picklefile(self):
        return '%s/pickle' % self.outdir

    @classmethod
    def load(cls, outdir):
        # the reference re-dumps the pickle it has just read
        # (constructor.py:82-86 runs inside load); with several ranks that
        # rewrite could truncate the file under a slower reader, so load()
        # only reads
        with open('%s/pickle' % outdir, 'rb') as handle:
            state = pickle.load(handle)
        return cls(outdir=outdir, _write_pickle=False, **state)

    def _design(self):
        return np.asarray(self.design.values).astype(bool)

    def load_bias(self, chrom):
        """analysis/core.py:35-60; returns a numpy (n_bins, n_reps) matrix."""
        return self._bias_device(chrom).cpu().numpy()

    def _bias_device(self, chrom):
        c = self._cache.setdefault(chrom, {})
        if 'bias' not in c:
            raw = np.array([_loadtxt(p.replace('<chrom>', chrom))
                            for p in self.bias_patterns]).T
            c['bias'] = ops.filter_bias(np.ascontiguousarray(raw),
                                        self.bias_thresh)
        return c['bias']

    # which pixel set the rows of a per-pixel output are aligned with
    # (docs/data_layout.md of the reference): the union pixels, the tested
    # ones (row[disp_idx]) or the loop pixels among those
    _ALIGNED_WITH = dict(
        [(n, 'union') for n in ('raw', 'size_factors', 'scaled', 'disp_idx')] +
        [(n, 'tested') for n in ('loop_idx', 'disp', 'mu_hat_null',
                                 'mu_hat_alt', 'llr', 'pvalues')] +
        [('qvalues', 'loops')])
    _NOT_PER_PIXEL = ('row', 'col', 'bias', 'cov_per_bin', 'disp_per_bin')

    def _data_file(self, name, chrom=None):
        return '%s/%s.npy' % (self.outdir, name) if chrom is None else \
            '%s/%s_%s.npy' % (self.outdir, name, chrom)

    def _read_rows(self, path, keep=None):
        """the rows ``keep`` (boolean mask) of a saved array, all if None; a
        masked read maps the file instead of loading it whole"""
        if keep is None:
            return np.load(path)
        return np.load(path, mmap_mode='r')[keep]

    def _pixel_coordinates(self, name, chrom):
        """(row, col) of the pixels the rows of ``name`` belong to."""
        level = self._ALIGNED_WITH[name]
        keep = None
        if level != 'union':
            keep = self.load_data('disp_idx', chrom)
            if level == 'loops':
                keep = self._chain_masks(keep,
                                         self.load_data('loop_idx', chrom))
        return (self._read_rows(self._data_file('row', chrom), keep),
                self._read_rows(self._data_file('col', chrom), keep))

    @staticmethod
    def _chain_masks(outer, inner):
        """mask over everything that selects the ``inner``-selected ones of
        the ``outer``-selected entries (``x[outer][inner] == x[chained]``)"""
        chained = np.zeros(len(outer), dtype=bool)
        chained[np.flatnonzero(outer)[inner]] = True
        return chained

    def load_data(self, name, chrom=None, idx=None, rep=None, cond=None,
                  coo=False):
        """The reference's loader (analysis/core.py:62-195): one saved array,
        whole (``chrom=None``), of one chromosome, or of all of them
        concatenated (``chrom='all'``: returns (data, offsets)); ``idx``: a
        boolean row mask or a (mask, mask-of-the-selected) pair; ``rep`` /
        ``cond``: one column by replicate / condition name; ``coo=True``:
        (row, col, data) of one chromosome."""
        self._flush_writes()
        if name == 'loop_idx' and self.loop_patterns is None and idx is None \
                and chrom != 'all':
            # no loops given: every tested pixel counts as one
            return np.ones(int(self.load_data('disp_idx', chrom).sum()),
                           dtype=bool)
        column = None
        if rep is not None:
            column = list(self.design.index).index(rep)
        elif cond is not None:
            column = list(self.design.columns).index(cond)
        pick = (lambda a: a) if column is None else (lambda a: a[:, column])
        if coo:
            if chrom == 'all' or idx is not None:
                raise ValueError("cannot pass coo=True with chrom='all' or idx")
            if name in self._NOT_PER_PIXEL:
                raise ValueError('data with name %s cannot be loaded as COO'
                                 % name)
            if name not in self._ALIGNED_WITH:
                raise ValueError('data name %s not recognized' % name)
            row, col = self._pixel_coordinates(name, chrom)
            return row, col, pick(self.load_data(name, chrom))
        keep = self._chain_masks(*idx) if type(idx) == tuple else idx
        if chrom != 'all':
            return pick(self._read_rows(self._data_file(name, chrom), keep))
        # genome-wide: ``keep`` runs over the concatenation of the chromosomes
        pieces, start = [], 0
        for c in self.chroms:
            path = self._data_file(name, c)
            if keep is None:
                pieces.append(np.load(path))
            else:
                rows = np.load(path, mmap_mode='r')
                pieces.append(rows[keep[start:start + rows.shape[0]]])
                start += rows.shape[0]
        offsets = np.concatenate(
            [[0], np.cumsum([p.shape[0] for p in pieces])]).astype(np.int64)
        return pick(np.concatenate(pieces)), offsets

    def save_data(self, data, name, chrom=None):
        """analysis/core.py:197-218."""
        if isinstance(data, torch.Tensor):
            data = data.cpu().numpy()
        if chrom is None:
            np.save('%s/%s.npy' % (self.outdir, name), data)
        elif isinstance(chrom, np.ndarray):
            for i, c in enumerate(self.chroms):
                self.save_data(data[chrom[i]:chrom[i + 1]], name, c)
        else:
            np.save('%s/%s_%s.npy' % (self.outdir, name, chrom), data)

    def load_disp_fn(self, cond):
        """analysis/core.py:220-237."""
        with open('%s/disp_fn_%s.pickle' % (self.outdir, cond), 'rb') as h:
            return pickle.load(h)

    def save_disp_fn(self, cond, disp_fn):
        """analysis/core.py:239-253."""
        with open('%s/disp_fn_%s.pickle' % (self.outdir, cond), 'wb') as h:
            return pickle.dump(disp_fn, h, -1)

    # ------------------------------------------------------------ helpers
    def _row_sharded(self):
        """True when every rank works on a row range of every chromosome
        (BASELINE config 4) rather than on whole
        chromosomes: ``H3D_SHARD=rows``, or by default when there are fewer
        chromosomes than ranks (``H3D_SHARD=chroms`` forces the other way)."""
        ws = hdist.world_size()
        if ws == 1:
            return False
        mode = os.environ.get('H3D_SHARD', 'auto')
        if mode not in ('auto', 'rows', 'chroms'):
            raise ValueError('H3D_SHARD must be auto, rows or chroms')
        return mode == 'rows' or (mode == 'auto' and len(self.chroms) < ws)

    def _my_chroms(self):
        if self._row_sharded():
            return list(self.chroms)
        return hdist.shard_chroms(self.chroms, self._chrom_weight)

    def _shard_bounds(self, chrom):
        """(lo, hi, dlo, dhi): this rank's slice of the union-aligned and of
        the disp_idx-aligned arrays of a row-sharded chromosome.  Known from
        ``prepare_data`` in the same process; in a fresh process the pixels are
        cut evenly (the later steps accept any contiguous partition)."""
        b = self._shards.get(chrom)
        if b is None:
            self._flush_writes()
            ws, me = hdist.world_size(), hdist.rank()
            disp_idx = np.load('%s/disp_idx_%s.npy' % (self.outdir, chrom),
                               mmap_mode='r')
            n = disp_idx.shape[0]
            lo, hi = n * me // ws, n * (me + 1) // ws
            dlo = int(np.count_nonzero(disp_idx[:lo]))
            b = (lo, hi, dlo, dlo + int(np.count_nonzero(disp_idx[lo:hi])))
            self._shards[chrom] = b
        return b

    def _save_sharded(self, items):
        """items: list of (tensor, name, chrom), each this rank's slice (rank
        order = array order) of one output array: rank 0 creates the ``.npy``
        at its full size, every rank writes its rows in place."""
        ws, me = hdist.world_size(), hdist.rank()
        host = [(d.cpu().numpy() if isinstance(d, torch.Tensor) else
                 np.asarray(d)) for d, _, _ in items]
        lens = hdist._all_gather_counts(
            np.array([h.shape[0] for h in host], dtype=np.int64))  # (ws, items)
        paths = ['%s/%s_%s.npy' % (self.outdir, n, c) for _, n, c in items]
        if me == 0:
            for h, p, tot in zip(host, paths, lens.sum(axis=0)):
                m = np.lib.format.open_memmap(
                    p, mode='w+', dtype=h.dtype, shape=(int(tot),) + h.shape[1:])
                del m
        hdist.barrier()
        for k, (h, p) in enumerate(zip(host, paths)):
            if h.shape[0]:
                m = np.load(p, mmap_mode='r+')
                a = int(lens[:me, k].sum())
                m[a:a + h.shape[0]] = h
                m.flush()
                del m
        hdist.barrier()

    def _chrom_weight(self, chrom):
        try:
            return os.path.getsize(
                self.raw_npz_patterns[0].replace('<chrom>', chrom))
        except OSError:
            return 1

    def _add_time(self, key, t0):
        """host seconds since ``t0`` added to ``self.timings[key]`` (where the
        wall time of a files-in -> files-out run goes); returns now."""
        now = time.perf_counter()
        self.timings[key] = self.timings.get(key, 0.0) + now - t0
        return now

    def _io_threads(self, n_threads):
        if n_threads is None or n_threads == 0:
            return 1
        if n_threads < 0:
            return max(1, min(os.cpu_count() or 1, 16))
        return n_threads

    def _save_many(self, items, n_threads=-1):
        """items: list of (tensor-or-array, name, chrom).  The files are
        written by background threads; they are complete when the calling step
        returns (or, inside ``run_to_qvalues``, when that returns)."""
        if self._writer is None:
            self._writer = _AsyncWriter(self._io_threads(n_threads))
        for data, name, chrom in items:
            path = '%s/%s.npy' % (self.outdir, name) if chrom is None else \
                '%s/%s_%s.npy' % (self.outdir, name, chrom)
            self._writer.submit(data, path)
        if not self._defer_writes:
            self._writer.wait()

    def _flush_writes(self):
        if self._writer is not None:
            self._writer.wait()

    def _submit_inputs(self, pool, chrom):
        """Queues the host side of ``prepare_data`` for one chromosome on
        ``pool``, one job per file (bias text parse, npz inflate -- zlib
        releases the GIL -- and the loop cluster files); returns a function
        that waits for them and gives (bias_raw, mats, loop_pixels)."""
        bias = [pool.submit(_loadtxt, p.replace('<chrom>', chrom))
                for p in self.bias_patterns]
        # H3D_PIN_INPUTS=1: the I/O threads also copy the matrices to
        # page-locked memory and the main thread only queues asynchronous
        # uploads.  Worth it for a process that makes many runs (genome scale,
        # B200: 0.78 -> 0.63 s per run); off by default because the first run
        # of a process pays ~3 s for the page-locked allocations.
        device = None
        if os.environ.get('H3D_PIN_INPUTS', '0') not in ('', '0') and \
                torch.cuda.is_available():
            device = torch.cuda.current_device()

        def load_matrix(path):
            m = hostio.load_npz(path).tocsr()
            # O(nnz) scan, cached on the matrix: ops.DeviceCSR asks for it on
            # the main thread
            m.has_canonical_format
            return m if device is None else hostio.pin_csr(m, device)
        mats = [pool.submit(load_matrix, p.replace('<chrom>', chrom))
                for p in self.raw_npz_patterns]
        loops = [pool.submit(load_clusters, pattern.replace('<chrom>', chrom))
                 for pattern in self.loop_patterns.values()] \
            if self.loop_patterns else None

        def collect():
            bias_raw = np.ascontiguousarray(
                np.array([f.result() for f in bias]).T)
            loop_pixels = None
            if loops is not None:
                loop_pixels = set().union(
                    *sum((f.result() for f in loops), []))
            return bias_raw, [f.result() for f in mats], loop_pixels
        return collect

    def _load_inputs(self, chrom, n_threads=-1):
        """Host side of ``prepare_data`` for one chromosome: bias vectors,
        replicate matrices (each npz inflated once, in parallel) and loop
        pixels."""
        with ThreadPoolExecutor(self._io_threads(n_threads)) as pool:
            return self._submit_inputs(pool, chrom)()

    def _prefetched_inputs(self, chroms, n_threads=-1):
        """Yields (chrom, inputs) in order while the files of the following
        chromosomes are read on a shared pool of I/O threads: as many
        chromosomes in flight as the threads can serve at one file each
        (with the arithmetic on the GPU the
        inflate of the input files is the critical path of a run)."""
        n_io = self._io_threads(n_threads)
        depth = max(1, n_io // max(1, len(self.raw_npz_patterns)))
        with ThreadPoolExecutor(n_io) as pool:
            todo = iter(chroms)
            queue = collections.deque(
                (c, self._submit_inputs(pool, c))
                for c in itertools.islice(todo, depth))
            while queue:
                c, collect = queue.popleft()
                inputs = collect()
                nxt = next(todo, None)
                if nxt is not None:
                    queue.append((nxt, self._submit_inputs(pool, nxt)))
                yield c, inputs
                del inputs

    def _chrom_state(self, chrom, names):
        """device tensors of one chromosome, from the cache or from disk."""
        c = self._cache.setdefault(chrom, {})
        for name in names:
            if name in c:
                continue
            self._flush_writes()
            if name == 'bias':
                self._bias_device(chrom)
            elif name == 'disp_index':
                di = self._chrom_state(chrom, ['disp_idx'])['disp_idx']
                c['disp_index'] = ops.mask_to_index(di)
            elif self._row_sharded():
                lo, hi, dlo, dhi = self._shard_bounds(chrom)
                arr = np.load('%s/%s_%s.npy' % (self.outdir, name, chrom),
                              mmap_mode='r')
                if name in ('row', 'col', 'raw', 'scaled', 'disp_idx') or \
                        (name == 'size_factors' and arr.ndim == 2):
                    arr = arr[lo:hi]
                elif name != 'size_factors':
                    arr = arr[dlo:dhi]
                arr = np.ascontiguousarray(arr)
                if arr.dtype == np.bool_:
                    arr = arr.view(np.uint8)
                c[name] = ops.dev(arr)
            else:
                arr = self.load_data(name, chrom)
                if arr.dtype == np.bool_:
                    arr = arr.view(np.uint8)
                c[name] = ops.dev(arr)
        return c

    def free_device_cache(self):
        self._cache = {}
        torch.cuda.empty_cache()

    # ------------------------------------------------------- prepare_data
    def prepare_data(self, chrom=None, norm='conditional_mor', n_bins=-1,
                     n_threads=-1, verbose=True, _inputs=None):
        """analysis/analysis.py:28-133."""
        if n_bins == -1:
            n_bins = int(self.dist_thresh_max / 5)
        if norm not in ops.NORMS:
            raise KeyError(norm)
        if chrom is None:
            # the following chromosomes' files are read (npz inflate, text
            # parse) while this one is on the GPU, and this one's outputs are
            # written while the next ones are computed
            chroms = self._my_chroms()
            defer, self._defer_writes = self._defer_writes, True
            try:
                t_wait = time.perf_counter()
                for c, inputs in self._prefetched_inputs(chroms, n_threads):
                    self._add_time('prepare/wait_for_files', t_wait)
                    self.prepare_data(chrom=c, norm=norm, n_bins=n_bins,
                                      n_threads=n_threads, verbose=False,
                                      _inputs=inputs)
                    del inputs
                    t_wait = time.perf_counter()
            finally:
                self._defer_writes = defer
            if not defer:
                self._flush_writes()
            hdist.barrier()
            return
        eprint('preparing data for chrom %s' % chrom)
        eprint('  loading bias', skip=not verbose)
        self._cache.pop(chrom, None)
        eprint('  computing union pixel set', skip=not verbose)
        bias_raw, mats, loop_pixels = _inputs if _inputs is not None else \
            self._load_inputs(chrom, n_threads)
        sharded = self._row_sharded()
        if sharded:
            bounds = hdist.row_ranges(staging.row_weights(mats))
            me = hdist.rank()
            mats = staging.shard_rows(mats, int(bounds[me]),
                                      int(bounds[me + 1]))
        t0 = time.perf_counter()
        csr = ops.DeviceCSR(mats)
        del mats
        t0 = self._add_time('prepare/upload', t0)
        if self.loop_patterns:
            eprint('  making loop_idx', skip=not verbose)
        eprint('  loading raw data', skip=not verbose)
        eprint('  loading balanced data', skip=not verbose)
        eprint('  computing size factors', skip=not verbose)
        eprint('  computing disp_idx', skip=not verbose)
        prepare = engine.prepare_chrom_sharded if sharded else \
            engine.prepare_chrom
        st = prepare(
            csr, bias_raw, self._design(), self.dist_thresh_min,
            self.dist_thresh_max, self.bias_thresh, self.mean_thresh, norm,
            n_bins, loop_pixels)
        del csr
        t0 = self._add_time('prepare/kernels_and_readbacks', t0)
        eprint('  saving data to disk', skip=not verbose)
        to_save = [(st[k].bool() if k in ('disp_idx', 'loop_idx') else st[k],
                    k, chrom)
                   for k in ('loop_idx', 'row', 'col', 'raw', 'size_factors',
                             'scaled', 'disp_idx') if k in st]
        if sharded:
            whole = [t for t in to_save
                     if t[1] == 'size_factors' and t[0].dim() == 1]
            to_save = [t for t in to_save if t not in whole]
            if whole and hdist.rank() == 0:      # (R,) factors: same everywhere
                self._save_many(whole, n_threads)
            counts = hdist._all_gather_counts(np.array(
                [st['row'].numel(), st['disp_index'].numel()]))
            me = hdist.rank()
            lo, dlo = [int(v) for v in counts[:me].sum(axis=0)]
            self._shards[chrom] = (lo, lo + int(counts[me, 0]),
                                   dlo, dlo + int(counts[me, 1]))
            self._save_sharded(to_save)
        else:
            self._save_many(to_save, n_threads)
        st.pop('scaled')
        self._cache[chrom] = st
        self._add_time('prepare/queue_writes', t0)

    # ------------------------------------------------------ estimate_disp
    def estimate_disp(self, estimator='qcml', frac=None, auto_frac_factor=15.,
                      weighted_lowess=True, n_threads=-1):
        """analysis/analysis.py:135-223."""
        eprint('estimating dispersion')
        if not callable(estimator) and estimator not in ops.ESTIMATORS:
            raise ValueError(
                "estimator must be 'qcml', 'cml', 'mme' (device kernels) or a "
                "function (data, f=...) -> dispersion, which is called on the "
                "host with the pooled pixels of every (distance, condition) "
                "bin, analysis/analysis.py:164-165")
        eprint('  loading data')
        mine = self._my_chroms()
        states = [self._chrom_state(
            c, ['row', 'col', 'raw', 'size_factors', 'disp_idx', 'disp_index',
                'bias']) for c in mine]
        disp_per_dist, fns, stats = engine.estimate_disp(
            states, self._design(), self.dist_thresh_max,
            cond_names=list(self.design.columns), estimator=estimator,
            frac=frac, auto_frac_factor=auto_frac_factor,
            weighted_lowess=weighted_lowess, log=eprint)
        self.timings['qcml_stats'] = stats
        eprint('  saving estimated dispersions to disk')
        if hdist.rank() == 0:
            for cond, fn in zip(self.design.columns, fns):
                self.save_disp_fn(cond, fn)
            self.save_data(disp_per_dist, 'disp_per_dist')
        save = self._save_sharded if self._row_sharded() else \
            (lambda items: self._save_many(items, n_threads))
        save([(s['disp'], 'disp', c) for c, s in zip(mine, states)])
        hdist.barrier()

    # ---------------------------------------------------------------- lrt
    def lrt(self, chrom=None, refit_mu=True, n_threads=-1, verbose=True):
        """analysis/analysis.py:225-284."""
        if chrom is None:
            defer, self._defer_writes = self._defer_writes, True
            try:
                for c in self._my_chroms():
                    self.lrt(chrom=c, refit_mu=refit_mu, n_threads=n_threads,
                             verbose=False)
            finally:
                self._defer_writes = defer
            if not defer:
                self._flush_writes()
            hdist.barrier()
            return
        eprint('running LRT for chrom %s' % chrom)
        eprint('  loading data', skip=not verbose)
        s = self._chrom_state(chrom, ['bias', 'size_factors', 'disp_idx',
                                      'disp_index', 'row', 'col', 'raw',
                                      'disp'])
        eprint('  computing LRT results', skip=not verbose)
        engine.lrt_chrom(s, self._design(), refit_mu)
        eprint('  saving results to disk', skip=not verbose)
        save = self._save_sharded if self._row_sharded() else \
            (lambda items: self._save_many(items, n_threads))
        save([(s[k], k, chrom) for k in
              ('pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt')])
        for k in ('llr', 'mu_hat_null', 'mu_hat_alt'):
            s.pop(k)

    # ----------------------------------------------------------------- bh
    def bh(self):
        """analysis/analysis.py:286-303."""
        eprint('applying BH-FDR correction')
        mine = self._my_chroms()
        names = ['pvalues'] + (['loop_idx'] if self.loop_patterns else [])
        states = [self._chrom_state(c, names) for c in mine]
        engine.bh(states, use_loop_idx=bool(self.loop_patterns))
        save = self._save_sharded if self._row_sharded() else self._save_many
        save([(s['qvalues'], 'qvalues', c) for c, s in zip(mine, states)])
        hdist.barrier()

    def run_to_qvalues(self, norm='conditional_mor', n_bins_norm=-1,
                       estimator='qcml', frac=None, auto_frac_factor=15.,
                       weighted_lowess=True, refit_mu=True, n_threads=-1,
                       verbose=True):
        """analysis/analysis.py:305-364."""
        # one wait for the background writers at the very end: the files of a
        # step are written while the next step computes
        self._defer_writes = True
        try:
            self.prepare_data(norm=norm, n_bins=n_bins_norm,
                              n_threads=n_threads)
            self.estimate_disp(
                estimator=estimator, frac=frac,
                auto_frac_factor=auto_frac_factor,
                weighted_lowess=weighted_lowess, n_threads=n_threads)
            self.lrt(refit_mu=refit_mu, n_threads=n_threads)
            self.bh()
        finally:
            self._defer_writes = False
            self._flush_writes()
        hdist.barrier()

    # ------------------------------------------- threshold / classify / collect
    @staticmethod
    def _as_list(v):
        return list(v) if hasattr(v, '__len__') else [v]

    def _cluster_chroms(self):
        """threshold / classify work on whole chromosomes from the files: in a
        row-sharded run rank 0 takes them all."""
        if self._row_sharded():
            return list(self.chroms) if hdist.rank() == 0 else []
        return self._my_chroms()

    def _save_cluster_files(self, clusters, outfile, chrom):
        hclusters.save_clusters(clusters, outfile)
        if self.res is not None:
            hclusters.clusters_to_table(clusters, chrom, self.res).to_csv(
                outfile.replace('.json', '.tsv'), sep='\t')

    def threshold(self, chrom=None, fdr=0.05, cluster_size=3, n_threads=-1):
        """analysis/analysis.py:366-430: pixels with q < fdr ("sig") and with
        q >= fdr ("insig") are clustered separately (4-connected components,
        on the GPU) and clusters below ``cluster_size`` pixels dropped; writes
        ``sig_<fdr>_<size>_<chrom>.json`` / ``insig_...json`` (+ ``.tsv`` when
        ``res`` is set).  A list of FDRs is a sweep over each value (the
        reference passes the whole list to the comparison, analysis.py:411-413,
        which only works for one value)."""
        if chrom is None:
            for c in self._cluster_chroms():
                self.threshold(chrom=c, fdr=fdr, cluster_size=cluster_size)
            hdist.barrier()
            return
        eprint('thresholding and clustering chrom %s' % chrom)
        row, col, qvalues = self.load_data('qvalues', chrom, coo=True)
        r, c, q = ops.dev(row, torch.int32), ops.dev(col, torch.int32), \
            ops.dev(qvalues, torch.float64)
        for f in self._as_list(fdr):
            for name, sel in (('sig', q < f), ('insig', q >= f)):
                rs, cs = r[sel], c[sel]
                label, size = ops.connected_components(rs, cs)
                for s in self._as_list(cluster_size):
                    found = ops.clusters_from_labels(rs, cs, label, size, s)
                    self._save_cluster_files(
                        found, '%s/%s_%g_%i_%s.json'
                        % (self.outdir, name, f, s, chrom), chrom)

    def classify(self, chrom=None, fdr=0.05, cluster_size=3, n_threads=-1):
        """analysis/analysis.py:432-496 + util/classification.py:7-49: the
        pixels of the significant clusters are assigned to the condition with
        the largest ``mu_hat_alt`` and re-clustered per condition; writes
        ``<cond>_<fdr>_<size>_<chrom>.json`` (+ ``.tsv`` when ``res`` is
        set)."""
        if chrom is None:
            for c in self._cluster_chroms():
                self.classify(chrom=c, fdr=fdr, cluster_size=cluster_size)
            hdist.barrier()
            return
        eprint('classifying differential interactions on chrom %s' % chrom)
        disp_idx = self.load_data('disp_idx', chrom)
        loop_idx = self.load_data('loop_idx', chrom)
        row = self.load_data('row', chrom, idx=(disp_idx, loop_idx))
        col = self.load_data('col', chrom, idx=(disp_idx, loop_idx))
        mu_hat_alt = self.load_data('mu_hat_alt', chrom, idx=loop_idx)
        keys = hclusters.pixel_keys(row, col)
        for f in self._as_list(fdr):
            for s in self._as_list(cluster_size):
                infile = '%s/sig_%g_%i_%s.json' % (self.outdir, f, s, chrom)
                if not os.path.isfile(infile):
                    self.threshold(chrom=chrom, fdr=f, cluster_size=s)
                sig = hclusters.load_clusters(infile)
                sig_px = np.concatenate(sig) if sig else \
                    np.zeros((0, 2), dtype=np.int64)
                idx = np.isin(keys, hclusters.pixel_keys(sig_px[:, 0],
                                                         sig_px[:, 1]))
                # np.argmax: the first condition wins a tie
                classes = np.argmax(mu_hat_alt[idx, :], axis=1) \
                    if idx.any() else np.zeros(0, dtype=np.int64)
                for ci, cond in enumerate(self.design.columns):
                    pick = classes == ci
                    found = ops.find_clusters(row[idx][pick], col[idx][pick])
                    self._save_cluster_files(
                        found, '%s/%s_%g_%i_%s.json'
                        % (self.outdir, cond, f, s, chrom), chrom)

    # -------------------------------------------------------------- simulate
    def simulate(self, cond, chrom=None, beta=0.5, p_diff=0.4, skip_bias=False,
                 loop_pattern=None, outdir='sim', n_threads=-1, verbose=True):
        """analysis/simulation.py:22-144: simulates raw contact matrices from
        the fitted scaled means and the dispersion trend of condition ``cond``,
        perturbing the loops of ``loop_pattern``; writes
        ``<outdir>/<rep>_<chrom>_raw.npz``, ``labels_<chrom>.txt`` and
        ``design.csv``.  The sampling runs on the GPU (csrc/simulate.cu)."""
        from hic3defdr_b200 import simulation as hsim
        if chrom is None:
            for c in self._cluster_chroms():
                self.simulate(cond, chrom=c, beta=beta, p_diff=p_diff,
                              skip_bias=skip_bias, loop_pattern=loop_pattern,
                              outdir=outdir, verbose=False)
            hdist.barrier()
            return
        eprint('simulating data for chrom %s' % chrom)
        if loop_pattern is None:
            loop_pattern = self.loop_patterns[cond]
        sel = np.asarray(self.design[cond].values).astype(bool)
        bias = self.load_bias(chrom)[:, sel]
        size_factors = self.load_data('size_factors', chrom)
        size_factors = size_factors[:, sel] if size_factors.ndim == 2 \
            else size_factors[sel]
        row = self.load_data('row', chrom)
        col = self.load_data('col', chrom)
        scaled = self.load_data('scaled', chrom)[:, sel]
        disp_fn = self.load_disp_fn(cond)
        clusters = load_clusters(loop_pattern.replace('<chrom>', chrom))
        mean = np.mean(scaled, axis=1)
        os.makedirs(outdir, exist_ok=True)
        n_sim_per_cond = size_factors.shape[-1]
        repnames = sum((['%s%i' % (c, i + 1) for i in range(n_sim_per_cond)]
                        for c in ['A', 'B']), [])
        design_file = '%s/design.csv' % outdir
        if not os.path.isfile(design_file):
            pd.DataFrame(
                {'A': [1] * n_sim_per_cond + [0] * n_sim_per_cond,
                 'B': [0] * n_sim_per_cond + [1] * n_sim_per_cond},
                dtype=bool, index=repnames).to_csv(design_file)
        if size_factors.ndim == 2:
            # one row per distance (the first pixel at that distance)
            eprint('  converting size factors', skip=not verbose)
            dist = col - row
            n_dists = int(dist.max()) + 1
            first = np.full(n_dists, 0, dtype=np.int64)
            seen = np.zeros(n_dists, dtype=bool)
            idx = np.arange(len(dist))[::-1]
            first[dist[::-1]] = idx
            seen[dist] = True
            table = size_factors[first, :]
            table[~seen] = size_factors[0, :]    # np.argmax of an all-False mask
            size_factors = table
        if skip_bias:
            bias = np.ones_like(bias)
            size_factors = np.ones_like(size_factors)
        bias = np.tile(bias, 2)
        size_factors = np.tile(size_factors, 2)
        classes, sim_iter = hsim.simulate(
            row, col, mean, disp_fn, bias, size_factors, clusters, beta=beta,
            p_diff=p_diff, trend='dist', verbose=verbose)
        np.savetxt('%s/labels_%s.txt' % (outdir, chrom), classes, fmt='%s')
        for rep, csr in zip(repnames, sim_iter):
            sparse.save_npz('%s/%s_%s_raw.npz' % (outdir, rep, chrom), csr)

    # -------------------------------------------------------------- evaluate
    def evaluate(self, cluster_pattern, label_pattern, min_dist=None,
                 max_dist=None, rerun_bh=False, outfile=None):
        """analysis/simulation.py:146-239: ROC / FDR-control curves of the
        q-values against the ground-truth labels of a simulation; writes
        ``<outdir>/eval.npz`` (or ``eval_<min_dist>_<max_dist>.npz``) with
        ``fdr``, ``fpr``, ``tpr``, ``thresh``.  Runs on rank 0."""
        from hic3defdr_b200 import evaluation as hev
        if outfile is None:
            outfile = 'eval.npz' if min_dist is None and max_dist is None \
                else 'eval_%s_%s.npz' % (min_dist, max_dist)
        if self.loop_patterns and cluster_pattern in self.loop_patterns.keys():
            cluster_pattern = self.loop_patterns[cluster_pattern]
        if hdist.rank() == 0:
            y_true, pvalues, qvalues = [], [], []
            for chrom in self.chroms:
                disp_idx = self.load_data('disp_idx', chrom)
                loop_idx = self.load_data('loop_idx', chrom)
                row = self.load_data('row', chrom, idx=(disp_idx, loop_idx))
                col = self.load_data('col', chrom, idx=(disp_idx, loop_idx))
                clusters = load_clusters(
                    cluster_pattern.replace('<chrom>', chrom))
                labels = np.loadtxt(label_pattern.replace('<chrom>', chrom),
                                    dtype='U7')
                dist = col - row
                dist_idx = np.ones(len(dist), dtype=bool)
                if min_dist is not None:
                    dist_idx[dist < min_dist] = False
                if max_dist is not None:
                    dist_idx[dist > max_dist] = False
                y_true.append(hev.make_y_true(row[dist_idx], col[dist_idx],
                                              clusters, labels))
                if min_dist is not None or max_dist is not None:
                    if rerun_bh:
                        pvalues.append(self.load_data(
                            'pvalues', chrom, idx=(loop_idx, dist_idx)))
                    else:
                        qvalues.append(self.load_data('qvalues', chrom,
                                                      idx=dist_idx))
            y_true = np.concatenate(y_true)
            if pvalues:
                with hdist.single_process():
                    qvalues = ops.adjust_pvalues(
                        np.concatenate(pvalues)).cpu().numpy()
            elif qvalues:
                qvalues = np.concatenate(qvalues)
            else:
                qvalues, _ = self.load_data('qvalues', 'all')
            fdr, fpr, tpr, thresh = hev.evaluate(y_true, qvalues)
            np.savez('%s/%s' % (self.outdir, outfile),
                     **{'fdr': fdr, 'fpr': fpr, 'tpr': tpr, 'thresh': thresh})
        hdist.barrier()

    def collect(self, fdr=0.05, cluster_size=3, n_threads=-1):
        """analysis/analysis.py:498-574: one ``results_<fdr>_<size>.tsv`` with
        the constitutive (insig) and per-condition clusters of every
        chromosome."""
        if self.res is None:
            raise ValueError(
                'the collect() step can only be run if the res kwarg was '
                'passed during construction of the HiC3DeFDR object; please '
                'run the classify() step instead or re-create the HiC3DeFDR '
                'object (you do not need to re-run any other steps)')
        eprint('collecting differential interactions')
        for f in self._as_list(fdr):
            for s in self._as_list(cluster_size):
                pattern = '%s/<class>_%g_%i_<chrom>.tsv' % (self.outdir, f, s)
                path = lambda cls, chrom: pattern.replace('<class>', cls) \
                    .replace('<chrom>', chrom)
                # threshold() / classify() contain barriers: every rank must
                # take the same branch, so rank 0 looks at the files and the
                # others follow its decision
                need = hdist.broadcast_flags([
                    not all(os.path.isfile(path('insig', c))
                            for c in self.chroms),
                    not all(os.path.isfile(path(cond, c))
                            for cond in self.design.columns
                            for c in self.chroms)])
                if need[0]:
                    self.threshold(fdr=f, cluster_size=s)
                if need[1]:
                    self.classify(fdr=f, cluster_size=s)
                if hdist.rank() != 0:
                    continue
                tables = []
                for chrom in self.chroms:
                    df = hclusters.load_cluster_table(path('insig', chrom))
                    df['classification'] = 'constitutive'
                    tables.append(df)
                    for cond in self.design.columns:
                        df = hclusters.load_cluster_table(path(cond, chrom))
                        df['classification'] = cond
                        tables.append(df)
                hclusters.sort_cluster_table(pd.concat(tables)).to_csv(
                    '%s/results_%g_%i.tsv' % (self.outdir, f, s), sep='\t')
        hdist.barrier()
