"""
Host-side mirrors of the reference's ``hic3defdr.util`` functions on the
run_to_qvalues path, each a thin call into libh3d (include/h3d.h).  Inputs may
be numpy arrays (copied to the current CUDA device) or CUDA torch tensors;
outputs are CUDA torch tensors unless stated.  No CPU fallback.
"""
import ctypes
import warnings

import numpy as np
import torch

from hic3defdr_b200._native import H3DError, lib, ptr

NORMS = {'conditional_mor': 0, 'conditional_scaling': 1,
         'median_of_ratios': 2, 'simple_scaling': 3}
ESTIMATORS = {'qcml': 0, 'cml': 1, 'mme': 2}
_DTYPES = {np.dtype(np.int64): 0, np.dtype(np.float64): 1,
           np.dtype(np.int32): 2, np.dtype(np.float32): 3}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def dev(a, dtype=None):
    """numpy / torch -> contiguous CUDA tensor of ``dtype``."""
    if not torch.cuda.is_available():
        raise H3DError('no CUDA device: hic3defdr_b200 has no CPU fallback')
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.ascontiguousarray(a)
        if a.dtype == np.bool_:
            a = a.view(np.uint8)
        if a.flags.writeable:
            t = torch.from_numpy(a)
        else:
            # views of inflated file buffers (hostio.load_npz) are read-only;
            # the tensor is only ever the source of the upload below
            with warnings.catch_warnings():
                warnings.simplefilter('ignore', UserWarning)
                t = torch.from_numpy(a)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.cuda().contiguous()


def workspace(nbytes):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device='cuda')


def design_bytes(design):
    return np.ascontiguousarray(np.asarray(design).astype(bool)).view(np.uint8)


_MAILBOX = {}


def to_host(t):
    """Small CUDA tensor (<= 64 KiB) -> numpy copy without the copy engine
    (h3d_publish into a pinned mailbox + a stream synchronisation): control
    values such as pixel counts must not wait behind bulk output copies."""
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes == 0:
        return np.empty(tuple(t.shape), dtype=torch.empty(
            0, dtype=t.dtype).numpy().dtype)
    padded = (nbytes + 7) // 8 * 8
    if padded > 65536 or padded != nbytes:
        return t.cpu().numpy()
    import threading
    dev_i = (torch.cuda.current_device(), threading.get_ident())
    box = _MAILBOX.get(dev_i)
    if box is None:
        box = torch.zeros(65536, dtype=torch.uint8).pin_memory()
        _MAILBOX[dev_i] = box
    lib().call('h3d_publish', ptr(t), ptr(box), nbytes, _stream())
    torch.cuda.current_stream().synchronize()
    return box[:nbytes].view(t.dtype).numpy().reshape(tuple(t.shape)).copy()


class Readback(object):
    """Deferred ``to_host``: the publish kernel is enqueued on the current
    stream now, ``get()`` synchronises THAT stream later and returns the numpy
    copy.  Lets several chromosomes (one stream each) keep their kernels
    queued while the host waits for one of them."""
    _box = {}
    _next = {}
    SLOT = 64          # bytes per slot
    N_SLOTS = 1024

    def __init__(self, t):
        import threading
        t = t.contiguous()
        self.dtype, self.shape = t.dtype, tuple(t.shape)
        self.nbytes = t.numel() * t.element_size()
        assert self.nbytes % 8 == 0 and 0 < self.nbytes <= self.SLOT
        key = (torch.cuda.current_device(), threading.get_ident())
        if key not in self._box:
            self._box[key] = torch.zeros(self.SLOT * self.N_SLOTS,
                                         dtype=torch.uint8).pin_memory()
            self._next[key] = 0
        slot = self._next[key]
        self._next[key] = (slot + 1) % self.N_SLOTS
        self.view = self._box[key][slot * self.SLOT:
                                   slot * self.SLOT + self.nbytes]
        self.stream = torch.cuda.current_stream()
        self.keep = t
        lib().call('h3d_publish', ptr(t), ptr(self.view), self.nbytes,
                   self.stream.cuda_stream)

    def get(self):
        self.stream.synchronize()
        self.keep = None
        return self.view.view(self.dtype).numpy().reshape(self.shape).copy()


def _check_failed(counter, what):
    n = int(to_host(counter.to(torch.int64))[0])
    if n:
        raise AssertionError(
            '%s: %d pixel(s) have no positive root (all-zero counts within a '
            'condition); the reference raises here too '
            '(hic3defdr/util/scaled_nb.py:139-183)' % (what, n))


# --------------------------------------------------------------------------
# prepare_data
# --------------------------------------------------------------------------

def fp64_peak_tflops():
    """Measured FP64 FMA rate of the current device (TFLOP/s, FMA = 2)."""
    scratch = torch.zeros(8, dtype=torch.float64, device='cuda')
    out = np.zeros(1)
    lib().call('h3d_fp64_peak', ptr(scratch), ptr(out), _stream())
    return float(out[0])


def filter_bias(bias, bias_thresh):
    """hic3defdr/analysis/core.py:58-59 on an (n_bins, n_reps) matrix."""
    b = dev(bias, torch.float64).clone()
    lib().call('h3d_bias_filter', ptr(b), b.shape[0], b.shape[1],
               float(bias_thresh), _stream())
    return b


class DeviceCSR(object):
    """The replicate matrices of one chromosome as device CSR arrays."""

    def __init__(self, mats):
        import scipy.sparse as sparse
        self.n_reps = len(mats)
        self.indptr, self.indices, self.data = [], [], []
        self.n_bins = mats[0].shape[0]
        dtypes, is64 = set(), set()
        pinned = []
        for m in mats:
            if not (sparse.issparse(m) and m.format == 'csr'):
                m = sparse.csr_matrix(m)
            # (data, indices, indptr) in page-locked memory, hostio.pin_csr
            pin = getattr(m, '_h3d_pinned', None)
            if not m.has_canonical_format:
                m = m.copy()
                m.sum_duplicates()
                pin = None
            dt = m.data.dtype
            if dt not in _DTYPES:
                dt = np.dtype(np.float64) if dt.kind == 'f' else \
                    np.dtype(np.int64)
            dtypes.add(dt)
            is64.add(m.indptr.dtype == np.int64)
            self.indptr.append(m.indptr)
            self.indices.append(m.indices.astype(np.int32, copy=False))
            self.data.append(m.data)
            pinned.append(pin or (None, None, None))
        self.dtype = dtypes.pop() if len(dtypes) == 1 else np.dtype(np.float64)
        self.is64 = int(any(is64))
        ip_t = np.int64 if self.is64 else np.int32

        def upload(a, t, dtype):
            # the pinned tensor holds exactly this array: asynchronous copy on
            # the current stream (the host allocator keeps the block until the
            # copy is done); anything else: blocking copy of the array
            if t is not None and a.dtype == dtype and a.size == t.numel() \
                    and a.size and a.ctypes.data == t.data_ptr():
                return t.cuda(non_blocking=True)
            return dev(a.astype(dtype, copy=False))
        self.indptr = [upload(a, p[2], ip_t)
                       for a, p in zip(self.indptr, pinned)]
        self.indices = [upload(a, p[1], np.int32)
                        for a, p in zip(self.indices, pinned)]
        self.data = [upload(a, p[0], self.dtype)
                     for a, p in zip(self.data, pinned)]
        self.nnz = sum(int(d.numel()) for d in self.data)

    def pointer_arrays(self):
        mk = lambda ts: (ctypes.c_void_p * self.n_reps)(*[t.data_ptr()
                                                           for t in ts])
        return mk(self.indptr), mk(self.indices), mk(self.data)


def union_gather(csr, dist_thresh, bias=None):
    """sparse_union + the raw / balanced gathers
    (hic3defdr/util/matrices.py:92-129, analysis/analysis.py:92-101).

    Returns dict(row, col, dist, raw, balanced) of CUDA tensors."""
    n = csr.n_bins
    ip, ix, dt = csr.pointer_arrays()
    offs = torch.empty(n + 1, dtype=torch.int32, device='cuda')
    wsb = lib().query('h3d_union_ws_bytes', n)
    ws = workspace(wsb)
    bptr = ptr(bias) if bias is not None else None
    lib().call('h3d_union_count', csr.n_reps, ip, csr.is64, ix, dt,
               _DTYPES[csr.dtype], bptr, n, int(dist_thresh), ptr(offs),
               ptr(ws), wsb, _stream())
    n_px = int(Readback(offs[n:n + 1].to(torch.int64)).get()[0])
    return union_emit(csr, dist_thresh, bias, offs, n_px)


def union_count_async(csr, dist_thresh, bias=None):
    """First half of ``union_gather``: per-row pixel counts / offsets on the
    device and a ``Readback`` of the total; finish with ``union_emit``."""
    n = csr.n_bins
    ip, ix, dt = csr.pointer_arrays()
    offs = torch.empty(n + 1, dtype=torch.int32, device='cuda')
    wsb = lib().query('h3d_union_ws_bytes', n)
    ws = workspace(wsb)
    bptr = ptr(bias) if bias is not None else None
    lib().call('h3d_union_count', csr.n_reps, ip, csr.is64, ix, dt,
               _DTYPES[csr.dtype], bptr, n, int(dist_thresh), ptr(offs),
               ptr(ws), wsb, _stream())
    return offs, Readback(offs[n:n + 1].to(torch.int64))


def union_emit(csr, dist_thresh, bias, offs, n_px):
    """Second half of ``union_gather`` given the row offsets and the total."""
    n = csr.n_bins
    ip, ix, dt = csr.pointer_arrays()
    bptr = ptr(bias) if bias is not None else None
    row = torch.empty(n_px, dtype=torch.int32, device='cuda')
    col = torch.empty_like(row)
    dist = torch.empty_like(row)
    raw = torch.empty((n_px, csr.n_reps), dtype=torch.int64, device='cuda')
    bal = torch.empty((n_px, csr.n_reps), dtype=torch.float64, device='cuda')
    if n_px:
        lib().call('h3d_union_emit', csr.n_reps, ip, csr.is64, ix, dt,
                   _DTYPES[csr.dtype], bptr, n, int(dist_thresh), ptr(offs),
                   ptr(row), ptr(col), ptr(dist), ptr(raw), ptr(bal),
                   _stream())
    return dict(row=row, col=col, dist=dist, raw=raw, balanced=bal)


def sparse_union(mats, dist_thresh=1000, bias=None):
    """hic3defdr/util/matrices.py:92-129 -> (row, col) int32 CUDA tensors.
    ``mats``: scipy sparse matrices or npz file names."""
    import scipy.sparse as sparse
    mats = [sparse.load_npz(m) if isinstance(m, str) else m for m in mats]
    b = dev(bias, torch.float64) if bias is not None else None
    out = union_gather(DeviceCSR(mats), dist_thresh, b)
    return out['row'], out['col']


def size_factor_table(balanced, dist, dist_max, n_bins, norm):
    """Device size-factor table ((dist_max + 1, R) for the conditional norms,
    (R,) otherwise); see ``conditional_mor`` & co. for the reference shapes."""
    bal = dev(balanced, torch.float64)
    n_px, n_reps = bal.shape
    conditional = 'conditional' in norm
    dd = dev(dist, torch.int32) if conditional else None
    shape = (dist_max + 1, n_reps) if conditional else (n_reps,)
    table = torch.empty(shape, dtype=torch.float64, device='cuda')
    wsb = lib().query('h3d_size_factors_ws_bytes', n_px, n_reps, dist_max)
    ws = workspace(wsb)
    lib().call('h3d_size_factors', ptr(dd), ptr(bal), n_px, n_reps,
               int(dist_max), int(n_bins or 0), NORMS[norm], ptr(table),
               ptr(ws), wsb, _stream())
    return table


# the stages of ``size_factor_table`` one by one: hic3defdr_b200.dist puts the
# collectives between them when a chromosome is sharded by row range
def sf_num_groups(dist_max, n_bins, norm):
    return int(lib().query('h3d_sf_num_groups', int(dist_max),
                           int(n_bins or 0), NORMS[norm]))


def sf_group_bounds(n_total, dist_max, n_bins, norm, key_start):
    """First position (chromosome-wide distance order) of every equal-count
    bin (util/binning.py:4-25) -> (n_groups + 1,) int64 CUDA tensor."""
    n_groups = sf_num_groups(dist_max, n_bins, norm)
    ks = dev(key_start, torch.int64) if key_start is not None else None
    gstart = torch.empty(n_groups + 1, dtype=torch.int64, device='cuda')
    lib().call('h3d_sf_group_bounds', int(n_total), int(dist_max),
               int(n_bins or 0), NORMS[norm], ptr(ks), ptr(gstart), _stream())
    return gstart


def sf_values(balanced, rank, norm):
    """Per-pixel ratios to the geometric mean over replicates
    (util/scaling.py:41-47; the plain values for the scaling norms), written
    replicate-major at position ``rank`` -> (R, n) float64 CUDA tensor."""
    bal = dev(balanced, torch.float64)
    n_px, n_reps = bal.shape
    out = torch.empty((n_reps, n_px), dtype=torch.float64, device='cuda')
    lib().call('h3d_sf_values', ptr(bal), ptr(rank), n_px, n_reps, NORMS[norm],
               ptr(out), _stream())
    return out


def sf_group_reduce(values, gstart, norm):
    """Exact median (sum for the scaling norms) of every (group, replicate)
    slice of ``values`` (R, ld) -> ((n_groups, R) float64, (n_groups,) int64
    number of ratios per group)."""
    n_reps, ld = values.shape
    gs = dev(gstart, torch.int64)
    n_groups = gs.numel() - 1
    red = torch.empty((n_groups, n_reps), dtype=torch.float64, device='cuda')
    valid = torch.zeros(max(n_groups, 1), dtype=torch.int64, device='cuda')
    lib().call('h3d_sf_group_reduce', ptr(values), int(ld), ptr(gs), n_groups,
               n_reps, NORMS[norm], ptr(red), ptr(valid), _stream())
    return red, valid[:n_groups]


def sf_table(red, gstart, key_start, dist_max, n_bins, norm):
    """util/scaling.py:92-104 (bin means, interpolation over distance) from
    the per-group reductions -> the table of ``size_factor_table``."""
    red = dev(red, torch.float64)
    n_groups, n_reps = red.shape
    conditional = 'conditional' in norm
    gs = dev(gstart, torch.int64)
    ks = dev(key_start, torch.int64) if conditional else None
    shape = (dist_max + 1, n_reps) if conditional else (n_reps,)
    table = torch.empty(shape, dtype=torch.float64, device='cuda')
    wsb = lib().query('h3d_sf_table_ws_bytes', n_groups, n_reps)
    ws = workspace(wsb)
    lib().call('h3d_sf_table', ptr(red), ptr(gs), ptr(ks), n_groups, n_reps,
               int(dist_max), int(n_bins or 0), NORMS[norm], ptr(table),
               ptr(ws), wsb, _stream())
    return table


def _conditional(data, dist, n_bins, norm):
    d = np.asarray(dist.cpu() if isinstance(dist, torch.Tensor) else dist)
    dist_max = int(d.max())
    table = size_factor_table(data, d.astype(np.int32), dist_max, n_bins, norm)
    return table[dev(d.astype(np.int64))]


def conditional_mor(data, dist, n_bins=None):
    """hic3defdr/util/scaling.py:108-127 -> (N, R)."""
    return _conditional(data, dist, n_bins, 'conditional_mor')


def conditional_scaling(data, dist, n_bins=None):
    """hic3defdr/util/scaling.py:130-149 -> (N, R)."""
    return _conditional(data, dist, n_bins, 'conditional_scaling')


def median_of_ratios(data):
    """hic3defdr/util/scaling.py:27-47 -> (R,)."""
    return size_factor_table(data, None, 0, 0, 'median_of_ratios')


def simple_scaling(data):
    """hic3defdr/util/scaling.py:50-65 -> (R,)."""
    return size_factor_table(data, None, 0, 0, 'simple_scaling')


def stable_rank(keys, n_keys):
    """Position of every element in a stable sort by ``keys`` (the canonical
    form of ``data.argsort().argsort()`` in hic3defdr/util/binning.py:25).
    Returns (rank int32 CUDA tensor, key_start int64 CUDA tensor)."""
    k = dev(keys, torch.int32)
    n = k.numel()
    rank = torch.empty(n, dtype=torch.int32, device='cuda')
    start = torch.empty(n_keys + 1, dtype=torch.int64, device='cuda')
    wsb = lib().query('h3d_stable_rank_ws_bytes', n, n_keys)
    ws = workspace(wsb)
    lib().call('h3d_stable_rank', ptr(k), n, int(n_keys), ptr(rank),
               ptr(start), ptr(ws), wsb, _stream())
    return rank, start


def equal_bin(data, n_bins):
    """hic3defdr/util/binning.py:4-25 (stable tie-break): bin index per
    element, as a numpy array."""
    data = np.asarray(data)
    rank, _ = stable_rank(data.astype(np.int32), int(data.max()) + 1)
    idx = np.linspace(0, n_bins, data.size, endpoint=0, dtype=int)
    return idx[rank.cpu().numpy()]


def scale_filter(row, col, balanced, sf_table, design, dist_max, mean_thresh,
                 dist_min):
    """hic3defdr/analysis/analysis.py:109-115.  ``balanced`` is overwritten
    with ``scaled``.  Returns (scaled, size_factors (N, R) or the (R,) table,
    disp_idx uint8)."""
    n_px, n_reps = balanced.shape
    design = np.asarray(design).astype(bool)
    per_dist = sf_table.dim() == 2
    sf_out = torch.empty_like(balanced) if per_dist else None
    disp_idx = torch.empty(n_px, dtype=torch.uint8, device='cuda')
    db = design_bytes(design)     # keep alive across the call
    lib().call('h3d_scale_filter', ptr(row), ptr(col), ptr(balanced),
               ptr(sf_table), int(per_dist), ptr(db), n_px,
               n_reps, design.shape[1], int(dist_max), float(mean_thresh),
               int(dist_min), ptr(sf_out), ptr(disp_idx), _stream())
    return balanced, (sf_out if per_dist else sf_table), disp_idx


def mask_to_index(mask):
    """Positions of the True entries of a uint8/bool CUDA mask (int32)."""
    m = dev(mask, torch.uint8)
    n = m.numel()
    idx = torch.empty(n, dtype=torch.int32, device='cuda')
    cnt = torch.zeros(1, dtype=torch.int64, device='cuda')
    wsb = lib().query('h3d_mask_to_index_ws_bytes', n)
    ws = workspace(wsb)
    lib().call('h3d_mask_to_index', ptr(m), n, ptr(idx), ptr(cnt), ptr(ws),
               wsb, _stream())
    return idx[:int(to_host(cnt)[0])]


def mask_to_index_async(mask):
    """``mask_to_index`` without the host wait: (index buffer of full length,
    Readback of the number of set entries)."""
    m = dev(mask, torch.uint8)
    n = m.numel()
    idx = torch.empty(n, dtype=torch.int32, device='cuda')
    cnt = torch.zeros(1, dtype=torch.int64, device='cuda')
    wsb = lib().query('h3d_mask_to_index_ws_bytes', n)
    ws = workspace(wsb)
    lib().call('h3d_mask_to_index', ptr(m), n, ptr(idx), ptr(cnt), ptr(ws),
               wsb, _stream())
    return idx, Readback(cnt)


def loop_membership(row, col, index, pixels):
    """hic3defdr/analysis/analysis.py:117-125: which of (row, col)[index] are
    in the iterable of (i, j) ``pixels``."""
    keys = np.array(sorted({(int(i) << 32) | int(j) for i, j in pixels}),
                    dtype=np.int64)
    kd = dev(keys)
    n = index.numel() if index is not None else row.numel()
    out = torch.empty(n, dtype=torch.uint8, device='cuda')
    lib().call('h3d_loop_membership', ptr(row), ptr(col), ptr(index), n,
               ptr(kd), kd.numel(), ptr(out), _stream())
    return out


# --------------------------------------------------------------------------
# estimate_disp
# --------------------------------------------------------------------------

def gather_counts_factors(row, col, index, raw, size_factors, bias, dest, ld,
                          x_out, f_out, dist_out):
    """hic3defdr/analysis/analysis.py:181-183 (+ raw[disp_idx]) into SoA."""
    n_sel = index.numel() if index is not None else row.numel()
    n_reps = raw.shape[1] if raw is not None else 0
    per_px = int(size_factors is not None and size_factors.dim() == 2)
    lib().call('h3d_gather_counts_factors', ptr(row), ptr(col), ptr(index),
               n_sel, ptr(raw), ptr(size_factors), per_px, ptr(bias), n_reps,
               ptr(dest), int(ld), ptr(x_out), ptr(f_out), ptr(dist_out),
               _stream())


def estimate_dispersion(x_soa, f_soa, seg_start, design, estimator='qcml',
                        runs=None):
    """Per-(segment, condition) dispersion of pooled SoA data.
    x_soa, f_soa: (R, ld) CUDA; seg_start: host int64 (n_seg + 1) boundaries of
    contiguous segments, or ``runs`` = (run_seg int32, run_lo int64, run_hi
    int64) host arrays when a segment is a list of runs (include/h3d.h,
    h3d_estimate_dispersion_runs; segment ids 0 .. max(run_seg)).
    Returns (disp (n_seg, C) numpy, stats dict)."""
    design = np.asarray(design).astype(bool)
    n_reps, n_conds = design.shape
    ld = x_soa.shape[1]
    stats = np.zeros(9, dtype=np.int64)
    db = design_bytes(design)
    if runs is None:
        seg = np.ascontiguousarray(seg_start, dtype=np.int64)
        n_seg = len(seg) - 1
        out = np.empty((n_seg, n_conds))
        wsb = lib().query('h3d_estimate_dispersion_ws_bytes', int(seg[-1]),
                          n_seg, n_reps, n_conds)
        ws = workspace(wsb)
        lib().call('h3d_estimate_dispersion', ptr(x_soa), ptr(f_soa), ld,
                   ptr(seg), n_seg, ptr(db), n_reps, n_conds,
                   ESTIMATORS[estimator], ptr(out), ptr(stats), ptr(ws), wsb,
                   _stream())
    else:
        run_seg = np.ascontiguousarray(runs[0], dtype=np.int32)
        run_lo = np.ascontiguousarray(runs[1], dtype=np.int64)
        run_hi = np.ascontiguousarray(runs[2], dtype=np.int64)
        n_runs = len(run_seg)
        n_seg = int(seg_start) if np.isscalar(seg_start) else \
            (int(run_seg.max()) + 1 if n_runs else 1)
        n_px = int((run_hi - run_lo).sum())
        out = np.empty((n_seg, n_conds))
        wsb = lib().query('h3d_estimate_dispersion_runs_ws_bytes', n_px,
                          n_runs, n_seg, n_reps, n_conds)
        ws = workspace(wsb)
        lib().call('h3d_estimate_dispersion_runs', ptr(x_soa), ptr(f_soa), ld,
                   ptr(run_seg), ptr(run_lo), ptr(run_hi), n_runs, n_seg,
                   ptr(db), n_reps, n_conds, ESTIMATORS[estimator], ptr(out),
                   ptr(stats), ptr(ws), wsb, _stream())
    if stats[8]:
        import sys
        print('  warning: the qCML fixed point of %d (distance, condition) bins '
              'did not settle within its tolerance (tiny bins with a large '
              'dispersion; the reference loops forever there, '
              'hic3defdr/util/dispersion.py:36-42); their last iterate is used'
              % int(stats[8]), file=sys.stderr)
    return out, dict(outer_iterations=int(stats[0]),
                     nll_evaluations=int(stats[1]),
                     pixel_equalizations=int(stats[2]),
                     launches=int(stats[3]),
                     equalize_launches=int(stats[4]),
                     equalize_us=int(stats[5]),
                     nll_launches=int(stats[6]), nll_us=int(stats[7]),
                     capped_segments=int(stats[8]))


def equalize(data, f, alpha):
    """hic3defdr/util/scaled_nb.py:186-214: pseudo-data of one bin, (n, R)
    counts and factors, scalar dispersion -> (n, R) float64 CUDA tensor."""
    data = np.asarray(data.cpu() if isinstance(data, torch.Tensor) else data,
                      dtype=float)
    n, r = data.shape
    x = dev(np.ascontiguousarray(data.T))
    fd = dev(np.ascontiguousarray(np.asarray(f, dtype=float).T))
    out = torch.empty((r, max(n, 1)), dtype=torch.float64, device='cuda')
    failed = torch.zeros(1, dtype=torch.int32, device='cuda')
    wsb = lib().query('h3d_equalize_ws_bytes', n)
    ws = workspace(wsb)
    lib().call('h3d_equalize', ptr(x), ptr(fd), max(n, 1), n, r, float(alpha),
               ptr(out), ptr(failed), ptr(ws), wsb, _stream())
    _check_failed(failed, 'equalize')
    return out[:, :n].t().contiguous()


def cml_nll(data, delta):
    """The objective of ``cml`` (hic3defdr/util/dispersion.py:72-75) at
    ``delta`` for (n, R) data -> float."""
    data = np.asarray(data.cpu() if isinstance(data, torch.Tensor) else data,
                      dtype=float)
    n, r = data.shape
    x = dev(np.ascontiguousarray(data.T))
    out = torch.zeros(1, dtype=torch.float64, device='cuda')
    wsb = lib().query('h3d_cml_nll_ws_bytes', n)
    ws = workspace(wsb)
    lib().call('h3d_cml_nll', ptr(x), n, n, r, float(delta), ptr(out),
               ptr(ws), wsb, _stream())
    return float(out.item())


def _single_bin(data, f, estimator):
    data = np.asarray(data.cpu() if isinstance(data, torch.Tensor) else data,
                      dtype=float)
    n, r = data.shape
    f = np.ones_like(data) if f is None else np.asarray(f, dtype=float)
    x = dev(np.ascontiguousarray(data.T))
    fd = dev(np.ascontiguousarray(f.T))
    d, _ = estimate_dispersion(x, fd, [0, n], np.ones((r, 1), bool), estimator)
    return float(d[0, 0])


def qcml(data, f=None):
    """hic3defdr/util/dispersion.py:10-43."""
    return _single_bin(data, f, 'qcml')


def cml(data, f=None):
    """hic3defdr/util/dispersion.py:46-80 (float division by f)."""
    return _single_bin(data, f, 'cml')


def mme(data, f=None):
    """hic3defdr/util/dispersion.py:108-131 (float division by f)."""
    return _single_bin(data, f, 'mme')


def gather_table(dist, table):
    """disp[:, c] = table[dist, c] (analysis/analysis.py:218)."""
    t = dev(table, torch.float64)
    n = dist.numel()
    out = torch.empty((n, t.shape[1]), dtype=torch.float64, device='cuda')
    lib().call('h3d_gather_table', ptr(dist), n, ptr(t), t.shape[1],
               t.shape[0], ptr(out), _stream())
    return out


# --------------------------------------------------------------------------
# lrt / bh
# --------------------------------------------------------------------------

def fit_mu_hat(x, b, alpha):
    """hic3defdr/util/scaled_nb.py:71-183 (same broadcasting forms)."""
    x = np.asarray(x) if not isinstance(x, torch.Tensor) else x
    b = np.asarray(b) if not isinstance(b, torch.Tensor) else b
    if x.ndim == 1:
        x, b = x[None, :], b[None, :]
    n, r = x.shape
    xd, bd = dev(x, torch.float64), dev(b, torch.float64)
    al = alpha if isinstance(alpha, torch.Tensor) else np.asarray(
        alpha, dtype=float)
    if al.ndim == 0:
        spx, srep = 0, 0
        al = al.reshape(1)
    elif al.ndim == 1:
        assert al.shape[0] == r
        spx, srep = 0, 1
    elif al.shape[1] == 1:
        spx, srep = 1, 0
    else:
        spx, srep = r, 1
    ad = dev(al, torch.float64)
    out = torch.empty(n, dtype=torch.float64, device='cuda')
    failed = torch.zeros(1, dtype=torch.int32, device='cuda')
    lib().call('h3d_fit_mu_hat', ptr(xd), ptr(bd), ptr(ad), spx, srep, n, r,
               ptr(out), ptr(failed), _stream())
    _check_failed(failed, 'fit_mu_hat')
    return out


def lrt(raw, f, disp, design, refit_mu=True):
    """hic3defdr/util/lrt.py:7-50.  ``disp`` is the per-condition (n, C)
    dispersion (the reference widens it with ``disp @ design.T`` before the
    call; here the widening happens in the kernel)."""
    design = np.asarray(design).astype(bool)
    n_reps, n_conds = design.shape
    rd, fd, dd = dev(raw, torch.float64), dev(f, torch.float64), \
        dev(disp, torch.float64)
    n = rd.shape[0]
    p = torch.empty(n, dtype=torch.float64, device='cuda')
    llr = torch.empty_like(p)
    mu0 = torch.empty_like(p)
    mu1 = torch.empty((n, n_conds), dtype=torch.float64, device='cuda')
    failed = torch.zeros(1, dtype=torch.int32, device='cuda')
    db = design_bytes(design)
    lib().call('h3d_lrt', ptr(rd), ptr(fd), ptr(dd), ptr(db),
               n, n_reps, n_conds, int(bool(refit_mu)), ptr(p), ptr(llr),
               ptr(mu0), ptr(mu1), ptr(failed), _stream())
    _check_failed(failed, 'lrt')
    return p, llr, mu0, mu1


def check_failed(counter, what):
    """raises the reference's AssertionError if the device counter of pixels
    without a positive root is non-zero"""
    _check_failed(counter, what)


def lrt_fused(row, col, index, raw, size_factors, bias, disp, design,
              refit_mu=True, failed=None):
    """The LRT of hic3defdr/analysis/analysis.py:261-278 reading the
    union-aligned device arrays directly.  ``failed``: optional int32 device
    counter shared by several calls; the caller then runs ``check_failed``
    once (one host synchronisation instead of one per chromosome)."""
    design = np.asarray(design).astype(bool)
    n_reps, n_conds = design.shape
    n = index.numel()
    p = torch.empty(n, dtype=torch.float64, device='cuda')
    llr = torch.empty_like(p)
    mu0 = torch.empty_like(p)
    mu1 = torch.empty((n, n_conds), dtype=torch.float64, device='cuda')
    deferred = failed is not None
    if not deferred:
        failed = torch.zeros(1, dtype=torch.int32, device='cuda')
    db = design_bytes(design)
    lib().call('h3d_lrt_fused', ptr(row), ptr(col), ptr(index), n, ptr(raw),
               ptr(size_factors), int(size_factors.dim() == 2), ptr(bias),
               ptr(disp), ptr(db), n_reps, n_conds,
               int(bool(refit_mu)), ptr(p), ptr(llr), ptr(mu0), ptr(mu1),
               ptr(failed), _stream())
    if not deferred:
        _check_failed(failed, 'lrt')
    return p, llr, mu0, mu1


def adjust_pvalues(pvalues):
    """lib5c.util.statistics.adjust_pvalues (BH) as called at
    hic3defdr/analysis/analysis.py:300."""
    p = dev(pvalues, torch.float64)
    n = p.numel()
    q = torch.empty_like(p)
    if n == 0:
        return q
    wsb = lib().query('h3d_bh_ws_bytes', n)
    ws = workspace(wsb)
    lib().call('h3d_bh', ptr(p), n, ptr(q), ptr(ws), wsb, _stream())
    return q


def adjust_pvalues_ranked(pvalues, rank_offset, n_total):
    """One bucket of a multi-GPU BH correction (include/h3d.h,
    h3d_bh_ranked): returns (q before the carry of the higher buckets, the
    bucket's minimum raw ratio as a 1-element CUDA tensor)."""
    p = dev(pvalues, torch.float64)
    n = p.numel()
    q = torch.empty_like(p)
    mn = torch.full((1,), float('inf'), dtype=torch.float64, device='cuda')
    if n == 0:
        return q, mn
    wsb = lib().query('h3d_bh_ws_bytes', n)
    ws = workspace(wsb)
    lib().call('h3d_bh_ranked', ptr(p), n, int(rank_offset), int(n_total),
               ptr(q), ptr(mn), ptr(ws), wsb, _stream())
    return q, mn


def apply_bh_carry(q, carry):
    """q <- min(q, carry) in place, NaN kept; ``carry`` a float or a 0-dim
    CUDA tensor (h3d_bh_apply_carry / h3d_bh_apply_carry_dev: the device form
    needs no host round trip)."""
    if isinstance(carry, torch.Tensor):
        c = carry.to(torch.float64).contiguous().view(1)
        lib().call('h3d_bh_apply_carry_dev', ptr(q), q.numel(), ptr(c),
                   _stream())
        return q
    lib().call('h3d_bh_apply_carry', ptr(q), q.numel(), float(carry),
               _stream())
    return q


# --------------------------------------------------------------------------
# threshold / classify
# --------------------------------------------------------------------------

def connected_components(row, col):
    """4-connected components of the pixels (row, col), which must be sorted
    by (row, col) and unique (hic3defdr/util/clusters.py:69-96, find_clusters
    with connectivity 1).  Returns (label, size) int32 CUDA tensors: label[i]
    is the position of the first pixel of pixel i's component, size[i] the
    component's pixel count at its first pixel and 0 elsewhere."""
    r = dev(row, torch.int32)
    c = dev(col, torch.int32)
    n = r.numel()
    label = torch.empty(n, dtype=torch.int32, device='cuda')
    size = torch.empty(n, dtype=torch.int32, device='cuda')
    if n == 0:
        return label, size
    wsb = lib().query('h3d_connected_components_ws_bytes', n)
    ws = workspace(wsb)
    lib().call('h3d_connected_components', ptr(r), ptr(c), n, ptr(label),
               ptr(size), ptr(ws), wsb, _stream())
    return label, size


def clusters_from_labels(row, col, label, size, min_size=1):
    """The components of ``connected_components`` with at least ``min_size``
    pixels (util/thresholding.py:47-61 size_filter) as a list of (k, 2) int
    arrays of [row, col] pairs, ordered by their first pixel."""
    if row.numel() == 0:
        return []
    keep = size[label.long()] >= int(min_size)
    lab = label[keep].cpu().numpy()
    px = torch.stack([row[keep], col[keep]], dim=1).cpu().numpy()
    if len(lab) == 0:
        return []
    order = np.argsort(lab, kind='stable')
    lab, px = lab[order], px[order]
    return np.split(px, np.flatnonzero(np.diff(lab)) + 1)


def find_clusters(row, col, min_size=1):
    """hic3defdr/util/clusters.py:69-96 (connectivity 1) on pixels sorted by
    (row, col): list of (k, 2) int arrays of [row, col] pairs."""
    r = dev(row, torch.int32)
    c = dev(col, torch.int32)
    label, size = connected_components(r, c)
    return clusters_from_labels(r, c, label, size, min_size)
