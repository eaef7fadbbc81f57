"""
Host <-> device staging for the end-to-end path: inputs travel host -> HBM on
a copy stream one chromosome ahead of the kernels that consume them, and every
output array starts its way back to pinned host memory on a second copy stream
the moment it is final, so the PCIe transfers (2.3 GB in / 7.6 GB out for the
mouse genome at 10 kb) overlap the FP64-bound dispersion and LRT kernels
instead of bracketing them.  The uploads have priority on the link: the output
drain starts once the last chromosome is uploaded (``OutputDrain(after=...)``;
measured on B200: both directions at once run at 47 GB/s each, 55 GB/s alone,
and only the uploads are on the critical path -- 237.7 -> 224.0 ms per step).  The reference has no counterpart: every step
round-trips through ``.npy`` files (hic3defdr/analysis/core.py:62-218).

Lifetime rule: a tensor that a copy stream reads or writes stays referenced
(by the prefetcher / the chromosome state / the drain) until ``wait()``.
"""
import numpy as np
import torch

from hic3defdr_b200 import ops


def pinned_csr(mats):
    """scipy CSR matrices of one chromosome -> dict lists of pinned host
    tensors (indptr, indices, data), canonical format, common dtype."""
    import scipy.sparse as sparse
    out = []
    dtypes = set()
    for m in mats:
        m = sparse.csr_matrix(m)
        if not m.has_canonical_format:
            m = m.copy()
            m.sum_duplicates()
        dt = m.data.dtype
        if dt not in ops._DTYPES:
            dt = np.dtype(np.float64) if dt.kind == 'f' else np.dtype(np.int64)
        dtypes.add(dt)
        out.append(m)
    dtype = dtypes.pop() if len(dtypes) == 1 else np.dtype(np.float64)
    is64 = any(m.indptr.dtype == np.int64 for m in out)
    ip_t = np.int64 if is64 else np.int32
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return [dict(indptr=pin(m.indptr.astype(ip_t, copy=False)),
                 indices=pin(m.indices.astype(np.int32, copy=False)),
                 data=pin(m.data.astype(dtype, copy=False))) for m in out]


def row_weights(mats):
    """stored entries per row, summed over the replicate matrices: the weight
    ``dist.row_ranges`` balances a row-sharded chromosome by"""
    import scipy.sparse as sparse
    return np.sum([np.diff(sparse.csr_matrix(m).indptr) for m in mats], axis=0)


def shard_rows(mats, lo, hi):
    """The rows [lo, hi) of every replicate matrix as CSR matrices of the
    ORIGINAL shape (the other rows empty): row and column numbers keep their
    chromosome-wide meaning, only the stored entries of the range are kept
    (and later uploaded)."""
    import scipy.sparse as sparse
    out = []
    for m in mats:
        m = sparse.csr_matrix(m)
        if not m.has_canonical_format:
            m = m.copy()
            m.sum_duplicates()
        a, b = int(m.indptr[lo]), int(m.indptr[hi])
        indptr = np.clip(m.indptr, a, b) - a
        out.append(sparse.csr_matrix(
            (m.data[a:b], m.indices[a:b], indptr.astype(m.indptr.dtype)),
            shape=m.shape))
    return out


def csr_to_device(host_mats, n_bins):
    """pinned (or device) CSR pieces -> ops.DeviceCSR on the current stream"""
    c = ops.DeviceCSR.__new__(ops.DeviceCSR)
    c.n_reps = len(host_mats)
    c.n_bins = n_bins
    c.indptr = [m['indptr'].cuda(non_blocking=True) for m in host_mats]
    c.indices = [m['indices'].cuda(non_blocking=True) for m in host_mats]
    c.data = [m['data'].cuda(non_blocking=True) for m in host_mats]
    c.dtype = np.dtype({torch.int64: np.int64, torch.float64: np.float64,
                        torch.int32: np.int32, torch.float32: np.float32}
                       [host_mats[0]['data'].dtype])
    c.is64 = int(host_mats[0]['indptr'].dtype == torch.int64)
    c.nnz = sum(int(d.numel()) for d in c.data)
    return c


class InputPrefetcher(object):
    """Iterates (DeviceCSR, bias) over host chromosomes; the copies of
    chromosome i + depth are enqueued on a copy stream before chromosome i is
    handed to the consumer, which only waits on that chromosome's event.

    ``host_chroms``: list of (list of pinned CSR dicts, pinned bias (n, R))."""

    def __init__(self, host_chroms, depth=2):
        self.host = host_chroms
        self.depth = depth
        self.stream = torch.cuda.Stream()
        self.keep = []
        # recorded on the copy stream behind the last chromosome's upload
        self.uploaded = torch.cuda.Event()
        self.n_issued = 0

    def _issue(self, i):
        mats, bias = self.host[i]
        cur = torch.cuda.current_stream()
        # device buffers come from the consumer stream's pool (no cross-stream
        # allocator traffic); the copy stream may fill them once the consumer
        # stream has passed this point (their previous users are done by then)
        dev = [{k: torch.empty(v.shape, dtype=v.dtype, device='cuda')
                for k, v in m.items()} for m in mats]
        b = torch.empty(bias.shape, dtype=bias.dtype, device='cuda')
        here = torch.cuda.Event()
        here.record(cur)
        self.stream.wait_event(here)
        with torch.cuda.stream(self.stream):
            for m, d in zip(mats, dev):
                for k, v in m.items():
                    d[k].copy_(v, non_blocking=True)
            b.copy_(bias, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
            self.n_issued += 1
            if self.n_issued == len(self.host):
                self.uploaded.record(self.stream)
        return csr_to_device(dev, bias.shape[0]), b, ev

    def __iter__(self):
        pending = []
        nxt = 0
        n = len(self.host)
        while nxt < min(self.depth, n):
            pending.append(self._issue(nxt))
            nxt += 1
        for i in range(n):
            csr, b, ev = pending.pop(0)
            if nxt < n:
                pending.append(self._issue(nxt))
                nxt += 1
            torch.cuda.current_stream().wait_event(ev)
            self.keep.append((csr, b))
            yield csr, b

    def __len__(self):
        return len(self.host)


_UNPACK_POOL = None


def _unpack_pool():
    global _UNPACK_POOL
    if _UNPACK_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _UNPACK_POOL = ThreadPoolExecutor(2)
    return _UNPACK_POOL


class OutputDrain(object):
    """``drain(i, name, tensor)``: device -> pinned host copy of a final
    output on a copy stream, ordered after the kernels that produced it.
    ``wait()`` returns {(i, name): pinned tensor} once everything landed.

    ``compact`` (default): three arrays do not cross the link as they are
    stored (csrc/hostpack.cu): ``size_factors`` and ``disp`` travel as their
    per-distance tables and are rebuilt in the pinned output buffer by host
    threads from ``row`` / ``col`` (/ ``disp_idx``), ``raw`` travels as int32
    and is widened there.  Same arrays for the caller, 2.9 GB less on the wire
    for the mouse genome; the rebuilding overlaps the dispersion estimate."""

    def __init__(self, pool=None, after=None, compact=True, host_threads=None):
        """``after``: an ``InputPrefetcher`` whose uploads go first.  The two
        directions of the link share ~92 GB/s on this pool's hosts (55 GB/s
        each alone), and the uploads are on the critical path (the kernels wait
        for them) while the outputs of ``prepare_data`` can drain during the
        ~120 ms of dispersion estimation that follow; so the drain holds its
        copies back until every chromosome has been queued for upload and the
        last upload has finished."""
        import os
        self.stream = torch.cuda.Stream()
        self.pool = pool if pool is not None else {}
        self.out = {}
        self.keep = []
        self.nbytes = 0
        self.after = after
        self.held = []
        self.compact = compact
        self.jobs = []
        self.host_threads = host_threads or max(
            1, min(16, (os.cpu_count() or 1) // max(1, _world())))

    # ---- entry points ---------------------------------------------------
    def __call__(self, i, name, tensor):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._submit(lambda: self._copy(i, name, tensor, ev))

    def table(self, i, name, table, n_rows, mask=None):
        """the (n_rows, C) array ``table[col - row]`` of chromosome i (over the
        pixels with ``mask`` set, the name of an output already handed in):
        only the table crosses the link"""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._submit(lambda: self._table(i, name, table, n_rows, mask, ev))

    def narrow(self, i, name, tensor):
        """an int64 array whose values fit 32 bits: copied as int32, widened
        on the host (copied as it is if a value does not fit)"""
        from hic3defdr_b200._native import lib, ptr
        t32 = torch.empty(tensor.shape, dtype=torch.int32, device='cuda')
        flag = torch.zeros(1, dtype=torch.int32, device='cuda')
        lib().call('h3d_narrow_i64', ptr(tensor), tensor.numel(), ptr(t32),
                   ptr(flag), torch.cuda.current_stream().cuda_stream)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._submit(lambda: self._narrow(i, name, tensor, t32, flag, ev))

    def _submit(self, fn):
        if self.after is not None:
            self.held.append(fn)
            if self.after.n_issued < len(self.after.host):
                return
            self.stream.wait_event(self.after.uploaded)
            held, self.held, self.after = self.held, [], None
            for f in held:
                f()
            return
        fn()

    # ---- copies ---------------------------------------------------------
    def _host(self, key, shape, dtype):
        host = self.pool.get(key)
        if host is None or tuple(host.shape) != tuple(shape) or \
                host.dtype != dtype:
            host = torch.empty(tuple(shape), dtype=dtype).pin_memory()
            self.pool[key] = host
        return host

    def _copy(self, i, name, tensor, ev, publish=True):
        key = (i, name)
        host = self._host(key, tensor.shape, tensor.dtype)
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            host.copy_(tensor, non_blocking=True)
        self.keep.append(tensor)
        if publish:
            self.out[key] = host
        self.nbytes += tensor.numel() * tensor.element_size()
        return host

    def _landed(self):
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return ev

    def _table(self, i, name, table, n_rows, mask, ev):
        from hic3defdr_b200._native import lib, ptr
        th = self._copy(i, name + '__table', table, ev, publish=False)
        out = self._host((i, name), (n_rows, table.shape[1]), torch.float64)
        self.out[(i, name)] = out
        row, col = self.out[(i, 'row')], self.out[(i, 'col')]
        m = self.out[(i, mask)] if mask else None
        landed = self._landed()          # row, col, mask and the table are in
        n_threads = self.host_threads

        def job():
            landed.synchronize()
            lib().call('h3d_host_expand_by_distance', ptr(th), th.shape[0],
                       th.shape[1], ptr(row), ptr(col), ptr(m), row.numel(),
                       ptr(out), n_threads)
        self.jobs.append(_unpack_pool().submit(job))

    def _narrow(self, i, name, tensor, t32, flag, ev):
        from hic3defdr_b200._native import lib, ptr
        h32 = self._copy(i, name + '__i32', t32, ev, publish=False)
        hflag = self._copy(i, name + '__overflow', flag, ev, publish=False)
        out = self._host((i, name), tensor.shape, torch.int64)
        self.out[(i, name)] = out
        landed = self._landed()
        n_threads = self.host_threads

        def job():
            landed.synchronize()
            if int(hflag[0]):
                return (i, name, tensor)         # redo as a plain copy
            lib().call('h3d_host_widen_i32', ptr(h32), ptr(out), h32.numel(),
                       n_threads)
        self.keep.append(tensor)
        self.jobs.append(_unpack_pool().submit(job))

    def wait(self):
        if self.held:                  # uploads never completed the hand-over
            held, self.held, self.after = self.held, [], None
            for f in held:
                f()
        self.stream.synchronize()
        redo = [r for r in (j.result() for j in self.jobs) if r is not None]
        self.jobs = []
        for i, name, tensor in redo:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._copy(i, name, tensor, ev)
        if redo:
            self.stream.synchronize()
        self.keep = []
        return self.out


def _world():
    import torch.distributed as td
    return td.get_world_size() if td.is_available() and td.is_initialized() \
        else 1
