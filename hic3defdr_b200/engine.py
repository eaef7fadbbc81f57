"""
Device-resident stages of ``run_to_qvalues``: the orchestration shared by the
drop-in class (hic3defdr_b200/analysis.py, which adds the file I/O) and by
bench.py (which times it with inputs already in HBM).  Each function takes
and returns CUDA tensors; the arithmetic is in libh3d (include/h3d.h).

A chromosome's state is a dict of tensors named after the reference's files
(hic3defdr/analysis/analysis.py:125-133, 219-223, 281-284, 302-303):
row, col, raw, size_factors, scaled, disp_idx (+ disp_index, the positions of
the True entries), bias, [loop_idx], disp, pvalues, llr, mu_hat_null,
mu_hat_alt, qvalues.
"""
import os

import numpy as np
import torch

from hic3defdr_b200 import dist as hdist
from hic3defdr_b200 import ops
from hic3defdr_b200.trace import stage
from hic3defdr_b200.trend import fit_many


def prepare_chrom_steps(csr, bias_raw, design, dist_min=4, dist_max=200,
                        bias_thresh=0.1, mean_thresh=1.0,
                        norm='conditional_mor', n_bins=-1, loop_pixels=None):
    """hic3defdr/analysis/analysis.py:63-133 for one chromosome as a generator:
    it yields an ``ops.Readback`` wherever the host needs a count from the
    device (number of union pixels, number of tested pixels) and is resumed
    with the value, so that a driver can keep several chromosomes in flight on
    separate streams (``prepare_many``).  Returns the chromosome state."""
    if n_bins == -1:
        n_bins = int(dist_max / 5)
    bias = ops.filter_bias(bias_raw, bias_thresh)
    offs, pending = ops.union_count_async(csr, dist_max, bias)
    n_px = int((yield pending)[0])
    u = ops.union_emit(csr, dist_max, bias, offs, n_px)
    st = dict(bias=bias, row=u['row'], col=u['col'], raw=u['raw'])
    if n_px == 0:
        return _empty_prepared(st, csr.n_reps, u['balanced'], loop_pixels)
    table = ops.size_factor_table(u['balanced'], u['dist'], dist_max, n_bins,
                                  norm)
    scaled, sf, disp_idx = ops.scale_filter(
        u['row'], u['col'], u['balanced'], table, design, dist_max,
        mean_thresh, dist_min)
    index, pending = ops.mask_to_index_async(disp_idx)
    n_d = int((yield pending)[0])
    st.update(size_factors=sf, scaled=scaled, disp_idx=disp_idx,
              disp_index=index[:n_d])
    if table.dim() == 2:
        st['sf_table'] = table          # (D + 1, R): size_factors = table[dist]
    if loop_pixels is not None:
        st['loop_idx'] = ops.loop_membership(st['row'], st['col'],
                                             st['disp_index'], loop_pixels)
    return st


def prepare_chrom(csr, bias_raw, design, *args, **kw):
    """One chromosome, synchronously (same arguments as
    ``prepare_chrom_steps``)."""
    gen = prepare_chrom_steps(csr, bias_raw, design, *args, **kw)
    try:
        pending = next(gen)
        while True:
            pending = gen.send(pending.get())
    except StopIteration as stop:
        return stop.value


def _empty_prepared(st, n_reps, balanced, loop_pixels=None):
    st.update(size_factors=torch.empty((0, n_reps), dtype=torch.float64,
                                       device='cuda'),
              scaled=balanced,
              disp_idx=torch.empty(0, dtype=torch.uint8, device='cuda'),
              disp_index=torch.empty(0, dtype=torch.int32, device='cuda'))
    if loop_pixels is not None:
        # every rank must hold the same set of arrays: the sharded writers
        # (analysis._save_sharded) run one collective per array
        st['loop_idx'] = torch.empty(0, dtype=torch.uint8, device='cuda')
    return st


def prepare_chrom_sharded(csr, bias_raw, design, dist_min=4, dist_max=200,
                          bias_thresh=0.1, mean_thresh=1.0,
                          norm='conditional_mor', n_bins=-1, loop_pixels=None):
    """``prepare_chrom`` for a chromosome sharded over the ranks by row range
    (BASELINE config 4): ``csr`` holds this rank's rows
    only (``staging.shard_rows``), ``bias_raw`` the whole chromosome's bias.
    Union, gathers, scaling and the filter are local to the row range; the
    size factors are the chromosome's (``dist.sharded_size_factor_table``), so
    every rank must call this for the same chromosomes in the same order.
    The state holds this rank's pixels; the ranks' states concatenated in rank
    order are the single-GPU state."""
    if n_bins == -1:
        n_bins = int(dist_max / 5)
    bias = ops.filter_bias(bias_raw, bias_thresh)
    u = ops.union_gather(csr, dist_max, bias)
    n_px = int(u['row'].numel())
    st = dict(bias=bias, row=u['row'], col=u['col'], raw=u['raw'])
    table = hdist.sharded_size_factor_table(u['balanced'], u['dist'], dist_max,
                                            n_bins, norm)
    if n_px == 0 or table is None:
        if 'conditional' not in norm and table is not None:
            _empty_prepared(st, csr.n_reps, u['balanced'], loop_pixels)
            st['size_factors'] = table
            return st
        return _empty_prepared(st, csr.n_reps, u['balanced'], loop_pixels)
    scaled, sf, disp_idx = ops.scale_filter(
        u['row'], u['col'], u['balanced'], table, design, dist_max,
        mean_thresh, dist_min)
    st.update(size_factors=sf, scaled=scaled, disp_idx=disp_idx,
              disp_index=ops.mask_to_index(disp_idx))
    if table.dim() == 2:
        st['sf_table'] = table
    if loop_pixels is not None:
        st['loop_idx'] = ops.loop_membership(st['row'], st['col'],
                                             st['disp_index'], loop_pixels)
    return st


_PREPARE_STREAMS = {}


def prepare_many(chrom_inputs, design, n_streams=None, sink=None, **kw):
    """``prepare_chrom`` over an iterable of (csr, bias_raw), up to
    ``n_streams`` chromosomes in flight, one CUDA stream each: a chromosome is
    ~25 short kernels, several of them latency-bound (rank scan, median
    select), and two host read-backs; interleaving hides both.  Results are
    identical to the sequential loop (each chromosome's kernels keep their
    order on their own stream).  ``sink(i, name, tensor)`` as in
    ``run_to_qvalues``."""
    if n_streams is None:
        n_streams = int(os.environ.get('H3D_PREPARE_STREAMS', '4'))
    main = torch.cuda.current_stream()
    dev_i = torch.cuda.current_device()
    streams = _PREPARE_STREAMS.setdefault(dev_i, [])
    while len(streams) < n_streams:
        streams.append(torch.cuda.Stream())
    states = {}
    active = []          # [index, stream, generator, pending Readback]

    def finish(i, stream, st):
        states[i] = st
        if sink is not None:
            with torch.cuda.stream(stream):
                for k in PREPARE_OUTPUTS:
                    emit_output(sink, i, k, st)

    def advance():
        i, stream, gen, pending = active.pop(0)
        with torch.cuda.stream(stream):
            try:
                active.append([i, stream, gen, gen.send(pending.get())])
            except StopIteration as stop:
                finish(i, stream, stop.value)

    for i, (csr, b) in enumerate(chrom_inputs):
        # chromosome i always runs on stream i mod n_streams: the caching
        # allocator then sees the same request sequence per stream every run
        stream = streams[i % n_streams]
        while any(a[1] is stream for a in active):
            advance()
        stream.wait_stream(main)              # inputs were made ready on main
        gen = prepare_chrom_steps(csr, b, design, **kw)
        with torch.cuda.stream(stream):
            try:
                active.append([i, stream, gen, next(gen)])
            except StopIteration as stop:
                finish(i, stream, stop.value)
    while active:
        advance()
    for s in streams:
        main.wait_stream(s)
    return [states[i] for i in sorted(states)]


def _pool_local_order(states, dist_max):
    """Distances of the tested pixels (chromosome order), their stable rank by
    pooling key and the key boundaries: shared first half of both pooling
    paths.  The key is the distance itself in a one-process run and (owner
    rank, distance) otherwise (hic3defdr_b200.dist.distance_keys)."""
    counts = [int(s['disp_index'].numel()) for s in states]
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    n_tot = int(offs[-1])
    dist_cat = torch.empty(n_tot, dtype=torch.int32, device='cuda')
    for s, o, n in zip(states, offs[:-1], counts):
        if s['size_factors'].dim() != 2:
            raise IndexError(
                'estimate_disp needs per-pixel size factors (a conditional '
                'norm), as in the reference (analysis/analysis.py:181)')
        if n:
            ops.gather_counts_factors(s['row'], s['col'], s['disp_index'],
                                      None, None, None, None, n_tot, None,
                                      None, dist_cat[o:o + n])
    key_of_dist, per = hdist.distance_keys(dist_max + 1)
    n_keys = per * hdist.world_size()
    if n_tot:
        keys = dist_cat if hdist.world_size() == 1 else \
            ops.dev(key_of_dist)[dist_cat.long()]
        rank, key_start = ops.stable_rank(keys, n_keys)
    else:
        rank = None
        key_start = torch.zeros(n_keys + 1, dtype=torch.int64, device='cuda')
    return counts, offs, dist_cat, rank, key_start, key_of_dist, per


def _pool_gather(states, counts, offs, rank, key_start, key_of_dist, per,
                 n_reps, owner_of_key, shift_of_key, bases, ld):
    """The two pooling passes (csrc/peer.cu): records at the pooled
    positions, then the coalesced gather into ``bases[owner]``."""
    import ctypes
    from hic3defdr_b200._native import lib, ptr
    n_tot = int(offs[-1])
    if not n_tot:
        return
    n_keys = len(owner_of_key)
    rec = torch.empty((n_tot, 2), dtype=torch.int32, device='cuda')
    table = np.zeros((len(states), 4), dtype=np.int64)
    for ci, (s, o, n) in enumerate(zip(states, offs[:-1], counts)):
        sf, mode = (s['sf_table'], 2) if 'sf_table' in s else \
            (s['size_factors'], 1)
        table[ci] = (s['raw'].data_ptr(), s['bias'].data_ptr(), sf.data_ptr(),
                     mode)
        if n:
            lib().call('h3d_pool_index', ptr(s['row']), ptr(s['disp_index']),
                       n, ci, ptr(rank[o:o + n]), ptr(rec), ops._stream())
    dist_of_key = np.full(n_keys, -1, dtype=np.int32)
    dist_of_key[key_of_dist] = np.arange(len(key_of_dist), dtype=np.int32)
    peer_arr = (ctypes.c_void_p * len(bases))(*bases)
    # named, so that the small tables stay allocated until the launch is queued
    # (a temporary's block would be handed to the next temporary at once)
    d_dist = ops.dev(dist_of_key)
    d_owner = ops.dev(np.asarray(owner_of_key, dtype=np.int32))
    d_shift = ops.dev(np.asarray(shift_of_key, dtype=np.int64))
    d_table = ops.dev(table)
    lib().call('h3d_pool_pull', ptr(rec), n_tot, ptr(key_start), n_keys,
               ptr(d_dist), ptr(d_owner), ptr(d_shift), ptr(d_table),
               len(states), n_reps, peer_arr, len(bases), int(ld),
               ops._stream())


def pool_by_distance(states, dist_max, n_reps=None):
    """Pools the disp_idx pixels of the given chromosomes by distance
    (analysis/analysis.py:169-183, 196-197) on this device: returns (x, f) SoA
    (R, n) in (pooling key, chromosome, row, col) order -- the two halves of one
    (2 R, n) buffer --, the per-pixel distances in chromosome order, the segment
    boundaries over the keys (host int64) and the per-chromosome offsets."""
    counts, offs, dist_cat, rank, key_start, key_of_dist, per = \
        _pool_local_order(states, dist_max)
    n_tot = int(offs[-1])
    if n_reps is None:
        # a rank without chromosomes must still take part in the exchange with
        # the same number of replicate rows as the others: callers that may run
        # multi-process pass the design's replicate count
        n_reps = states[0]['raw'].shape[1] if states else 1
    ld = max(n_tot, 1)
    xf = torch.empty((2 * n_reps, ld), dtype=torch.float64, device='cuda')
    n_keys = per * hdist.world_size()
    _pool_gather(states, counts, offs, rank, key_start, key_of_dist, per,
                 n_reps, np.zeros(n_keys, np.int32), np.zeros(n_keys, np.int64),
                 [xf.data_ptr()], ld)
    seg_start = ops.to_host(key_start)
    return xf[:n_reps], xf[n_reps:], dist_cat, seg_start, offs


def pool_to_owners(states, dist_max, n_reps):
    """``pool_by_distance`` + the exchange by distance owner in one pass over
    the pixels (multi-GPU, NCCL, one node): after the local stable rank by
    distance and ONE all-gather of the per-(rank, distance) counts, the owners
    are dealt by longest-processing-time on the genome-wide pixel counts
    (``dist.lpt_layout``) and the pooling gather writes every pixel's counts
    and factors straight into the buffer of the rank that owns its distance,
    over NVLink (csrc/peer.cu).  Returns (x, f, dist_cat, seg_start, offs,
    layout) with x, f the two halves of this rank's receive buffer and
    seg_start the boundaries of its owned distances."""
    ws, me = hdist.world_size(), hdist.rank()
    n_dist = dist_max + 1
    with hdist.single_process():         # local order: plain distance keys
        counts, offs, dist_cat, rank, key_start, key_of_dist, per = \
            _pool_local_order(states, dist_max)
    # the collective also orders the ranks: nobody writes into a buffer whose
    # previous contents are still being read (stream order on every rank)
    all_counts = hdist._all_gather_flat(
        (key_start[1:] - key_start[:-1]).contiguous()).cpu().numpy()
    lay = hdist.lpt_layout(all_counts, me)
    ld = (int(lay['n_recv'].max()) + 255) // 256 * 256 + 256
    ptrs = hdist._PEERS.ensure(2 * n_reps * ld * 8)
    _pool_gather(states, counts, offs, rank, key_start, key_of_dist, per,
                 n_reps, lay['owner'], lay['shift'], ptrs, ld)
    hdist.fence_peer_writes()
    own = ptrs[me]
    x = hdist.RawMatrix(own, (n_reps, ld))
    f = hdist.RawMatrix(own + n_reps * ld * 8, (n_reps, ld))
    return x, f, dist_cat, lay['seg_start'], offs, lay


def fit_trends(disp_per_dist, dist_max, cond_names, frac=None,
               auto_frac_factor=15., weighted_lowess=True, log=None):
    """analysis/analysis.py:208-218: one trend per condition; returns
    (list of callables, (dist_max + 1, C) table of their values).  The
    smoothing of all conditions is one kernel launch (trend.fit_many)."""
    n_conds = disp_per_dist.shape[1]
    table = np.full((dist_max + 1, n_conds), np.nan)
    specs = []
    for c in range(n_conds):
        if log:
            log('  estimating dispersion for condition %s' % cond_names[c])
            log('  fitting distance vs dispersion relationship')
        idx = np.isfinite(disp_per_dist[:, c])
        xs = np.arange(dist_max + 1)[idx]
        ys = disp_per_dist[:, c][idx]
        kwargs = {'left_boundary': ys[0]}
        if frac is not None:
            kwargs['frac'] = frac
        if weighted_lowess:
            kwargs['auto_frac_factor'] = auto_frac_factor
        specs.append((xs, ys, kwargs))
    fns = fit_many(specs, weighted=weighted_lowess)
    grid = np.arange(dist_max + 1)
    for c, fn in enumerate(fns):
        table[:, c] = fn(grid)
    return fns, table


def _estimate_with_callable(x, f, runs, n_seg, design, estimator):
    """analysis/analysis.py:164-165, 186-206 with a user-supplied Python
    estimator ``(data (pixels, replicates), f=...) -> float``: the pooled
    pixels of every (distance, condition) bin are handed to it on the host, as
    the reference does; everything around it (pooling, exchange, trend, gather)
    stays on the device."""
    xs, fs = x.cpu().numpy(), f.cpu().numpy()
    out = np.full((n_seg, design.shape[1]), np.nan)
    run_seg, run_lo, run_hi = runs
    for s in range(n_seg):
        cols = np.concatenate(
            [np.arange(lo, hi) for g, lo, hi in zip(run_seg, run_lo, run_hi)
             if g == s] + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
        if not len(cols):
            continue
        for c in range(design.shape[1]):
            reps = np.flatnonzero(design[:, c])
            raw_slice = np.ascontiguousarray(
                xs[np.ix_(reps, cols)].T).astype(np.int64)
            f_slice = np.ascontiguousarray(fs[np.ix_(reps, cols)].T)
            out[s, c] = estimator(raw_slice, f=f_slice)
    return out, dict(outer_iterations=0, nll_evaluations=0,
                     pixel_equalizations=0, launches=0, equalize_launches=0,
                     equalize_us=0, nll_launches=0, nll_us=0,
                     capped_segments=0)


def estimate_disp(states, design, dist_max, cond_names=None, estimator='qcml',
                  frac=None, auto_frac_factor=15., weighted_lowess=True,
                  log=None):
    """analysis/analysis.py:163-223 on device states (this rank's
    chromosomes).  Sets ``disp`` in every state; returns
    (disp_per_dist (D+1, C) numpy, list of trend callables, stats)."""
    n_conds = design.shape[1]
    if cond_names is None:
        cond_names = [str(c) for c in range(n_conds)]
    if hdist.peer_exchange_enabled() and not callable(estimator):
        # multi-GPU: the pooling gather writes into the owners' buffers
        with stage('estimate_disp/pool'):
            x, f, dist_cat, seg_own, offs, lay = pool_to_owners(
                states, dist_max, design.shape[0])
        n_tot = int(offs[-1])
        with stage('estimate_disp/qcml'):
            if len(seg_own) > 1:
                disp_own, stats = ops.estimate_dispersion(
                    x, f, seg_own, design, estimator)
            else:
                disp_own, stats = np.zeros((0, n_conds)), None
        with stage('estimate_disp/merge'):
            disp_per_dist = hdist.merge_disp_owned(disp_own, lay['owned'],
                                                   dist_max + 1)
        if stats is None:
            stats = dict(outer_iterations=0, nll_evaluations=0,
                         pixel_equalizations=0, launches=0,
                         equalize_launches=0, equalize_us=0, nll_launches=0,
                         nll_us=0, capped_segments=0)
    else:
        with stage('estimate_disp/pool'):
            x, f, dist_cat, seg_start, offs = pool_by_distance(
                states, dist_max, n_reps=design.shape[0])
        n_tot = int(offs[-1])
        # multi-GPU: every distance is estimated on the rank that owns it
        n_keys = len(seg_start) - 1
        with stage('estimate_disp/exchange'):
            x, f, runs = hdist.exchange_by_distance(x, f, seg_start, n_tot)
        n_owned = n_keys // hdist.world_size()
        with stage('estimate_disp/qcml'):
            if callable(estimator):
                disp_per_dist, stats = _estimate_with_callable(
                    x, f, runs, n_owned, design, estimator)
            else:
                disp_per_dist, stats = ops.estimate_dispersion(
                    x, f, n_owned, design, estimator, runs=runs)
        with stage('estimate_disp/merge'):
            disp_per_dist = hdist.merge_disp_per_dist(disp_per_dist,
                                                      dist_max + 1)
    del x, f
    with stage('estimate_disp/trend'):
        fns, table = fit_trends(disp_per_dist, dist_max, cond_names, frac,
                                auto_frac_factor, weighted_lowess, log)
        disp = ops.gather_table(dist_cat, table) if n_tot else torch.empty(
            (0, n_conds), dtype=torch.float64, device='cuda')
    disp_table = ops.dev(table)
    for i, s in enumerate(states):
        s['disp'] = disp[offs[i]:offs[i + 1]]
        s['disp_table'] = disp_table         # disp = disp_table[col - row]
    return disp_per_dist, fns, stats


def lrt_chrom(st, design, refit_mu=True, failed=None):
    """analysis/analysis.py:261-284 for one chromosome state."""
    p, llr, mu0, mu1 = ops.lrt_fused(
        st['row'], st['col'], st['disp_index'], st['raw'], st['size_factors'],
        st['bias'], st['disp'], design, refit_mu, failed)
    st.update(pvalues=p, llr=llr, mu_hat_null=mu0, mu_hat_alt=mu1)
    return st


def bh(states, use_loop_idx=False):
    """analysis/analysis.py:295-303 over this rank's chromosome states (and,
    through hic3defdr_b200.dist, every other rank's)."""
    ps = []
    for s in states:
        p = s['pvalues']
        if use_loop_idx:
            p = p[s['loop_idx'].bool()]
        ps.append(p)
    counts = [int(p.numel()) for p in ps]
    local = torch.cat(ps) if ps else torch.empty(0, dtype=torch.float64,
                                                 device='cuda')
    q = hdist.global_bh(local)
    offs = np.concatenate([[0], np.cumsum(counts)])
    for i, s in enumerate(states):
        s['qvalues'] = q[offs[i]:offs[i + 1]]
    return states


PREPARE_OUTPUTS = ('row', 'col', 'disp_idx', 'raw', 'size_factors', 'scaled')
LRT_OUTPUTS = ('pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt')


def emit_output(sink, i, name, st):
    """Hands output ``name`` of chromosome state ``st`` to ``sink``.  A sink
    that declares ``compact`` (staging.OutputDrain) receives the arrays that
    are functions of the pixel distance as their per-distance tables and
    ``raw`` through its narrowing entry; a plain callable gets every tensor."""
    if sink is None:
        return
    if getattr(sink, 'compact', False):
        if name == 'size_factors' and 'sf_table' in st:
            return sink.table(i, name, st['sf_table'], int(st['row'].numel()))
        if name == 'disp' and 'disp_table' in st:
            return sink.table(i, name, st['disp_table'],
                              int(st['disp_index'].numel()), mask='disp_idx')
        if name == 'raw' and st['raw'].dtype == torch.int64:
            return sink.narrow(i, name, st['raw'])
    sink(i, name, st[name])


def run_to_qvalues(chrom_inputs, design, dist_min=4, dist_max=200,
                   bias_thresh=0.1, mean_thresh=1.0, norm='conditional_mor',
                   n_bins=-1, estimator='qcml', frac=None,
                   auto_frac_factor=15., weighted_lowess=True, refit_mu=True,
                   sink=None, row_sharded=False):
    """All four steps on device inputs: ``chrom_inputs`` is an iterable of
    (ops.DeviceCSR, bias_raw CUDA tensor) for THIS rank's chromosomes (a list,
    or a ``staging.InputPrefetcher`` that uploads ahead of the kernels).
    ``sink(i, name, tensor)``, if given, is called for every output array of
    chromosome i as soon as it is final (``staging.OutputDrain`` starts its
    device -> host copy there).
    ``row_sharded``: every rank holds a row range of EVERY chromosome
    (``staging.shard_rows``) instead of whole chromosomes; the states are then
    this rank's pixels of each chromosome.
    Returns (states, disp_per_dist, trend callables, qcml stats)."""
    design = np.asarray(design).astype(bool)
    emit = sink if sink is not None else (lambda i, name, t: None)
    with stage('prepare_data'):
        kw = dict(dist_min=dist_min, dist_max=dist_max,
                  bias_thresh=bias_thresh, mean_thresh=mean_thresh, norm=norm,
                  n_bins=n_bins)
        if row_sharded and hdist.world_size() > 1:
            states = []
            for i, (csr, b) in enumerate(chrom_inputs):
                states.append(prepare_chrom_sharded(csr, b, design, **kw))
                for k in PREPARE_OUTPUTS:
                    emit_output(sink, i, k, states[-1])
        else:
            states = prepare_many(chrom_inputs, design, sink=sink, **kw)
    with stage('estimate_disp'):
        dpd, fns, stats = estimate_disp(states, design, dist_max,
                                        estimator=estimator, frac=frac,
                                        auto_frac_factor=auto_frac_factor,
                                        weighted_lowess=weighted_lowess)
        for i, s in enumerate(states):
            emit_output(sink, i, 'disp', s)
    with stage('lrt'):
        failed = torch.zeros(1, dtype=torch.int32, device='cuda')
        for i, s in enumerate(states):
            lrt_chrom(s, design, refit_mu, failed)
            for k in LRT_OUTPUTS:
                emit(i, k, s[k])
        ops.check_failed(failed, 'lrt')
    with stage('bh'):
        bh(states)
        for i, s in enumerate(states):
            emit(i, 'qvalues', s['qvalues'])
    return states, dpd, fns, stats
