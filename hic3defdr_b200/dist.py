"""
Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over
NVLink on the GPU box, gloo in the CPU tests).

How the path shards:
  * prepare_data / lrt: per chromosome, no exchange -> chromosomes are dealt to
    ranks by longest-processing-time on the input size;
  * estimate_disp: dispersion is pooled genome-wide per distance
    (hic3defdr/analysis/analysis.py:169-206) -> ONE all-to-all moves every
    rank's pixels of distance d to the rank owning d (distances dealt round
    robin: balances pixel counts and qCML iterations); the (distance,
    condition) optimisations then run without any collective, and the (D+1, C)
    result is all-gathered;
  * bh: one global correction (analysis.py:296-303) -> distributed sort/rank:
    splitters from a gathered sample, one all-to-all of the p-values (8 B per
    tested pixel over NVLink) to the owner of their value range, local ranking
    with global rank offsets, all-gather of one carry per rank, q-values back
    by the mirrored all-to-all.
  * a chromosome too large (or too alone) for one GPU is sharded by ROW RANGE
    (``row_ranges``): union, scaling, LRT stay local; only the size factors
    need the whole chromosome -> ``sharded_size_factor_table``: all-gather of
    the per-distance pixel counts (fixes every pixel's chromosome-wide rank and
    so its equal-count bin), one all-to-all of the per-pixel ratios to the
    owner of their bin, exact medians there, all-gather of the (bins, R)
    medians.
Single-process runs take none of these branches.
"""
import numpy as np
import torch
import torch.distributed as td


import contextlib

_SINGLE = [False]


@contextlib.contextmanager
def single_process():
    """Inside this context the calling process behaves as a one-process run
    even though a process group exists (no collectives are issued): used to
    recompute a result on one rank for comparison with the sharded run."""
    _SINGLE[0] = True
    try:
        yield
    finally:
        _SINGLE[0] = False


def initialized():
    return td.is_available() and td.is_initialized() and not _SINGLE[0]


def rank():
    return td.get_rank() if initialized() else 0


def world_size():
    return td.get_world_size() if initialized() else 1


def barrier():
    if initialized() and world_size() > 1:
        td.barrier()


def broadcast_flags(flags):
    """rank 0's list of booleans on every rank (decisions that guard code
    containing collectives must be the same everywhere)"""
    if not (initialized() and world_size() > 1):
        return [bool(v) for v in flags]
    t = torch.tensor([int(bool(v)) for v in flags], dtype=torch.int64)
    if td.get_backend() == 'nccl':
        t = t.cuda()
    td.broadcast(t, src=0)
    return [bool(v) for v in t.cpu().tolist()]


def lpt_assign(weights, n_ranks):
    """Longest-processing-time assignment: returns owner rank per item."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    load = [0] * n_ranks
    owner = [0] * len(weights)
    for i in order:
        r = min(range(n_ranks), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += weights[i]
    return owner


def shard_chroms(chroms, weight_fn):
    """The chromosomes this rank owns (all of them in a 1-process run), in the
    order of ``chroms``."""
    ws = world_size()
    if ws == 1:
        return list(chroms)
    owner = lpt_assign([weight_fn(c) for c in chroms], ws)
    return [c for c, o in zip(chroms, owner) if o == rank()]


def row_ranges(weight_per_row, n_ranks=None):
    """Row range [lo, hi) of every rank for a row-sharded chromosome: cuts the
    rows where the cumulative weight (stored entries per row, summed over the
    replicates) crosses k / n_ranks of the total.  Returns (n_ranks + 1,) int64
    boundaries, identical on every rank."""
    ws = world_size() if n_ranks is None else n_ranks
    w = np.asarray(weight_per_row, dtype=np.float64)
    n = len(w)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    if cum[-1] <= 0:
        return np.linspace(0, n, ws + 1).astype(np.int64)
    cuts = np.searchsorted(cum, cum[-1] * np.arange(1, ws) / ws, side='left')
    return np.concatenate([[0], np.minimum(cuts, n), [n]]).astype(np.int64)


def _all_gather_flat(t):
    """every rank's tensor ``t`` (same shape everywhere) stacked along a new
    first dimension: one collective into one flat buffer (the layout both the
    NCCL and the gloo backend accept)"""
    ws = world_size()
    out = torch.empty(ws * t.numel(), dtype=t.dtype, device=t.device)
    td.all_gather_into_tensor(out, t.contiguous().view(-1))
    return out.view((ws,) + tuple(t.shape))


def _all_gather_counts(local_counts):
    """(world_size, len(local_counts)) int64 table of every rank's counts: one
    collective into one tensor, one device -> host copy."""
    t = torch.from_numpy(np.ascontiguousarray(local_counts, dtype=np.int64))
    if td.get_backend() == 'nccl':
        t = t.cuda()
    return _all_gather_flat(t).cpu().numpy()


def regroup_positions(recv_counts, device):
    """Received layout is [source rank][owned distance] (runs of
    ``recv_counts[s, d]`` pixels); wanted is [owned distance][source rank].
    Returns the int64 destination of every received element: run (s, d) starts
    at R[s, d] on arrival and at T[d, s] in the regrouped array, so
    ``pos = arange + repeat(T - R)`` -- no sort, no host-side expansion."""
    c = np.asarray(recv_counts, dtype=np.int64)               # (ws, owned)
    n = int(c.sum())
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=device)
    flat = c.ravel()                                          # arrival order
    r_start = np.concatenate([[0], np.cumsum(flat)[:-1]])
    t_start = np.concatenate([[0], np.cumsum(c.T.ravel())[:-1]]) \
        .reshape(c.shape[1], c.shape[0]).T.ravel()            # T[d, s] at (s, d)
    shift = torch.from_numpy(t_start - r_start).to(device)
    reps = torch.from_numpy(flat).to(device)
    pos = torch.arange(n, dtype=torch.int64, device=device)
    if n:
        pos += torch.repeat_interleave(shift, reps, output_size=n)
    return pos


def distance_owner(n_dist, n_ranks=None):
    """Owner rank of every distance.  Distances are dealt in boustrophedon
    order (0 .. G-1, G-1 .. 0, 0 .. G-1, ...): the cost of a (distance,
    condition) bin falls smoothly with the distance (fewer pixels, fewer counts
    per pixel, shorter incomplete-gamma loops), so plain round robin hands rank
    0 the more expensive member of every group of G (equalize 11 % above the
    mean at G = 8); reversing every other group cancels the trend to first
    order."""
    ws = world_size() if n_ranks is None else n_ranks
    d = np.arange(n_dist)
    grp, pos = d // ws, d % ws
    return np.where(grp % 2 == 0, pos, ws - 1 - pos).astype(np.int64)


def distance_keys(n_dist, n_ranks=None):
    """Pooling key of every distance: the key orders a rank's pooled pixels by
    (owner, distance): key(d) = owner(d) * per + d // n_ranks,
    per = ceil(n_dist / n_ranks) (every rank owns one distance of every group
    of n_ranks consecutive ones, ``distance_owner``).  Returns (key per
    distance int32, per).  Identity for one process."""
    ws = world_size() if n_ranks is None else n_ranks
    per = -(-n_dist // ws)
    d = np.arange(n_dist)
    return (distance_owner(n_dist, ws) * per + d // ws).astype(np.int32), per


def owned_distances(n_dist, me=None, n_ranks=None):
    """distances owned by rank ``me``, in key order"""
    ws = world_size() if n_ranks is None else n_ranks
    me = rank() if me is None else me
    return np.flatnonzero(distance_owner(n_dist, ws) == me)


def exchange_by_distance(x, f, seg_start, n_local):
    """x, f: (R, ld) pooled on this rank in ``distance_keys`` order;
    seg_start: (ws * per + 1) local group boundaries over the keys.  Returns
    (x, f, runs): the arrays hold ALL ranks' pixels of the distances this rank
    owns, ``runs`` = (run_seg, run_lo, run_hi) int arrays describing where:
    run r is the pixels [run_lo[r], run_hi[r]) of owned distance run_seg[r]
    (``per`` owned keys, the trailing ones empty when the distances do not
    divide evenly).  A distance's pixels arrive as one run per source rank, in
    rank order, and stay where they land: the dispersion kernels sum a bin
    exactly, whatever its pixel order (ops.estimate_dispersion, ``runs=``).

    One grouped set of point-to-point transfers (a single NCCL group: every
    replicate row of x and f, to and from every peer) instead of one collective
    per row."""
    ws, me = world_size(), rank()
    n_keys = len(seg_start) - 1
    if ws == 1:
        seg = np.asarray(seg_start, dtype=np.int64)
        return x, f, (np.arange(n_keys, dtype=np.int32), seg[:-1].copy(),
                      seg[1:].copy())
    per = n_keys // ws
    local_counts = np.diff(seg_start)
    all_counts = _all_gather_counts(local_counts)          # (ws, ws * per)
    n_reps = x.shape[0]
    soff = np.asarray(seg_start, dtype=np.int64)[::per]     # (ws + 1,) send offsets
    recv_counts = all_counts[:, me * per:(me + 1) * per]    # (ws, per)
    roff = np.concatenate([[0], np.cumsum(recv_counts.sum(axis=1))]) \
        .astype(np.int64)                                   # (ws + 1,) receive offsets
    n_recv = int(roff[-1])
    xr = torch.empty((n_reps, max(n_recv, 1)), dtype=x.dtype, device=x.device)
    fr = torch.empty_like(xr)
    p2p = []
    for src, dst in ((x, xr), (f, fr)):
        a, b = int(soff[me]), int(soff[me + 1])
        if b > a:                                   # this rank's own share
            dst[:, int(roff[me]):int(roff[me + 1])].copy_(src[:, a:b])
        for r in range(n_reps):
            for k in range(ws):
                if k == me:
                    continue
                a, b = int(soff[k]), int(soff[k + 1])
                if b > a:
                    p2p.append(td.P2POp(td.isend, src[r, a:b], k))
                a, b = int(roff[k]), int(roff[k + 1])
                if b > a:
                    p2p.append(td.P2POp(td.irecv, dst[r, a:b], k))
    if p2p:
        for req in td.batch_isend_irecv(p2p):
            req.wait()
    # run (s, j): source rank s, owned key j
    lo = roff[:-1, None] + np.cumsum(recv_counts, axis=1) - recv_counts
    hi = lo + recv_counts
    seg_of = np.broadcast_to(np.arange(per, dtype=np.int32)[None, :],
                             recv_counts.shape)
    order = np.argsort(seg_of.ravel(), kind='stable')       # by key, then source
    runs = (np.ascontiguousarray(seg_of.ravel()[order].astype(np.int32)),
            np.ascontiguousarray(lo.ravel()[order].astype(np.int64)),
            np.ascontiguousarray(hi.ravel()[order].astype(np.int64)))
    return xr, fr, runs


def owner_layout(all_counts, per, me):
    """Where every rank's pooled pixels live after the exchange by distance
    owner.  ``all_counts``: (ws, ws * per) pixels per (source rank, key), key =
    owner * per + owned index (``distance_keys``).  The buffer of owner k holds
    one run per (source s, owned key j), source-major:
    ``lo_k[s, j] = sum(counts[:s, k's keys]) + sum(counts[s, k's keys before j])``.
    Returns (n_recv (ws,), shift (ws * per,), runs): n_recv[k] = pixels owner k
    receives; shift[key] = position of THIS rank's first pixel of ``key`` in
    its owner's buffer minus its position in this rank's own pooled order (so a
    pixel at local pooled position p goes to p + shift[key]); runs = (run_seg,
    run_lo, run_hi) of THIS rank's buffer, ordered by key, then source."""
    c = np.asarray(all_counts, dtype=np.int64)
    ws = c.shape[0]
    blocks = c.reshape(ws, ws, per)                         # [source, owner, j]
    n_recv = blocks.sum(axis=(0, 2))                        # (ws,)
    per_src = blocks.sum(axis=2)                            # [source, owner]
    src_off = np.cumsum(per_src, axis=0) - per_src          # [source, owner]
    within = np.cumsum(blocks, axis=2) - blocks             # [source, owner, j]
    lo = src_off[:, :, None] + within                       # [source, owner, j]
    local_start = np.cumsum(c[me]) - c[me]                  # (ws * per,)
    shift = lo[me].reshape(-1) - local_start
    mine_lo = lo[:, me, :]                                  # [source, j]
    mine_cnt = blocks[:, me, :]
    seg_of = np.broadcast_to(np.arange(per, dtype=np.int32)[None, :],
                             mine_cnt.shape)
    order = np.argsort(seg_of.ravel(), kind='stable')
    runs = (np.ascontiguousarray(seg_of.ravel()[order].astype(np.int32)),
            np.ascontiguousarray(mine_lo.ravel()[order]),
            np.ascontiguousarray((mine_lo + mine_cnt).ravel()[order]))
    return n_recv, shift, runs


def lpt_layout(all_counts, me, weight=None):
    """Ownership and buffer layout of the peer-memory pooling when the owners
    are chosen AFTER the per-(rank, distance) pixel counts are known
    (``all_counts``: (ws, n_dist), every rank's local pooled order is plain
    distance order).  Distances are dealt to ranks by longest-processing-time
    on ``weight`` (default: the genome-wide pixel count of the distance;
    distances without pixels weigh nothing), identically on every rank.  The
    buffer of owner k holds its owned distances in increasing order, inside a
    distance one run per source rank in rank order -- i.e. every owned distance
    is one contiguous segment.
    Returns dict(owner (n_dist,), owned (list of arrays per rank), shift
    (n_dist,) for THIS rank: pooled position + shift = position in the owner's
    buffer, seg_start (n_owned + 1,) of THIS rank's buffer, n_recv (ws,))."""
    c = np.asarray(all_counts, dtype=np.int64)
    ws, n_dist = c.shape
    tot = c.sum(axis=0)
    w = tot.astype(float) if weight is None else np.asarray(weight, float)
    owner = np.asarray(lpt_assign([float(v) for v in w], ws), dtype=np.int64)
    # distances without pixels: spread them, they cost nothing
    owned = [np.flatnonzero(owner == k) for k in range(ws)]
    shift = np.zeros(n_dist, dtype=np.int64)
    local_start = np.cumsum(c[me]) - c[me]
    n_recv = np.zeros(ws, dtype=np.int64)
    seg_start = None
    for k in range(ws):
        d = owned[k]
        seg = np.concatenate([[0], np.cumsum(tot[d])]).astype(np.int64)
        n_recv[k] = seg[-1]
        before_me = c[:me][:, d].sum(axis=0)             # lower ranks' share
        shift[d] = seg[:-1] + before_me - local_start[d]
        if k == me:
            seg_start = seg
    return dict(owner=owner, owned=owned, shift=shift, seg_start=seg_start,
                n_recv=n_recv)


def merge_disp_owned(disp_owned, owned, n_dist):
    """``disp_owned``: (len(owned[me]), C) results of this rank's distances;
    ``owned``: every rank's distance list (``lpt_layout``).  Returns the full
    (n_dist, C) table on every rank."""
    ws, me = world_size(), rank()
    n_conds = disp_owned.shape[1]
    cap = max(1, max(len(d) for d in owned))
    pad = np.full((cap, n_conds), np.nan)
    pad[:len(owned[me])] = disp_owned[:len(owned[me])]
    out = _all_gather_flat(_coll_tensor(pad)).cpu().numpy()
    full = np.full((n_dist, n_conds), np.nan)
    for k in range(ws):
        full[owned[k]] = out[k][:len(owned[k])]
    return full


class RawMatrix(object):
    """a (rows, ld) float64 matrix in device memory that torch did not
    allocate (a peer-shared receive buffer): just what the C-ABI calls need"""

    def __init__(self, address, shape):
        self.address, self.shape = int(address), tuple(shape)

    def data_ptr(self):
        return self.address


class PeerBuffers(object):
    """This rank's receive buffer for the distance exchange and the other
    ranks' buffers opened through CUDA IPC (include/h3d.h, h3d_peer_*).  The
    buffer persists across steps and only grows; every rank calls ``ensure``
    with the same size at the same point."""

    def __init__(self):
        self.nbytes, self.own, self.ptrs = 0, None, None

    def ensure(self, nbytes):
        import ctypes
        from hic3defdr_b200._native import lib
        if nbytes <= self.nbytes:
            return self.ptrs
        self.release()
        nbytes = int(nbytes * 1.25) // 4096 * 4096 + 4096
        own = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        lib().call('h3d_peer_alloc', nbytes, ctypes.byref(own), handle)
        handles = [None] * world_size()
        td.all_gather_object(handles, bytes(handle))
        ptrs = []
        for k, h in enumerate(handles):
            if k == rank():
                ptrs.append(own.value)
                continue
            p = ctypes.c_void_p()
            lib().call('h3d_peer_open',
                       (ctypes.c_ubyte * 64).from_buffer_copy(h),
                       ctypes.byref(p))
            ptrs.append(p.value)
        self.nbytes, self.own, self.ptrs = nbytes, own.value, ptrs
        barrier()
        return ptrs

    def release(self):
        from hic3defdr_b200._native import lib
        if self.own is None:
            return
        torch.cuda.synchronize()
        barrier()                       # nobody still writes into a buffer
        for k, p in enumerate(self.ptrs):
            if k != rank():
                lib().call('h3d_peer_close', p)
        barrier()                       # every mapping is closed before the free
        lib().call('h3d_peer_free', self.own)
        self.nbytes, self.own, self.ptrs = 0, None, None


_PEERS = PeerBuffers()            # dispersion pooling
_BH_PEERS = PeerBuffers()         # p-values out, q-values back


class _CudaArray(object):
    def __init__(self, address, n, typestr):
        self.__cuda_array_interface__ = dict(
            shape=(int(n),), typestr=typestr, data=(int(address), False),
            version=2)


def _raw_tensor(address, n, dtype):
    """zero-copy 1-D torch view of device memory torch did not allocate"""
    if n == 0:
        return torch.empty(0, dtype=dtype, device='cuda')
    assert dtype == torch.float64
    return torch.as_tensor(_CudaArray(address, n, '<f8'), device='cuda')


def _peer_copy(src, src_off, ptrs, dst_off, nbytes):
    """slice k of ``src`` (byte offsets) into rank k's buffer"""
    import ctypes
    from hic3defdr_b200._native import lib, ptr
    ws = len(ptrs)
    for k0 in range(0, ws, 16):
        k1 = min(ws, k0 + 16)
        so = np.ascontiguousarray(src_off[k0:k1], dtype=np.int64)
        do = np.ascontiguousarray(dst_off[k0:k1], dtype=np.int64)
        nb = np.ascontiguousarray(nbytes[k0:k1], dtype=np.int64)
        bases = (ctypes.c_void_p * (k1 - k0))(*ptrs[k0:k1])
        lib().call('h3d_peer_copy', ptr(src), ptr(so), bases, ptr(do),
                   ptr(nb), k1 - k0, torch.cuda.current_stream().cuda_stream)


_PEER_OK = {}


def _probe_peer_memory():
    """One collective probe per process group: every rank allocates a small
    buffer, shares its CUDA IPC handle and opens everybody else's.  Peer
    memory is used only if that worked on EVERY rank (e.g. not across nodes,
    not where IPC is disabled); otherwise the NCCL transfers are."""
    import ctypes
    from hic3defdr_b200._native import lib
    ok, own, opened, why = 1, ctypes.c_void_p(), [], ''
    mine = None
    try:
        handle = (ctypes.c_ubyte * 64)()
        lib().call('h3d_peer_alloc', 4096, ctypes.byref(own), handle)
        mine = bytes(handle)
    except Exception as e:                       # noqa: BLE001
        ok, why = 0, str(e)
    handles = [None] * world_size()
    td.all_gather_object(handles, mine)          # every rank takes part
    if ok and all(h is not None for h in handles):
        try:
            for k, h in enumerate(handles):
                if k == rank():
                    continue
                p = ctypes.c_void_p()
                lib().call('h3d_peer_open',
                           (ctypes.c_ubyte * 64).from_buffer_copy(h),
                           ctypes.byref(p))
                opened.append(p.value)
        except Exception as e:                   # noqa: BLE001
            ok, why = 0, str(e)
    else:
        ok = 0
    if not ok:
        import sys
        print('hic3defdr_b200: CUDA IPC peer memory unavailable (%s); using '
              'NCCL transfers for the exchanges' % why, file=sys.stderr)
    t = torch.tensor([ok], dtype=torch.int32, device='cuda')
    td.all_reduce(t, op=td.ReduceOp.MIN)
    for p in opened:
        try:
            lib().call('h3d_peer_close', p)
        except Exception:                        # noqa: BLE001
            pass
    barrier()
    if own.value:
        try:
            lib().call('h3d_peer_free', own.value)
        except Exception:                        # noqa: BLE001
            pass
    return bool(int(t.item()))


def peer_exchange_enabled():
    """The pooling gather writes straight into the owners' buffers over
    NVLink (csrc/peer.cu) when the ranks run NCCL on one node and CUDA IPC
    works between all of them (probed once, collectively); ``H3D_EXCHANGE=
    nccl`` selects the grouped point-to-point transfers instead."""
    import os
    if not (initialized() and world_size() > 1 and
            td.get_backend() == 'nccl' and
            os.environ.get('H3D_EXCHANGE', 'peer') == 'peer'):
        return False
    key = world_size()
    if key not in _PEER_OK:
        _PEER_OK[key] = _probe_peer_memory()
    return _PEER_OK[key]


def fence_peer_writes():
    """every rank's peer writes issued so far (stream order) have landed when
    this returns on the stream: a one-element all-reduce"""
    t = torch.zeros(1, dtype=torch.int32, device='cuda')
    td.all_reduce(t)


def merge_disp_per_dist(disp_owned, n_dist):
    """``disp_owned``: (per, C) results of this rank's owned distances (key
    order).  Returns the full (n_dist, C) table on every rank."""
    ws = world_size()
    if ws == 1:
        return disp_owned[:n_dist]
    per, n_conds = disp_owned.shape
    t = torch.from_numpy(np.ascontiguousarray(disp_owned))
    if td.get_backend() == 'nccl':
        t = t.cuda()
    out = _all_gather_flat(t).cpu().numpy()
    full = np.full((n_dist, n_conds), np.nan)
    for k in range(ws):
        d = owned_distances(n_dist, k, ws)
        full[d] = out[k][:len(d)]
    return full


def all_gather_varlen(local):
    """Concatenation of every rank's 1-D tensor, plus the per-rank sizes."""
    ws = world_size()
    sizes = _all_gather_counts(np.array([local.numel()]))[:, 0]
    m = int(sizes.max()) if len(sizes) else 0
    pad = torch.zeros(max(m, 1), dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    out = [torch.empty_like(pad) for _ in range(ws)]
    td.all_gather(out, pad)
    return torch.cat([o[:int(s)] for o, s in zip(out, sizes)]), sizes


BH_SAMPLES = 4096      # splitter candidates contributed by each rank


def sample_positions(n, take, device=None):
    """``take`` evenly strided positions in [0, n), first and last included.
    Integer arithmetic: a float32 ``linspace(0, n - 1, take)`` rounds n - 1 up
    past the end once n exceeds 2^24 (a rank of the mouse genome holds 19 M
    p-values)."""
    return (torch.arange(take, dtype=torch.int64, device=device) *
            (n - 1)) // max(take - 1, 1)


def _partition(bucket, n_buckets):
    """Stable partition by bucket id: (position of every element in
    bucket-major order, bucket sizes as an int64 tensor on the same device)."""
    if bucket.is_cuda:
        from hic3defdr_b200 import ops
        pos, start = ops.stable_rank(bucket, n_buckets)
        return pos.long(), start[1:] - start[:-1]
    order = torch.argsort(bucket, stable=True)
    pos = torch.empty_like(order)
    pos[order] = torch.arange(len(order))
    return pos, torch.bincount(bucket.long(), minlength=n_buckets)


def global_bh(local_p, bh_fn=None, bh_ranked_fn=None, carry_fn=None):
    """BH over the p-values of all ranks (hic3defdr/analysis/analysis.py:
    296-303); returns this rank's q-values.

    Multi-process: a distributed sort/rank.  Splitters from an all-gathered
    sample cut [0, 1] into one value range per rank; ONE all-to-all moves every
    p-value (8 B) to the owner of its range; the owner ranks its bucket locally
    (global rank = local rank + sizes of the lower buckets), computes
    p n / rank and the running minimum inside the bucket; the bucket minima
    (one double per rank) are all-gathered for the carry across buckets; the
    q-values return by the mirrored all-to-all.  Equal p-values always land in
    the same bucket, so the result is identical to the single-process one.
    The host waits for the device once (the all-to-all split sizes)."""
    if bh_fn is None:
        from hic3defdr_b200 import ops
        bh_fn, bh_ranked_fn, carry_fn = ops.adjust_pvalues, \
            ops.adjust_pvalues_ranked, ops.apply_bh_carry
    ws, me = world_size(), rank()
    if ws == 1:
        return bh_fn(local_p)
    n = local_p.numel()
    device = local_p.device
    finite = torch.isfinite(local_p)
    n_fin_t = finite.sum()
    # splitters: an evenly strided local sample (non-finite entries count as
    # +inf and fall out below; the sample only decides the balance of the
    # buckets, never the result)
    sample = torch.full((BH_SAMPLES,), float('inf'), dtype=local_p.dtype,
                        device=device)
    if n:
        take = min(BH_SAMPLES, n)
        sel = sample_positions(n, take, device)
        vals = local_p[sel]
        sample[:take] = torch.where(torch.isfinite(vals), vals,
                                    torch.full_like(vals, float('inf')))
    gathered = _all_gather_flat(sample)
    allsamp = torch.sort(gathered.view(-1)).values
    n_samp = torch.isfinite(allsamp).sum()
    cut = torch.clamp((n_samp * torch.arange(1, ws, device=device)) // ws,
                      max=None, min=0)
    cut = torch.minimum(cut, torch.clamp(n_samp - 1, min=0))
    splitters = allsamp[cut]
    # bucket of every local value (non-finite: last bucket), stable partition
    bucket = torch.bucketize(local_p, splitters).to(torch.int32)
    bucket = torch.where(finite, bucket, torch.full_like(bucket, ws - 1))
    pos, sizes = _partition(bucket, ws)
    send = torch.empty_like(local_p)
    send[pos] = local_p
    info = torch.cat([sizes.to(torch.int64), n_fin_t.to(torch.int64).view(1)])
    table = _all_gather_flat(info).cpu().numpy()     # (ws, ws + 1): the one sync
    counts, n_total = table[:, :ws], int(table[:, ws].sum())
    send_splits = [int(v) for v in counts[me]]
    recv_splits = [int(v) for v in counts[:, me]]
    peer = local_p.is_cuda and peer_exchange_enabled()
    if peer:
        # both all-to-alls as direct stores into the peers' buffers (NVLink,
        # csrc/peer.cu): region A receives the p-values of this rank's value
        # range, region B the q-values of this rank's own pixels
        cap_a = int(counts.sum(axis=0).max())
        cap_b = int(counts.sum(axis=1).max())
        ptrs = _BH_PEERS.ensure((cap_a + cap_b) * 8)
        src_off = np.concatenate([[0], np.cumsum(counts[me])[:-1]]) * 8
        dst_off = counts[:me].sum(axis=0) * 8            # my slot in owner k's region A
        _peer_copy(send, src_off, ptrs, dst_off, counts[me] * 8)
        fence_peer_writes()
        recv = _raw_tensor(ptrs[me], sum(recv_splits), local_p.dtype)
    else:
        recv = torch.empty(sum(recv_splits), dtype=local_p.dtype,
                           device=device)
        td.all_to_all_single(recv, send, output_split_sizes=recv_splits,
                             input_split_sizes=send_splits)
    # non-finite values sit in the last bucket only, so the sizes of the
    # lower buckets are counts of finite values
    rank_offset = int(counts[:, :me].sum())
    q_bucket, bmin = bh_ranked_fn(recv, rank_offset, n_total)
    mins = _all_gather_flat(bmin.view(1)).view(-1)
    inf = torch.full((1,), float('inf'), dtype=mins.dtype, device=mins.device)
    carry = torch.cat([mins[me + 1:], inf]).min()       # stays on the device
    q_bucket = carry_fn(q_bucket, carry)
    if peer:
        src_off = np.concatenate([[0], np.cumsum(counts[:, me])[:-1]]) * 8
        # bucket ``me`` of source s starts after its lower buckets in s's order
        dst_off = (cap_a + counts[:, :me].sum(axis=1)) * 8
        _peer_copy(q_bucket, src_off, ptrs, dst_off, counts[:, me] * 8)
        fence_peer_writes()
        back = _raw_tensor(ptrs[me] + cap_a * 8, n, local_p.dtype)
    else:
        back = torch.empty_like(send)
        td.all_to_all_single(back, q_bucket, output_split_sizes=send_splits,
                             input_split_sizes=recv_splits)
    return back[pos] if n else torch.empty_like(send)


def _coll_tensor(a, like=None):
    """numpy -> tensor on the device the collectives of this backend use"""
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.cuda() if td.get_backend() == 'nccl' else t


def sharded_size_factor_table(balanced, dist, dist_max, n_bins, norm,
                              kernels=None):
    """Size-factor table (hic3defdr/util/scaling.py:27-149 with equal_bin,
    util/binning.py:4-25) of ONE chromosome whose union pixels are sharded
    over the ranks by row range: ``balanced`` (n_local, R) and ``dist``
    (n_local,) are this rank's pixels, rank order = row order.  Every rank
    returns the same table as ``ops.size_factor_table`` gives on the whole
    chromosome (medians are order statistics, so bit for bit; the sums of the
    scaling norms in a different summation order), or None if the chromosome
    has no pixels at all.

    ``kernels``: the per-stage arithmetic (default hic3defdr_b200.ops, i.e.
    libh3d; the CPU gloo test passes numpy stand-ins)."""
    if kernels is None:
        from hic3defdr_b200 import ops as kernels
    K = kernels
    ws, me = world_size(), rank()
    conditional = 'conditional' in norm
    nb = int(n_bins or 0) if conditional else 0
    n_local, n_reps = int(balanced.shape[0]), int(balanced.shape[1])
    n_groups = K.sf_num_groups(dist_max, nb, norm)
    # 1. chromosome-wide position of every local pixel in (distance, row) order
    if conditional:
        if n_local:
            rank_local, ks_local = K.stable_rank(dist, dist_max + 1)
            cnt_local = np.diff(np.asarray(ks_local.cpu()))
        else:
            rank_local, cnt_local = None, np.zeros(dist_max + 1, np.int64)
    else:
        rank_local, cnt_local = None, np.array([n_local], dtype=np.int64)
    all_cnt = _all_gather_counts(cnt_local)                 # (ws, keys)
    key_start = np.concatenate([[0], np.cumsum(all_cnt.sum(axis=0))]) \
        .astype(np.int64)
    n_total = int(key_start[-1])
    if n_total == 0:
        return None
    my_off = key_start[:-1] + all_cnt[:me].sum(axis=0)      # (keys,)
    cnt_me = all_cnt[me]
    # 2. equal-count bins are ranges of the chromosome-wide order; the local
    # pixels of a bin are a contiguous range of the local distance order too
    gstart = np.asarray(K.sf_group_bounds(
        n_total, dist_max, nb, norm,
        key_start if conditional else None).cpu()).astype(np.int64)
    below = np.clip(gstart[:, None] - my_off[None, :], 0, cnt_me[None, :])
    lb = below.sum(axis=1).astype(np.int64)                 # (n_groups + 1,)
    # 3. ratios in local distance order, sent to the owners of their bins
    values = K.sf_values(balanced, rank_local, norm)        # (R, n_local)
    gk = (np.arange(ws + 1) * n_groups) // ws               # owner k: [gk[k], gk[k+1])
    send_splits = [int(lb[gk[k + 1]] - lb[gk[k]]) for k in range(ws)]
    all_lc = _all_gather_counts(np.diff(lb))                # (ws, n_groups)
    recv_counts = all_lc[:, gk[me]:gk[me + 1]]              # (ws, own)
    recv_splits = [int(c.sum()) for c in recv_counts]
    n_recv = int(sum(recv_splits))
    recv = torch.empty((n_reps, max(n_recv, 1)), dtype=values.dtype,
                       device=values.device)
    works = [td.all_to_all_single(
        recv[r, :n_recv], values[r, :n_local],
        output_split_sizes=recv_splits, input_split_sizes=send_splits,
        async_op=True) for r in range(n_reps)]
    pos = regroup_positions(recv_counts, values.device)     # -> [bin][source]
    for w in works:
        w.wait()
    grouped = torch.empty_like(recv)
    if n_recv:
        grouped.index_copy_(1, pos, recv[:, :n_recv])
    own_start = np.concatenate([[0], np.cumsum(recv_counts.sum(axis=0))]) \
        .astype(np.int64)
    # 4. exact medians (sums) of the owned bins, shared with every rank
    n_own_max = int(np.max(np.diff(gk)))
    red_own = torch.zeros((max(n_own_max, 1), n_reps), dtype=values.dtype,
                          device=values.device)
    n_own = int(gk[me + 1] - gk[me])
    if n_own:
        red, _ = K.sf_group_reduce(grouped, own_start, norm)
        red_own[:n_own] = red
    parts = [torch.empty_like(red_own) for _ in range(ws)]
    td.all_gather(parts, red_own)
    red_all = torch.cat([p[:int(gk[k + 1] - gk[k])]
                         for k, p in enumerate(parts)])
    # 5. bin means and interpolation over distance, redundantly
    return K.sf_table(red_all, gstart, key_start if conditional else None,
                      dist_max, nb, norm)
