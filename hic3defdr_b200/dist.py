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


def distance_keys(n_dist, n_ranks=None):
    """Pooling key of every distance: distances are dealt to ranks round robin
    (owner(d) = d mod n_ranks -- the cost of a (distance, condition) bin varies
    smoothly with d, so interleaving balances pixels AND qCML iterations), and
    the key orders a rank's pooled pixels by (owner, distance):
    key(d) = owner(d) * per + d // n_ranks, per = ceil(n_dist / n_ranks).
    Returns (key per distance int32, per).  Identity for one process."""
    ws = world_size() if n_ranks is None else n_ranks
    per = -(-n_dist // ws)
    d = np.arange(n_dist)
    return ((d % ws) * per + d // ws).astype(np.int32), per


def owned_distances(n_dist, me=None, n_ranks=None):
    """distances owned by rank ``me``, in key order"""
    ws = world_size() if n_ranks is None else n_ranks
    me = rank() if me is None else me
    return np.arange(me, n_dist, ws)


def exchange_by_distance(x, f, seg_start, n_local):
    """x, f: (R, ld) pooled on this rank in ``distance_keys`` order;
    seg_start: (ws * per + 1) local group boundaries over the keys.  Returns
    (x, f, seg_start) holding ALL ranks' pixels of the distances this rank
    owns, one segment per owned key (``per`` segments, the trailing ones empty
    when the distances do not divide evenly)."""
    ws, me = world_size(), rank()
    if ws == 1:
        return x, f, seg_start
    per = (len(seg_start) - 1) // ws
    local_counts = np.diff(seg_start)
    all_counts = _all_gather_counts(local_counts)          # (ws, ws * per)
    n_reps = x.shape[0]
    send_splits = [int(seg_start[(k + 1) * per] - seg_start[k * per])
                   for k in range(ws)]
    recv_counts = all_counts[:, me * per:(me + 1) * per]    # (ws, per)
    recv_splits = [int(c.sum()) for c in recv_counts]
    n_recv = int(sum(recv_splits))
    xr = torch.empty((n_reps, max(n_recv, 1)), dtype=x.dtype, device=x.device)
    fr = torch.empty_like(xr)
    works = []
    for r in range(n_reps):
        for src, dst in ((x, xr), (f, fr)):
            works.append(td.all_to_all_single(
                dst[r, :n_recv], src[r, :n_local],
                output_split_sizes=recv_splits, input_split_sizes=send_splits,
                async_op=True))
    pos = regroup_positions(recv_counts, x.device)
    for w in works:
        w.wait()
    xo = torch.empty_like(xr)
    fo = torch.empty_like(fr)
    if n_recv:                  # pos is a permutation: a plain indexed copy
        xo.index_copy_(1, pos, xr[:, :n_recv])
        fo.index_copy_(1, pos, fr[:, :n_recv])
    seg = np.concatenate([[0], np.cumsum(recv_counts.sum(axis=0))]) \
        .astype(np.int64)
    return xo, fo, seg


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
    bucket-major order, bucket sizes as a host int64 array)."""
    if bucket.is_cuda:
        from hic3defdr_b200 import ops
        pos, start = ops.stable_rank(bucket, n_buckets)
        return pos.long(), np.diff(start.cpu().numpy())
    order = torch.argsort(bucket, stable=True)
    pos = torch.empty_like(order)
    pos[order] = torch.arange(len(order))
    return pos, np.bincount(bucket.numpy(), minlength=n_buckets).astype(np.int64)


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
    the same bucket, so the result is identical to the single-process one."""
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
    # splitters: evenly strided local sample of the finite values
    fin_vals = local_p[finite]
    n_fin = fin_vals.numel()
    sample = torch.full((BH_SAMPLES,), float('inf'), dtype=local_p.dtype,
                        device=device)
    if n_fin:
        take = min(BH_SAMPLES, n_fin)
        sel = sample_positions(n_fin, take, device)
        sample[:take] = fin_vals[sel]
    gathered = [torch.empty_like(sample) for _ in range(ws)]
    td.all_gather(gathered, sample)
    allsamp = torch.sort(torch.cat(gathered)).values
    n_samp = int(torch.isfinite(allsamp).sum().item())
    cut = [min(max(n_samp * k // ws, 0), max(n_samp - 1, 0))
           for k in range(1, ws)]
    splitters = allsamp[torch.tensor(cut, dtype=torch.long, device=device)]
    # bucket of every local value (non-finite: last bucket), stable partition
    bucket = torch.bucketize(local_p, splitters).to(torch.int32)
    bucket = torch.where(finite, bucket, torch.full_like(bucket, ws - 1))
    pos, sizes = _partition(bucket, ws)
    send = torch.empty_like(local_p)
    send[pos] = local_p
    info = np.concatenate([sizes, [n_fin]])
    table = _all_gather_counts(info)                 # (ws, ws + 1)
    counts, n_total = table[:, :ws], int(table[:, ws].sum())
    send_splits = [int(v) for v in counts[me]]
    recv_splits = [int(v) for v in counts[:, me]]
    recv = torch.empty(sum(recv_splits), dtype=local_p.dtype, device=device)
    td.all_to_all_single(recv, send, output_split_sizes=recv_splits,
                         input_split_sizes=send_splits)
    # non-finite values sit in the last bucket only, so the sizes of the
    # lower buckets are counts of finite values
    rank_offset = int(counts[:, :me].sum())
    q_bucket, bmin = bh_ranked_fn(recv, rank_offset, n_total)
    higher = _all_gather_flat(bmin).cpu().numpy().ravel()[me + 1:]
    carry = float(higher.min()) if len(higher) else float('inf')
    q_bucket = carry_fn(q_bucket, carry)
    back = torch.empty_like(send)
    td.all_to_all_single(back, q_bucket, output_split_sizes=send_splits,
                         input_split_sizes=recv_splits)
    return back[pos] if n else back


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
