"""
Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over
NVLink on the GPU box, gloo in the CPU tests).

How the path shards:
  * prepare_data / lrt: per chromosome, no exchange -> chromosomes are dealt to
    ranks by longest-processing-time on the input size;
  * estimate_disp: dispersion is pooled genome-wide per distance
    (hic3defdr/analysis/analysis.py:169-206) -> ONE all-to-all moves every
    rank's pixels of distance d to the rank owning d (contiguous distance
    ranges balanced by pixel count); the (distance, condition) optimisations
    then run without any collective, and the (D+1, C) result is all-gathered;
  * bh: one global correction (analysis.py:296-303) -> p-values are
    all-gathered (8 B per tested pixel over NVLink), every rank ranks the full
    set and keeps the q-values of its own pixels.
Single-process runs take none of these branches.
"""
import numpy as np
import torch
import torch.distributed as td


def initialized():
    return td.is_available() and td.is_initialized()


def rank():
    return td.get_rank() if initialized() else 0


def world_size():
    return td.get_world_size() if initialized() else 1


def barrier():
    if initialized() and world_size() > 1:
        td.barrier()


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


def distance_ranges(global_counts, n_ranks):
    """Contiguous distance ranges [lo_k, hi_k) with balanced pixel counts.
    Returns an int64 array of n_ranks + 1 boundaries."""
    total = int(global_counts.sum())
    cum = np.concatenate([[0], np.cumsum(global_counts)])
    bounds = [0]
    for k in range(1, n_ranks):
        target = total * k / float(n_ranks)
        b = int(np.searchsorted(cum, target, side='left'))
        b = min(max(b, bounds[-1]), len(global_counts))
        bounds.append(b)
    bounds.append(len(global_counts))
    return np.array(bounds, dtype=np.int64)


def _all_gather_counts(local_counts):
    t = torch.from_numpy(np.ascontiguousarray(local_counts, dtype=np.int64))
    if td.get_backend() == 'nccl':
        t = t.cuda()
    out = [torch.empty_like(t) for _ in range(world_size())]
    td.all_gather(out, t)
    return np.stack([o.cpu().numpy() for o in out])


def _regroup_order(keys):
    """positions that sort ``keys`` stably (receiver-side regrouping)."""
    if keys.is_cuda:
        from hic3defdr_b200 import ops
        n_keys = int(keys.max().item()) + 1 if keys.numel() else 1
        r, _ = ops.stable_rank(keys, n_keys)
        return r.long()
    order = torch.argsort(keys, stable=True)
    r = torch.empty_like(order)
    r[order] = torch.arange(len(order))
    return r


def exchange_by_distance(x, f, seg_start, n_local):
    """x, f: (R, ld) pooled by distance on this rank; seg_start: (D + 2) local
    group boundaries.  Returns (x, f, seg_start, owner) where the arrays hold
    ALL ranks' pixels of the distances this rank owns (other distances are
    empty segments) and ``owner`` is the boundary array (None when single
    process)."""
    if world_size() == 1:
        return x, f, seg_start, None
    ws, me = world_size(), rank()
    n_dist = len(seg_start) - 1
    local_counts = np.diff(seg_start)
    all_counts = _all_gather_counts(local_counts)          # (ws, n_dist)
    bounds = distance_ranges(all_counts.sum(axis=0), ws)
    n_reps = x.shape[0]
    send_splits = [int(seg_start[bounds[k + 1]] - seg_start[bounds[k]])
                   for k in range(ws)]
    lo, hi = int(bounds[me]), int(bounds[me + 1])
    recv_counts = all_counts[:, lo:hi]                      # (ws, owned)
    recv_splits = [int(c.sum()) for c in recv_counts]
    n_recv = int(sum(recv_splits))
    xr = torch.empty((n_reps, max(n_recv, 1)), dtype=x.dtype, device=x.device)
    fr = torch.empty_like(xr)
    for r in range(n_reps):
        for src, dst in ((x, xr), (f, fr)):
            td.all_to_all_single(dst[r, :n_recv], src[r, :n_local].contiguous(),
                                 output_split_sizes=recv_splits,
                                 input_split_sizes=send_splits)
    # received layout: [source rank][distance]; wanted: [distance][source rank]
    keys = torch.from_numpy(np.concatenate(
        [np.repeat(np.arange(hi - lo), c) for c in recv_counts]
        + [np.zeros(0, dtype=np.int64)]).astype(np.int32)).to(x.device)
    pos = _regroup_order(keys)
    xo = torch.empty_like(xr)
    fo = torch.empty_like(fr)
    if n_recv:
        xo[:, pos] = xr[:, :n_recv]
        fo[:, pos] = fr[:, :n_recv]
    owned = recv_counts.sum(axis=0)
    seg = np.zeros(n_dist + 1, dtype=np.int64)
    seg[lo + 1:hi + 1] = np.cumsum(owned)
    seg[hi + 1:] = seg[hi]
    return xo, fo, seg, bounds


def merge_disp_per_dist(disp_local, bounds):
    """Every rank contributes the rows of the distances it owns."""
    if bounds is None:
        return disp_local
    t = torch.from_numpy(np.nan_to_num(disp_local, nan=0.0))
    mask = torch.from_numpy(np.isfinite(disp_local).astype(np.float64))
    lo, hi = int(bounds[rank()]), int(bounds[rank() + 1])
    keep = torch.zeros_like(t)
    keep[lo:hi] = 1
    t, mask = t * keep, mask * keep
    if td.get_backend() == 'nccl':
        t, mask = t.cuda(), mask.cuda()
    td.all_reduce(t)
    td.all_reduce(mask)
    out = t.cpu().numpy()
    out[mask.cpu().numpy() == 0] = np.nan
    return out


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


def global_bh(local_p, bh_fn=None):
    """BH over the p-values of all ranks; returns this rank's q-values."""
    if bh_fn is None:
        from hic3defdr_b200 import ops
        bh_fn = ops.adjust_pvalues
    if world_size() == 1:
        return bh_fn(local_p)
    allp, sizes = all_gather_varlen(local_p)
    q = bh_fn(allp)
    start = int(sizes[:rank()].sum())
    return q[start:start + int(sizes[rank()])]
