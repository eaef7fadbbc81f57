"""CPU, property-based (hypothesis): the host-side partitioning helpers of the
multi-GPU path (hic3defdr_b200/dist.py) for arbitrary
counts -- empty ranks, empty distances, one rank, more ranks than items."""
import numpy as np
from hypothesis import given, settings, strategies as st

from hic3defdr_b200 import dist as hd


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(0, 10 ** 6), min_size=0, max_size=40),
       st.integers(1, 9))
def test_lpt_assign_is_a_balanced_partition(weights, n_ranks):
    owner = hd.lpt_assign(weights, n_ranks)
    assert len(owner) == len(weights)
    assert all(0 <= o < n_ranks for o in owner)
    loads = [sum(w for w, o in zip(weights, owner) if o == k)
             for k in range(n_ranks)]
    assert sum(loads) == sum(weights)
    # greedy on sorted weights: no rank exceeds another by more than one item
    if weights:
        assert max(loads) - min(loads) <= max(weights)
    # deterministic (every rank computes the same deal)
    assert owner == hd.lpt_assign(list(weights), n_ranks)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(0, 5000), min_size=0, max_size=300),
       st.integers(1, 9))
def test_row_ranges_cover_the_rows_in_order(weights, n_ranks):
    b = hd.row_ranges(weights, n_ranks)
    n = len(weights)
    assert b.dtype == np.int64 and len(b) == n_ranks + 1
    assert b[0] == 0 and b[-1] == n and (np.diff(b) >= 0).all()
    w = np.asarray(weights, dtype=np.int64)
    total = int(w.sum())
    if total:
        shares = [int(w[b[k]:b[k + 1]].sum()) for k in range(n_ranks)]
        assert sum(shares) == total
        # a cut lands on the row where the cumulative weight crosses k/n of
        # the total: no share exceeds the ideal one by more than a row
        assert max(shares) <= total / n_ranks + w.max()


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 6), st.integers(1, 7), st.data())
def test_owner_layout_tiles_the_owner_buffers(ws, per, data):
    counts = np.array(data.draw(st.lists(
        st.lists(st.integers(0, 30), min_size=ws * per, max_size=ws * per),
        min_size=ws, max_size=ws)), dtype=np.int64)
    lays = [hd.owner_layout(counts, per, me) for me in range(ws)]
    for k in range(ws):
        n_recv, _, (run_seg, run_lo, run_hi) = lays[k]
        assert int(n_recv[k]) == int(counts[:, k * per:(k + 1) * per].sum())
        cover = np.zeros(int(n_recv[k]), dtype=int)
        # what every source writes into owner k's buffer
        for me in range(ws):
            start = np.cumsum(counts[me]) - counts[me]
            for j in range(per):
                key = k * per + j
                lo = start[key] + lays[me][1][key]
                cover[lo:lo + counts[me, key]] += 1
        assert (cover == 1).all()
        # the runs the owner reads: ordered by key, then source, same tiling
        assert (np.diff(run_seg) >= 0).all()
        seen = np.zeros_like(cover)
        for s, lo, hi in zip(run_seg, run_lo, run_hi):
            seen[lo:hi] += 1
        assert (seen == 1).all()
        for j in range(per):
            sel = run_seg == j
            assert int((run_hi[sel] - run_lo[sel]).sum()) == \
                int(counts[:, k * per + j].sum())


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 6), st.integers(1, 40), st.data())
def test_lpt_layout_segments_are_contiguous(ws, n_dist, data):
    counts = np.array(data.draw(st.lists(
        st.lists(st.integers(0, 50), min_size=n_dist, max_size=n_dist),
        min_size=ws, max_size=ws)), dtype=np.int64)
    lays = [hd.lpt_layout(counts, me) for me in range(ws)]
    owner = lays[0]['owner']
    assert sorted(np.concatenate(lays[0]['owned']).tolist()) == \
        list(range(n_dist))
    for k in range(ws):
        d_own = lays[k]['owned'][k]
        seg = lays[k]['seg_start']
        assert len(seg) == len(d_own) + 1
        cover = np.zeros(int(seg[-1]), dtype=int)
        for me in range(ws):
            assert np.array_equal(lays[me]['owner'], owner)
            start = np.cumsum(counts[me]) - counts[me]
            for j, d in enumerate(d_own):
                lo = start[d] + lays[me]['shift'][d]
                assert seg[j] <= lo and lo + counts[me, d] <= seg[j + 1]
                cover[lo:lo + counts[me, d]] += 1
        assert (cover == 1).all()
