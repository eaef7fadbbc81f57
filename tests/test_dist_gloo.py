"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path
(hic3defdr_b200/dist.py): chromosome sharding, the distance all-to-all with
receiver-side regrouping, the merge of per-distance results and the global
BH gather.  The CUDA kernels are not involved (the arithmetic callbacks are
replaced by numpy stand-ins where the product would call libh3d)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    td.init_process_group('gloo', rank=rank, world_size=world)
    from hic3defdr_b200 import dist as hd
    from oracle import pipeline as op
    try:
        # 1. sharding is a partition, identical on every rank
        chroms = ['c%d' % i for i in range(7)]
        w = {c: (i * 37) % 11 + 1 for i, c in enumerate(chroms)}
        mine = hd.shard_chroms(chroms, lambda c: w[c])
        gathered = [None] * world
        td.all_gather_object(gathered, mine)
        assert sorted(sum(gathered, [])) == sorted(chroms)
        # 2. exchange by distance: every pixel ends on the owner of its
        # distance (round robin), grouped by distance, source ranks in order
        rng = np.random.default_rng(100 + rank)
        n_dist, n_reps, n = 13, 3, 200 + 50 * rank
        key_of_dist, per = hd.distance_keys(n_dist)
        assert sorted(key_of_dist) == sorted(set(key_of_dist))
        dist = rng.integers(2, n_dist, size=n)
        dist = dist[np.argsort(key_of_dist[dist], kind='stable')]   # pooled
        seg = np.concatenate([[0], np.cumsum(np.bincount(
            key_of_dist[dist], minlength=per * world))])
        x = np.stack([dist * 1000.0 + rank * 100 + r for r in range(n_reps)])
        ident = rng.random(n)                       # per-pixel tag
        f = np.stack([ident + r for r in range(n_reps)])
        xo, fo, runs = hd.exchange_by_distance(
            torch.from_numpy(x), torch.from_numpy(f), seg, n)
        mine_d = hd.owned_distances(n_dist)
        run_seg, run_lo, run_hi = runs
        assert len(run_seg) == per * world
        assert (np.diff(run_seg) >= 0).all()        # by owned key, then source
        n_got = int((run_hi - run_lo).sum())
        xo, fo = xo.numpy()[:, :n_got], fo.numpy()[:, :n_got]
        seg_o = np.concatenate([[0], np.cumsum(np.bincount(
            run_seg, weights=run_hi - run_lo, minlength=per))]).astype(int)
        covered = np.zeros(n_got, dtype=int)
        for j in range(per):
            last_src = -1
            for g, a, b in zip(run_seg, run_lo, run_hi):
                if g != j or a == b:
                    continue
                assert j < len(mine_d)
                covered[a:b] += 1
                assert (xo[0, a:b] // 1000 == mine_d[j]).all()
                src = (xo[0, a:b] % 1000) // 100
                assert len(set(src.tolist())) == 1 and src[0] > last_src
                last_src = src[0]                   # one run per source, in rank order
        assert (covered == 1).all()
        # the peer-memory exchange (csrc/peer.cu) derives the same layout from
        # the all-gathered counts alone: identical runs, and every local pixel
        # lands inside its run of the owner's buffer
        allc = hd._all_gather_counts(np.diff(seg))
        n_recv, shift, runs2 = hd.owner_layout(allc, per, rank)
        for a, b in zip(runs, runs2):
            assert np.array_equal(a, b)
        assert int(n_recv[rank]) == n_got
        layouts = [None] * world
        td.all_gather_object(layouts, (shift.tolist(), seg.tolist()))
        for src, (sh, sg) in enumerate(layouts):
            for key in range(per * world):
                if key // per != rank or sg[key + 1] == sg[key]:
                    continue
                j = key - rank * per
                sel = [(a, b) for g, a, b in zip(*runs2) if g == j]
                lo_run, hi_run = sel[src]           # one run per source, in order
                assert sg[key] + sh[key] == lo_run
                assert sg[key + 1] + sh[key] == hi_run
        # every pixel arrived exactly once, on exactly one rank
        tags = [None] * world
        td.all_gather_object(tags, fo[0].tolist())
        sent = [None] * world
        td.all_gather_object(sent, ident.tolist())
        assert sorted(sum(tags, [])) == sorted(sum(sent, []))
        # 3. merge of per-distance dispersions
        local = np.full((per, 2), np.nan)
        local[:len(mine_d)] = mine_d[:, None] + np.array([0.0, 0.5])
        local[np.diff(seg_o) == 0] = np.nan
        merged = hd.merge_disp_per_dist(local, n_dist)
        allc = [None] * world
        td.all_gather_object(allc, np.bincount(dist, minlength=n_dist).tolist())
        counts = np.sum(allc, axis=0)
        assert merged.shape == (n_dist, 2)
        assert np.array_equal(np.isnan(merged[:, 0]), counts == 0)
        ok = counts > 0
        assert np.array_equal(merged[ok, 1], np.arange(n_dist)[ok] + 0.5)
        # 4. global BH equals BH over the concatenation
        # (distributed sort/rank; ties across the splitter and NaN included)
        p_local = rng.random(3000 + 170 * rank) ** 2
        p_local[::7] = 0.25                 # heavy ties, the same on all ranks
        p_local[5::31] = np.nan
        p_local[3] = 1.0

        def ranked(t, off, ntot):
            p = t.numpy()
            q = np.full(p.shape, np.nan)
            fin = np.isfinite(p)
            order = np.argsort(p[fin], kind='stable')
            raw = p[fin][order] / ((off + np.arange(1, fin.sum() + 1))
                                   / float(ntot))
            sm = np.minimum.accumulate(raw[::-1])[::-1]
            sm[sm > 1] = 1
            out = np.empty(len(sm))
            out[order] = sm
            q[fin] = out
            return torch.from_numpy(q), torch.tensor(
                [raw.min() if len(raw) else np.inf])

        q_local = hd.global_bh(
            torch.from_numpy(p_local),
            bh_fn=lambda t: torch.from_numpy(op.bh(t.numpy())),
            bh_ranked_fn=ranked,
            carry_fn=lambda q, c: torch.where(q > c, c.to(q.dtype).expand_as(q), q))
        allp = [None] * world
        td.all_gather_object(allp, p_local.tolist())
        q_all = op.bh(np.concatenate([np.array(v) for v in allp]))
        start = sum(len(v) for v in allp[:rank])
        assert np.array_equal(q_local.numpy(), q_all[start:start + len(p_local)],
                              equal_nan=True)
        ret[rank] = 'ok'
    except Exception as e:          # surface the failure in the parent
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        td.destroy_process_group()


class _NumpyStages(object):
    """numpy stand-ins for the libh3d stages that
    ``dist.sharded_size_factor_table`` chains (same layouts and conventions as
    hic3defdr_b200/ops.py: sf_num_groups ... sf_table)."""

    @staticmethod
    def _conditional(norm):
        return 'conditional' in norm

    def sf_num_groups(self, dist_max, n_bins, norm):
        if not self._conditional(norm):
            return 1
        return n_bins if n_bins else dist_max + 1

    def stable_rank(self, keys, n_keys):
        k = np.asarray(keys)
        order = np.argsort(k, kind='stable')
        rank = np.empty(len(k), dtype=np.int64)
        rank[order] = np.arange(len(k))
        start = np.concatenate([[0], np.cumsum(np.bincount(k, minlength=n_keys))])
        return torch.from_numpy(rank), torch.from_numpy(start)

    def sf_group_bounds(self, n_total, dist_max, n_bins, norm, key_start):
        if not self._conditional(norm):
            return torch.tensor([0, n_total])
        if not n_bins:
            return torch.from_numpy(np.asarray(key_start[:dist_max + 2]))
        idx = np.linspace(0, n_bins, n_total, endpoint=0, dtype=int)
        return torch.from_numpy(np.searchsorted(idx, np.arange(n_bins + 1)))

    def sf_values(self, balanced, rank, norm):
        b = np.asarray(balanced)
        out = np.empty(b.shape[::-1])
        if norm in ('conditional_mor', 'median_of_ratios'):
            gm = np.exp(np.mean(np.log(b + 1), axis=1)) - 1
            with np.errstate(all='ignore'):
                v = b / gm[:, None]
            v[~np.all(b > 0, axis=1)] = np.nan
        else:
            v = b
        pos = np.arange(len(b)) if rank is None else np.asarray(rank)
        out[:, pos] = v.T
        return torch.from_numpy(out)

    def sf_group_reduce(self, values, gstart, norm):
        v = values.numpy()
        g = np.asarray(gstart)
        red = np.full((len(g) - 1, v.shape[0]), np.nan)
        for i in range(len(g) - 1):
            seg = v[:, g[i]:g[i + 1]]
            if norm in ('conditional_mor', 'median_of_ratios'):
                ok = ~np.isnan(seg[0])
                if ok.any():
                    red[i] = np.median(seg[:, ok], axis=1)
            else:
                red[i] = seg.sum(axis=1)
        return torch.from_numpy(red), None

    def sf_table(self, red, gstart, key_start, dist_max, n_bins, norm):
        from oracle import pipeline as op
        from oracle.thirdparty import gmean
        red, g = red.numpy(), np.asarray(gstart)
        occ = np.diff(g) > 0
        s_b = red[occ]
        if 'scaling' in norm:
            s_b = s_b / np.array([gmean(row) for row in s_b])[:, None]
        if not self._conditional(norm):
            return torch.from_numpy(s_b[0])
        d_of = np.repeat(np.arange(dist_max + 1), np.diff(key_start))
        d_b = np.array([d_of[g[i]:g[i + 1]].mean()
                        for i in range(len(g) - 1) if occ[i]])
        x = np.arange(dist_max + 1)
        if not n_bins:
            table = np.full((dist_max + 1, red.shape[1]), np.nan)
            table[occ] = s_b
            return torch.from_numpy(table)
        return torch.from_numpy(np.stack(
            [op.interp_extrap(d_b, s_b[:, r], x)
             for r in range(red.shape[1])], axis=1))


def _sf_worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    td.init_process_group('gloo', rank=rank, world_size=world)
    from hic3defdr_b200 import dist as hd
    from oracle import pipeline as op
    try:
        # one chromosome: pixels in (row, col) order, cut by row range
        rng = np.random.default_rng(77)                 # same on every rank
        n_rows, dist_max, n_reps = 400, 30, 4
        row = np.repeat(np.arange(n_rows), dist_max + 1)
        col = row + np.tile(np.arange(dist_max + 1), n_rows)
        keep = (col < n_rows) & (rng.random(len(row)) < 0.8)
        row, col = row[keep], col[keep]
        dist = (col - row).astype(np.int32)
        bal = rng.gamma(2.0, 40.0 / (1 + dist)[:, None], (len(row), n_reps))
        bal[rng.random(bal.shape) < 0.05] = 0.0          # pixels without a ratio
        w = np.bincount(row, minlength=n_rows)
        bounds = hd.row_ranges(w)
        assert bounds[0] == 0 and bounds[-1] == n_rows and \
            (np.diff(bounds) >= 0).all()
        if world == 3:           # a rank without a single pixel must still take part
            bounds = np.array([0, 150, 150, n_rows])
        mine = (row >= bounds[rank]) & (row < bounds[rank + 1])
        for norm, n_bins in (('conditional_mor', 6), ('conditional_mor', 0),
                             ('conditional_scaling', 5),
                             ('median_of_ratios', 0), ('simple_scaling', 0)):
            table = hd.sharded_size_factor_table(
                torch.from_numpy(bal[mine]), torch.from_numpy(dist[mine]),
                dist_max, n_bins, norm, kernels=_NumpyStages()).numpy()
            if norm == 'median_of_ratios':
                np.testing.assert_array_equal(table, op.median_of_ratios(bal))
                continue
            if norm == 'simple_scaling':
                np.testing.assert_allclose(table, op.simple_scaling(bal),
                                           rtol=1e-12)
                continue
            reducer = op.median_of_ratios if norm.endswith('mor') else \
                op.simple_scaling
            want = op.conditional_size_factors(bal, dist, n_bins, reducer)
            got = table[dist]
            if norm.endswith('mor'):
                np.testing.assert_array_equal(got, want)
            else:
                np.testing.assert_allclose(got, want, rtol=1e-12)
        # a chromosome without any pixel
        assert hd.sharded_size_factor_table(
            torch.empty((0, n_reps), dtype=torch.float64),
            torch.empty(0, dtype=torch.int32), dist_max, 6,
            'conditional_mor', kernels=_NumpyStages()) is None
        ret[rank] = 'ok'
    except Exception:
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_row_sharded_size_factors(world):
    """SURVEY.md section 8(e), pixel-range sharding of one chromosome: the
    collectives of ``dist.sharded_size_factor_table`` (counts all-gather, bin
    owner all-to-all, medians all-gather) reproduce the whole-chromosome size
    factors of the oracle, bit for bit for the median-based norms."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sf_worker, args=(world, _free_port(), ret), nprocs=world,
             join=True)
    for r in range(world):
        assert ret.get(r) == 'ok', ret.get(r)


@pytest.mark.parametrize('world', [2, 3])
def test_multi_rank_host_logic(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == 'ok', ret.get(r)


def test_lpt_and_ranges():
    from hic3defdr_b200 import dist as hd
    owner = hd.lpt_assign([10, 9, 8, 7, 1, 1], 2)
    loads = [sum(w for w, o in zip([10, 9, 8, 7, 1, 1], owner) if o == k)
             for k in range(2)]
    assert abs(loads[0] - loads[1]) <= 2
    keys, per = hd.distance_keys(201, 8)
    assert per == 26 and len(set(keys)) == 201 and keys.max() < 8 * per
    owned = [hd.owned_distances(201, k, 8) for k in range(8)]
    assert sorted(np.concatenate(owned).tolist()) == list(range(201))
    for k, d in enumerate(owned):
        # one distance of every group of 8, dealt 0..7, 7..0, 0..7, ...
        assert (d // 8 == np.arange(len(d))).all()
        assert (np.where((d // 8) % 2 == 0, d % 8, 7 - d % 8) == k).all()
        assert (keys[d] == k * per + np.arange(len(d))).all()
    # the boustrophedon deal balances a cost that falls with the distance
    cost = 1000.0 - 3.0 * np.arange(201)
    loads = [cost[d].sum() for d in owned]
    assert (max(loads) - min(loads)) / np.mean(loads) < 0.05
    keys, per = hd.distance_keys(7, 1)
    assert per == 7 and (keys == np.arange(7)).all()


def test_bh_sample_positions_stay_in_range():
    """regression: float32 linspace(0, n - 1, take) indexed one past the end
    for n > 2^24 (device-side assert in the 2-GPU row-sharded mouse run)"""
    from hic3defdr_b200 import dist as hd
    for n in (1, 2, 4095, 4096, 4097, 2 ** 24 + 2, 19_366_195, 38_732_388,
              2 ** 31 - 1):
        take = min(hd.BH_SAMPLES, n)
        sel = hd.sample_positions(n, take).numpy()
        assert sel[0] == 0 and sel[-1] == n - 1
        assert (np.diff(sel) >= 0).all() and len(sel) == take


def test_lpt_layout_tiles_every_owner_buffer():
    """dist.lpt_layout (ownership chosen after the counts are known): the
    slices every source rank writes tile every owner's buffer exactly, every
    owned distance is one contiguous segment (sources in rank order), empty
    distances weigh nothing, and the deal balances the pixel counts."""
    from hic3defdr_b200 import dist as hd
    rng = np.random.default_rng(3)
    ws, n_dist = 5, 61
    counts = rng.integers(0, 400, size=(ws, n_dist)) * \
        np.linspace(2.0, 0.5, n_dist).astype(int).clip(1)
    counts[:, :4] = 0                          # below dist_thresh_min: no pixels
    lays = [hd.lpt_layout(counts, me) for me in range(ws)]
    owner = lays[0]['owner']
    for lay in lays:
        assert np.array_equal(lay['owner'], owner)
    for k in range(ws):
        d_own = lays[k]['owned'][k]
        assert (np.diff(d_own) > 0).all() and (owner[d_own] == k).all()
        seg = lays[k]['seg_start']
        assert np.array_equal(np.diff(seg), counts[:, d_own].sum(axis=0))
        cover = np.zeros(int(lays[k]['n_recv'][k]), dtype=int)
        for me in range(ws):
            local_start = np.cumsum(counts[me]) - counts[me]
            last_hi = {}
            for j, d in enumerate(d_own):
                lo = local_start[d] + lays[me]['shift'][d]
                hi = lo + counts[me, d]
                assert seg[j] <= lo and hi <= seg[j + 1]
                cover[lo:hi] += 1
        assert (cover == 1).all()
    loads = np.array([counts[:, lays[0]['owned'][k]].sum() for k in range(ws)])
    assert (loads.max() - loads.min()) / loads.mean() < 0.05
