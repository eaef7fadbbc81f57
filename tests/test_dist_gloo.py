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
        xo, fo, seg_o = hd.exchange_by_distance(
            torch.from_numpy(x), torch.from_numpy(f), seg, n)
        mine_d = hd.owned_distances(n_dist)
        assert len(seg_o) == per + 1
        n_got = int(seg_o[-1])
        xo, fo = xo.numpy()[:, :n_got], fo.numpy()[:, :n_got]
        for j in range(per):
            a, b = int(seg_o[j]), int(seg_o[j + 1])
            if j >= len(mine_d):
                assert a == b
                continue
            assert (xo[0, a:b] // 1000 == mine_d[j]).all()
            src = (xo[0, a:b] % 1000) // 100
            assert (np.diff(src) >= 0).all()        # rank order inside a distance
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
            carry_fn=lambda q, c: torch.where(q > c, torch.full_like(q, c), q))
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
    for k in range(8):
        d = hd.owned_distances(201, k, 8)
        assert (d % 8 == k).all() and (keys[d] == k * per + np.arange(len(d))).all()
    keys, per = hd.distance_keys(7, 1)
    assert per == 7 and (keys == np.arange(7)).all()
