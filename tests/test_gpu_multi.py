"""GPU, world_size 2, NCCL: the sharded run (chromosomes dealt to ranks,
dispersion pixels exchanged by distance owner, global BH) against the
single-process run of the same inputs.  Needs two CUDA devices; skipped on a
one-GPU box (the host-side logic is covered on CPU by tests/test_dist_gloo.py).

Tolerances: indices / masks / raw bit-exact; disp_per_dist 1e-7 (inside a
distance the pooled pixel order is (rank, chromosome, row) instead of
(chromosome, row): only the summation order of the NLL partials changes);
p / q 1e-6 given that (end-to-end tolerance of SURVEY.md section 8(c))."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CHROMS = {'cA': 900, 'cB': 700, 'cC': 500}
DIST_MAX = 40


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(names):
    from hic3defdr_b200 import ops
    from hic3defdr_b200.synth import make_chrom
    out = []
    for c in names:
        mats, bias, _ = make_chrom(CHROMS[c], 4, DIST_MAX,
                                   seed=4242 + 10 * list(CHROMS).index(c),
                                   amp=250.0)
        out.append((ops.DeviceCSR(mats), ops.dev(bias)))
    return out


KEYS = ('row', 'col', 'raw', 'size_factors', 'scaled', 'disp_idx', 'disp',
        'pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt', 'qvalues')


def _to_host(states, names):
    return {c: {k: st[k].cpu().numpy() for k in KEYS}
            for c, st in zip(names, states)}


def _worker(rank, world, port, outdir):
    import torch
    import torch.distributed as td
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    td.init_process_group('nccl', rank=rank, world_size=world,
                          device_id=torch.device('cuda', rank))
    from hic3defdr_b200 import dist as hd
    from hic3defdr_b200 import engine
    design = np.array([[1, 0], [1, 0], [0, 1], [0, 1]], dtype=bool)
    names = list(CHROMS)
    mine = hd.shard_chroms(names, lambda c: CHROMS[c])
    states, dpd, fns, stats = engine.run_to_qvalues(
        _inputs(mine), design, dist_max=DIST_MAX)
    res = _to_host(states, mine)
    res['__dpd__'] = dpd
    import pickle
    with open(os.path.join(outdir, 'rank%d.pkl' % rank), 'wb') as h:
        pickle.dump(res, h)
    td.barrier()
    td.destroy_process_group()


def test_two_rank_nccl_matches_single_process(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 CUDA devices')
    import pickle
    import torch.multiprocessing as mp
    from hic3defdr_b200 import engine
    outdir = str(tmp_path)
    mp.spawn(_worker, args=(2, _free_port(), outdir), nprocs=2, join=True)
    got = {}
    dpds = []
    for r in range(2):
        with open(os.path.join(outdir, 'rank%d.pkl' % r), 'rb') as h:
            res = pickle.load(h)
        dpds.append(res.pop('__dpd__'))
        got.update(res)
    assert sorted(got) == sorted(CHROMS)
    np.testing.assert_array_equal(dpds[0], dpds[1])     # every rank has the table
    design = np.array([[1, 0], [1, 0], [0, 1], [0, 1]], dtype=bool)
    names = list(CHROMS)
    states, dpd, fns, stats = engine.run_to_qvalues(
        _inputs(names), design, dist_max=DIST_MAX)
    want = _to_host(states, names)
    ok = np.isfinite(dpd)
    assert np.array_equal(ok, np.isfinite(dpds[0]))
    np.testing.assert_allclose(dpds[0][ok], dpd[ok], rtol=1e-7, atol=1e-9)
    for c in names:
        for k in ('row', 'col', 'raw', 'disp_idx'):
            np.testing.assert_array_equal(got[c][k], want[c][k], err_msg=k)
        for k in ('size_factors', 'scaled'):
            np.testing.assert_allclose(got[c][k], want[c][k], rtol=1e-12,
                                       err_msg=k)
        for k in ('disp', 'mu_hat_null', 'mu_hat_alt'):
            np.testing.assert_allclose(got[c][k], want[c][k], rtol=1e-6,
                                       err_msg=k)
        np.testing.assert_allclose(got[c]['llr'], want[c]['llr'], rtol=1e-5,
                                   atol=1e-9)
        np.testing.assert_allclose(got[c]['pvalues'], want[c]['pvalues'],
                                   rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(got[c]['qvalues'], want[c]['qvalues'],
                                   rtol=1e-5, atol=1e-12)


def _sharded_worker(rank, world, port, outdir):
    import torch
    import torch.distributed as td
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    td.init_process_group('nccl', rank=rank, world_size=world,
                          device_id=torch.device('cuda', rank))
    from hic3defdr_b200 import dist as hd
    from hic3defdr_b200 import engine, ops, staging
    from hic3defdr_b200.synth import make_chrom
    design = np.array([[1, 0], [1, 0], [0, 1], [0, 1]], dtype=bool)
    inputs = []
    for c in CHROMS:
        mats, bias, _ = make_chrom(CHROMS[c], 4, DIST_MAX,
                                   seed=4242 + 10 * list(CHROMS).index(c),
                                   amp=250.0)
        bounds = hd.row_ranges(staging.row_weights(mats))
        part = staging.shard_rows(mats, int(bounds[rank]), int(bounds[rank + 1]))
        inputs.append((ops.DeviceCSR(part), ops.dev(bias)))
    states, dpd, fns, stats = engine.run_to_qvalues(
        inputs, design, dist_max=DIST_MAX, row_sharded=True)
    res = _to_host(states, list(CHROMS))
    res['__dpd__'] = dpd
    import pickle
    with open(os.path.join(outdir, 'rank%d.pkl' % rank), 'wb') as h:
        pickle.dump(res, h)
    td.barrier()
    td.destroy_process_group()


def test_two_rank_row_sharded_matches_single_process(tmp_path):
    """Pixel-range sharding (BASELINE config 4): every
    rank takes a row range of EVERY chromosome; size factors through
    dist.sharded_size_factor_table.  Concatenating the ranks' states in rank
    order must give the single-process state: indices / raw / masks and the
    median-based size factors bit-exact, the rest as in the chromosome-sharded
    test above."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 CUDA devices')
    import pickle
    import torch.multiprocessing as mp
    from hic3defdr_b200 import engine
    outdir = str(tmp_path)
    mp.spawn(_sharded_worker, args=(2, _free_port(), outdir), nprocs=2,
             join=True)
    parts, dpds = [], []
    for r in range(2):
        with open(os.path.join(outdir, 'rank%d.pkl' % r), 'rb') as h:
            res = pickle.load(h)
        dpds.append(res.pop('__dpd__'))
        parts.append(res)
    np.testing.assert_array_equal(dpds[0], dpds[1])
    design = np.array([[1, 0], [1, 0], [0, 1], [0, 1]], dtype=bool)
    names = list(CHROMS)
    states, dpd, fns, stats = engine.run_to_qvalues(
        _inputs(names), design, dist_max=DIST_MAX)
    want = _to_host(states, names)
    ok = np.isfinite(dpd)
    assert np.array_equal(ok, np.isfinite(dpds[0]))
    np.testing.assert_allclose(dpds[0][ok], dpd[ok], rtol=1e-7, atol=1e-9)
    # the trend fit amplifies 1e-7 differences of disp_per_dist (DESIGN.md
    # section 4, "Trend fit sensitivity"; the pooled pixel order inside a
    # distance differs between the two runs), so everything after it is
    # compared stage-isolated: the single-process stages fed with the sharded
    # run's disp_per_dist must reproduce the sharded run's outputs exactly
    import torch
    from hic3defdr_b200 import ops
    _, table = engine.fit_trends(dpds[0], DIST_MAX, ['0', '1'])
    for st in states:
        idx = st['disp_index'].long()
        st['disp'] = ops.gather_table(
            (st['col'][idx] - st['row'][idx]).to(torch.int32), table)
        engine.lrt_chrom(st, design)
    engine.bh(states)
    iso = _to_host(states, names)
    for c in names:
        assert len(parts[0][c]['row']) and len(parts[1][c]['row'])
        got = {k: np.concatenate([p[c][k] for p in parts]) for k in KEYS}
        for k in ('row', 'col', 'raw', 'disp_idx', 'size_factors', 'scaled'):
            np.testing.assert_array_equal(got[k], want[c][k], err_msg=k)
        for k in ('disp', 'mu_hat_null', 'mu_hat_alt', 'llr', 'pvalues'):
            np.testing.assert_array_equal(got[k], iso[c][k], err_msg=k)
        np.testing.assert_allclose(got['qvalues'], iso[c]['qvalues'],
                                   rtol=1e-12, atol=0)


def _class_worker(rank, world, port, kw, outdir, shard='rows'):
    import torch
    import torch.distributed as td
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    os.environ['H3D_SHARD'] = shard
    torch.cuda.set_device(rank)
    td.init_process_group('nccl', rank=rank, world_size=world,
                          device_id=torch.device('cuda', rank))
    from hic3defdr_b200 import HiC3DeFDR
    h = HiC3DeFDR(outdir=outdir, dist_thresh_max=DIST_MAX, **kw)
    h.run_to_qvalues(n_threads=0)
    # checkpoint / resume: a fresh object redoes lrt + bh from the files alone
    # (its pixel partition is the even split, not the row ranges)
    h2 = HiC3DeFDR.load(outdir)
    h2.lrt(n_threads=0)
    h2.bh()
    td.barrier()
    td.destroy_process_group()


def test_two_rank_row_sharded_class_writes_the_same_files(tmp_path):
    """The drop-in class with H3D_SHARD=rows on 2 ranks writes, slice by
    slice, the same per-chromosome .npy files as one process."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 CUDA devices')
    import torch.multiprocessing as mp
    from hic3defdr_b200 import HiC3DeFDR
    from hic3defdr_b200.synth import write_dataset
    root = str(tmp_path)
    kw = write_dataset(os.path.join(root, 'in'), {'cA': 700, 'cB': 500},
                       n_reps=4, dist_max=DIST_MAX, amp=250.0, loops=True)
    out1, out2 = os.path.join(root, 'one'), os.path.join(root, 'two')
    h = HiC3DeFDR(outdir=out1, dist_thresh_max=DIST_MAX, **kw)
    h.run_to_qvalues(n_threads=0)
    mp.spawn(_class_worker, args=(2, _free_port(), kw, out2), nprocs=2,
             join=True)
    exact = ('row', 'col', 'raw', 'disp_idx', 'loop_idx', 'size_factors',
             'scaled')
    later = ('disp', 'pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt', 'qvalues')
    ld = lambda out, k, c: np.load(os.path.join(out, '%s_%s.npy' % (k, c)))
    for c in kw['chroms']:
        for k in exact + later:
            a, b = ld(out1, k, c), ld(out2, k, c)
            assert a.shape == b.shape and a.dtype == b.dtype, k
            if k in exact:
                np.testing.assert_array_equal(a, b, err_msg=k)
    a = np.load(os.path.join(out1, 'disp_per_dist.npy'))
    b = np.load(os.path.join(out2, 'disp_per_dist.npy'))
    ok = np.isfinite(a)
    assert np.array_equal(ok, np.isfinite(b))
    # ~600-pixel bins: the pixel order inside a distance differs between the
    # runs, the NLL partial sums with it, and Brent (xatol 1e-5 on delta) lands
    # up to ~1e-8 apart -- the reference's own order sensitivity, DESIGN.md 5
    np.testing.assert_allclose(a[ok], b[ok], rtol=1e-6, atol=1e-7)
    # after the (ill-conditioned) trend fit: stage-isolated, as in the test
    # above -- one process redoes lrt + bh from the 2-rank run's disp files
    want = {(k, c): ld(out2, k, c) for c in kw['chroms'] for k in later}
    h3 = HiC3DeFDR.load(out2)
    h3.lrt(n_threads=0)
    h3.bh()
    for (k, c), b in want.items():
        a = ld(out2, k, c)
        if k == 'qvalues':
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=0, err_msg=k)
        else:
            np.testing.assert_array_equal(a, b, err_msg=k)
    # disp is one value per (distance, condition), the same on both ranks
    for c in kw['chroms']:
        di = ld(out2, 'disp_idx', c)
        d = (ld(out2, 'col', c) - ld(out2, 'row', c))[di]
        disp = ld(out2, 'disp', c)
        for dd in np.unique(d)[:10]:
            assert len(np.unique(disp[d == dd], axis=0)) == 1


def test_two_rank_chromosome_sharded_class_writes_the_same_files(tmp_path):
    """The drop-in class in its default multi-GPU mode (whole chromosomes per
    rank, background writers, one barrier at the end of run_to_qvalues): the
    files of both ranks together are the single-process files."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 CUDA devices')
    import torch.multiprocessing as mp
    from hic3defdr_b200 import HiC3DeFDR
    from hic3defdr_b200.synth import write_dataset
    root = str(tmp_path)
    kw = write_dataset(os.path.join(root, 'in'),
                       {'cA': 700, 'cB': 500, 'cC': 400}, n_reps=4,
                       dist_max=DIST_MAX, amp=250.0, loops=True)
    out1, out2 = os.path.join(root, 'one'), os.path.join(root, 'two')
    h = HiC3DeFDR(outdir=out1, dist_thresh_max=DIST_MAX, **kw)
    h.run_to_qvalues(n_threads=0)
    mp.spawn(_class_worker, args=(2, _free_port(), kw, out2, 'chroms'),
             nprocs=2, join=True)
    ld = lambda out, k, c: np.load(os.path.join(out, '%s_%s.npy' % (k, c)))
    for c in kw['chroms']:
        for k in ('row', 'col', 'raw', 'disp_idx', 'loop_idx', 'size_factors',
                  'scaled'):
            np.testing.assert_array_equal(ld(out1, k, c), ld(out2, k, c),
                                          err_msg=k)
        for k in ('disp', 'pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt',
                  'qvalues'):
            assert ld(out1, k, c).shape == ld(out2, k, c).shape, k
    a = np.load(os.path.join(out1, 'disp_per_dist.npy'))
    b = np.load(os.path.join(out2, 'disp_per_dist.npy'))
    ok = np.isfinite(a)
    assert np.array_equal(ok, np.isfinite(b))
    np.testing.assert_allclose(a[ok], b[ok], rtol=1e-6, atol=1e-7)
    # stage-isolated after the trend fit, as above
    later = ('pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt', 'qvalues')
    want = {(k, c): ld(out2, k, c) for c in kw['chroms'] for k in later}
    h3 = HiC3DeFDR.load(out2)
    h3.lrt(n_threads=0)
    h3.bh()
    for (k, c), b in want.items():
        if k == 'qvalues':
            np.testing.assert_allclose(ld(out2, k, c), b, rtol=1e-12, atol=0)
        else:
            np.testing.assert_array_equal(ld(out2, k, c), b, err_msg=k)
