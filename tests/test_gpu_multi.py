"""GPU, world_size 2, NCCL: the sharded run (chromosomes dealt to ranks,
dispersion pixels exchanged by distance owner, global BH) against the
single-process run of the same inputs.  Needs two CUDA devices; skipped on a
one-GPU box (the host-side logic is covered on CPU by tests/test_dist_gloo.py).

Tolerance: NONE -- every output of the N-rank run must equal the one-process
run bit for bit.  Inside a distance the pooled pixel order is (rank,
chromosome, row) instead of (chromosome, row), but the likelihood of a bin is
summed in 128-bit fixed point (csrc/disp.cu, Fix128), which does not depend on
the order; medians are order statistics; BH ranks are global."""
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
    np.testing.assert_array_equal(dpds[0], dpd)
    for c in names:
        for k in KEYS:
            np.testing.assert_array_equal(got[c][k], want[c][k], err_msg=k)


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
    np.testing.assert_array_equal(dpds[0], dpd)
    for c in names:
        assert len(parts[0][c]['row']) and len(parts[1][c]['row'])
        got = {k: np.concatenate([p[c][k] for p in parts]) for k in KEYS}
        for k in KEYS:
            np.testing.assert_array_equal(got[k], want[c][k], err_msg=k)


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
    names = ('row', 'col', 'raw', 'disp_idx', 'loop_idx', 'size_factors',
             'scaled', 'disp', 'pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt',
             'qvalues')
    ld = lambda out, k, c: np.load(os.path.join(out, '%s_%s.npy' % (k, c)))
    for c in kw['chroms']:
        for k in names:
            a, b = ld(out1, k, c), ld(out2, k, c)
            assert a.shape == b.shape and a.dtype == b.dtype, k
            np.testing.assert_array_equal(a, b, err_msg=k)
    np.testing.assert_array_equal(
        np.load(os.path.join(out1, 'disp_per_dist.npy')),
        np.load(os.path.join(out2, 'disp_per_dist.npy')))
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
                  'scaled', 'disp', 'pvalues', 'llr', 'mu_hat_null',
                  'mu_hat_alt', 'qvalues'):
            np.testing.assert_array_equal(ld(out1, k, c), ld(out2, k, c),
                                          err_msg=k)
    np.testing.assert_array_equal(
        np.load(os.path.join(out1, 'disp_per_dist.npy')),
        np.load(os.path.join(out2, 'disp_per_dist.npy')))
