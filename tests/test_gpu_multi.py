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
