"""GPU: the host-buffer (end-to-end) path.  ``staging.OutputDrain`` in its
compact mode (size_factors / disp as per-distance tables, raw as int32, rebuilt
on the host by csrc/hostpack.cu) must hand the caller exactly the arrays the
plain device -> host copies give."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NAMES = ('row', 'col', 'raw', 'size_factors', 'scaled', 'disp_idx', 'disp',
         'pvalues', 'llr', 'mu_hat_null', 'mu_hat_alt', 'qvalues')


def _host_inputs(big_count=False):
    import torch
    from hic3defdr_b200 import staging
    from hic3defdr_b200.synth import make_chrom
    out = []
    for i, n in enumerate((700, 500)):
        mats, bias, _ = make_chrom(n, 4, 40, seed=77 + i, amp=200.0)
        if big_count and i == 0:
            m = mats[1].tolil()
            m[10, 14] = 5_000_000_000            # does not fit 32 bits
            mats[1] = m.tocsr()
        out.append((staging.pinned_csr(mats),
                    torch.from_numpy(bias).pin_memory()))
    return out


def _run(host, compact):
    import torch
    from hic3defdr_b200 import engine, staging
    design = np.array([[1, 0], [1, 0], [0, 1], [0, 1]], dtype=bool)
    prefetch = staging.InputPrefetcher(host, depth=2)
    drain = staging.OutputDrain({}, after=prefetch, compact=compact)
    states, dpd, fns, stats = engine.run_to_qvalues(
        prefetch, design, dist_max=40, sink=drain)
    out = drain.wait()
    torch.cuda.synchronize()
    return {k: v.numpy().copy() for k, v in out.items()}, drain.nbytes


@pytest.mark.parametrize('big_count', [False, True])
def test_compact_drain_equals_plain_copies(big_count):
    host = _host_inputs(big_count)
    plain, nb_plain = _run(host, compact=False)
    compact, nb_compact = _run(host, compact=True)
    assert sorted(plain) == sorted(compact) == sorted(
        (i, k) for i in range(2) for k in NAMES)
    for key in plain:
        assert plain[key].dtype == compact[key].dtype, key
        np.testing.assert_array_equal(plain[key], compact[key], err_msg=str(key))
    assert plain[(0, 'raw')].dtype == np.int64
    if big_count:
        assert plain[(0, 'raw')].max() == 5_000_000_000
    assert nb_compact < 0.75 * nb_plain


@pytest.mark.parametrize('dtype', [np.int64, np.int32, np.float64])
def test_pinned_csr_upload_equals_plain(tmp_path, dtype):
    """Input matrices read by hostio.load_npz and re-based on page-locked
    memory (hostio.pin_csr) reach the device through asynchronous copies; the
    device CSR arrays equal those of the plain (blocking, pageable) upload of
    the matrix scipy.sparse.load_npz gives -- the reference's reader,
    analysis/analysis.py:85-86 -- including the fall-backs: a matrix with
    unsorted indices, an empty one, a row slice without pinned arrays."""
    import scipy.sparse as sparse
    import torch
    from hic3defdr_b200 import hostio, ops, staging
    rng = np.random.default_rng(11)
    dense = np.triu(rng.poisson(0.3, size=(400, 400))).astype(dtype)
    good = sparse.csr_matrix(dense)
    unsorted = good.copy()
    for i in range(0, 400, 7):                 # reverse the entries of some rows
        a, b = unsorted.indptr[i], unsorted.indptr[i + 1]
        unsorted.indices[a:b] = unsorted.indices[a:b][::-1].copy()
        unsorted.data[a:b] = unsorted.data[a:b][::-1].copy()
    unsorted.has_sorted_indices = False
    empty = sparse.csr_matrix((400, 400), dtype=dtype)
    mats, want = [], []
    for k, m in enumerate((good, unsorted, empty)):
        path = str(tmp_path / ('m%d.npz' % k))
        sparse.save_npz(path, m)
        got = hostio.pin_csr(hostio.load_npz(path).tocsr())
        assert (got != sparse.load_npz(path)).nnz == 0
        if k == 0:
            assert all(t.is_pinned() for t in got._h3d_pinned)
            assert got.data.ctypes.data == got._h3d_pinned[0].data_ptr()
            assert got.has_canonical_format
        mats.append(got)
        want.append(sparse.load_npz(path))
    a, b = ops.DeviceCSR(mats), ops.DeviceCSR(want)
    torch.cuda.synchronize()
    assert a.dtype == b.dtype and a.is64 == b.is64 and a.nnz == b.nnz
    for xs, ys in ((a.indptr, b.indptr), (a.indices, b.indices),
                   (a.data, b.data)):
        for x, y in zip(xs, ys):
            assert x.dtype == y.dtype and torch.equal(x, y)
    # row slices are new matrices: plain upload
    a = ops.DeviceCSR(staging.shard_rows(mats[:1], 100, 300))
    b = ops.DeviceCSR(staging.shard_rows(want[:1], 100, 300))
    assert torch.equal(a.data[0], b.data[0]) and \
        torch.equal(a.indptr[0], b.indptr[0])
