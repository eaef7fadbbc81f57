"""CPU: the zip/inflate reader of the input matrices
(hic3defdr_b200/hostio.py) gives what
``scipy.sparse.load_npz`` gives -- the call the reference makes at
analysis/analysis.py:85-86 -- for every container ``save_npz`` writes."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
import scipy.sparse as sparse

from hic3defdr_b200 import hostio


def _same(a, b):
    assert type(a) is type(b) and a.format == b.format
    assert a.shape == b.shape and a.dtype == b.dtype
    assert (a != b).nnz == 0
    for name in ('data', 'indices', 'indptr', 'row', 'col', 'offsets'):
        if hasattr(b, name):
            x, y = getattr(a, name), getattr(b, name)
            assert x.dtype == y.dtype and np.array_equal(x, y), name


@pytest.mark.parametrize('fmt', ['csr', 'csc', 'coo', 'bsr', 'dia'])
@pytest.mark.parametrize('dtype', [np.int64, np.int32, np.float64, np.float32])
@pytest.mark.parametrize('compressed', [True, False])
def test_load_npz_equals_scipy(tmp_path, fmt, dtype, compressed):
    rng = np.random.default_rng(3)
    dense = rng.poisson(0.3, size=(97, 97)).astype(dtype)
    dense = np.triu(dense)
    if fmt == 'coo':
        # matrices written by older scipy store row/col members; newer ones a
        # coords member (that container goes to scipy's own loader)
        m = sparse.coo_matrix(dense)
    else:
        m = getattr(sparse, fmt + '_matrix')(dense)
    path = str(tmp_path / 'm.npz')
    sparse.save_npz(path, m, compressed=compressed)
    want = sparse.load_npz(path)
    _same(hostio.load_npz(path), want)
    with ThreadPoolExecutor(3) as pool:
        _same(hostio.load_npz(path, pool), want)


def test_empty_and_sparse_array_containers(tmp_path):
    path = str(tmp_path / 'e.npz')
    sparse.save_npz(path, sparse.csr_matrix((50, 50), dtype=np.int64))
    _same(hostio.load_npz(path), sparse.load_npz(path))
    path = str(tmp_path / 'a.npz')
    sparse.save_npz(path, sparse.csr_array(np.eye(4)))
    got, want = hostio.load_npz(path), sparse.load_npz(path)
    assert type(got) is type(want) and (got != want).nnz == 0


def test_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(5)
    path = str(tmp_path / 'm.npz')
    sparse.save_npz(path, sparse.csr_matrix(
        rng.poisson(0.5, size=(300, 300)).astype(np.int64)))
    blob = bytearray(open(path, 'rb').read())
    blob[len(blob) // 2] ^= 0x5a                 # inside a deflate stream
    open(path, 'wb').write(bytes(blob))
    with pytest.raises(Exception):
        hostio.load_npz(path)
    open(path, 'wb').write(b'not a zip file')
    with pytest.raises(Exception):
        hostio.load_npz(path)


def test_reader_feeds_the_sharding_helpers(tmp_path):
    """read-only views are enough for the host code that touches the matrices
    before the upload (staging.row_weights / shard_rows)."""
    from hic3defdr_b200 import staging
    rng = np.random.default_rng(9)
    path = str(tmp_path / 'm.npz')
    m = sparse.csr_matrix(np.triu(rng.poisson(0.4, size=(120, 120))).astype(np.int64))
    sparse.save_npz(path, m)
    got = hostio.load_npz(path).tocsr()
    assert not got.data.flags.writeable and got.has_canonical_format
    np.testing.assert_array_equal(staging.row_weights([got]),
                                  staging.row_weights([m]))
    a, b = staging.shard_rows([got], 30, 90)[0], staging.shard_rows([m], 30, 90)[0]
    assert (a != b).nnz == 0
