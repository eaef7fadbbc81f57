"""CPU: the background ``.npy`` writer of the drop-in class
(hic3defdr_b200/analysis.py::_AsyncWriter) with
host data: files are complete after ``wait()``, failures surface there."""
import os

import numpy as np
import pytest
import torch


def test_writer_round_trip_and_error_propagation(tmp_path):
    from hic3defdr_b200.analysis import _AsyncWriter
    w = _AsyncWriter(3)
    rng = np.random.default_rng(0)
    arrays = {'a': rng.random((1000, 4)), 'b': rng.integers(0, 9, 5000),
              'c': rng.random(7) > 0.5, 'd': np.zeros((0, 2))}
    for k, v in arrays.items():
        w.submit(v, str(tmp_path / ('%s.npy' % k)))
    t = torch.arange(12, dtype=torch.float64).reshape(3, 4)
    w.submit(t, str(tmp_path / 't.npy'))
    w.wait()
    for k, v in arrays.items():
        got = np.load(str(tmp_path / ('%s.npy' % k)))
        assert got.dtype == v.dtype and np.array_equal(got, v)
    assert np.array_equal(np.load(str(tmp_path / 't.npy')), t.numpy())
    w.wait()                                   # idempotent
    w.submit(arrays['a'], str(tmp_path / 'no_such_dir' / 'x.npy'))
    with pytest.raises(OSError):
        w.wait()
    w.submit(arrays['b'], str(tmp_path / 'again.npy'))   # usable afterwards
    w.wait()
    assert os.path.exists(str(tmp_path / 'again.npy'))
