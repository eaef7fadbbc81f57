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


def _small_inputs(root, chroms, n_reps=4, loops=False):
    from hic3defdr_b200.synth import write_dataset
    return write_dataset(str(root), {c: 60 + 7 * i for i, c in enumerate(chroms)},
                         n_reps=n_reps, dist_max=20, config=7, loops=loops)


@pytest.mark.parametrize('n_threads,loops', [(-1, False), (0, True), (3, True)])
def test_prefetched_inputs_match_the_plain_loader(tmp_path, n_threads, loops):
    """The look-ahead loader of ``prepare_data`` (several chromosomes in flight
    on one pool) yields every chromosome once, in order, with the inputs the
    one-chromosome loader gives (reference: analysis/analysis.py:84-101 reads
    the same files one by one)."""
    from hic3defdr_b200 import HiC3DeFDR
    chroms = ['chr%d' % i for i in range(1, 8)]
    kw = _small_inputs(tmp_path / 'in', chroms, loops=loops)
    if not loops:
        kw.pop('loop_patterns')
    h = HiC3DeFDR(outdir=str(tmp_path / 'out'), dist_thresh_max=20, **kw)
    seen = []
    for c, (bias, mats, loop_pixels) in h._prefetched_inputs(chroms, n_threads):
        seen.append(c)
        wb, wm, wl = h._load_inputs(c, 0)
        assert np.array_equal(bias, wb) and bias.flags.c_contiguous
        assert len(mats) == len(wm) == 4
        for a, b in zip(mats, wm):
            assert a.format == 'csr' and (a != b).nnz == 0 and a.dtype == b.dtype
        assert loop_pixels == wl and (loop_pixels is not None) == loops
    assert seen == chroms
    assert list(h._prefetched_inputs([], n_threads)) == []


def test_prefetched_inputs_surface_a_missing_file(tmp_path):
    from hic3defdr_b200 import HiC3DeFDR
    chroms = ['chr1', 'chr2', 'chr3']
    kw = _small_inputs(tmp_path / 'in', chroms)
    kw.pop('loop_patterns')
    os.remove(kw['raw_npz_patterns'][1].replace('<chrom>', 'chr2'))
    h = HiC3DeFDR(outdir=str(tmp_path / 'out'), dist_thresh_max=20, **kw)
    it = h._prefetched_inputs(chroms, -1)
    assert next(it)[0] == 'chr1'
    with pytest.raises(OSError):
        next(it)


def test_chunked_npy_writer_equals_np_save(tmp_path):
    """The staged writer streams an array to disk chunk by chunk
    (``write_npy_chunks``): the file must be np.save's, byte for byte, for
    every dtype / shape the pipeline writes (int32 row, int64 (N, R) raw, bool
    masks, float64 vectors and matrices, empty arrays, 0-d)."""
    from hic3defdr_b200.analysis import _AsyncWriter
    rng = np.random.default_rng(1)
    arrays = [rng.random((1000, 4)), rng.integers(0, 9, 5000).astype(np.int32),
              rng.random(7) > 0.5, np.zeros((0, 2)), np.zeros(0, dtype=bool),
              rng.integers(0, 2 ** 40, (33, 2)), np.float64(3.5) * np.ones(()),
              rng.random((5, 3, 2)).astype(np.float32)]
    for i, a in enumerate(arrays):
        a = np.ascontiguousarray(a)
        ref, got = str(tmp_path / 'ref.npy'), str(tmp_path / 'got.npy')
        np.save(ref, a)
        raw = a.tobytes()
        for step in (1 << 20, 997):
            _AsyncWriter.write_npy_chunks(
                got, a.dtype, a.shape,
                (raw[j:j + step] for j in range(0, len(raw), step)))
            assert open(got, 'rb').read() == open(ref, 'rb').read(), (i, step)
            back = np.load(got)
            assert back.dtype == a.dtype and np.array_equal(back, a)
