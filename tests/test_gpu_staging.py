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
