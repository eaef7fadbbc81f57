"""
Edge-case fixture for the device pseudo-data:
outputs of the UNMODIFIED reference's ``equalize``
(/root/reference/hic3defdr/util/scaled_nb.py:186-214, q2qnbinom :217-275)
recorded in the build container.

    python tests/golden/make_golden_equalize.py  ->  ref_equalize_edges.npz

Cases (each at several dispersions):
  small   counts of 0..3 with small factors: mu_in or mu_out < 0.25 -> the
          order-dependent clamp of scaled_nb.py:240-242, x = 0 in the left tail
  tails   one replicate far above, the other far below the fitted mean: both
          tails, up to the underflow of the reference's tail probabilities
          (sf -> 0 -> isf = inf; cdf -> 0 -> ppf = -inf / 0 -> clipped to 0)
  grid    log-uniform means 0.05..2e4 and log-normal factors (sigma 1): series
          and continued-fraction branches, small and large shape parameters
  three   three replicates per condition (odd count: the R_c = 4 kernel
          instance with one masked replicate)
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import refrun  # noqa: E402

ALPHAS = (1e-4, 1e-3, 0.01, 0.05, 0.3, 2.0, 20.0)


def cases():
    rng = np.random.default_rng(20261019)
    out = {}
    # small
    x = rng.integers(0, 4, size=(600, 2))
    f = np.exp(rng.normal(-1.5, 0.8, size=(600, 2)))
    x[:4] = [[0, 1], [1, 0], [0, 3], [2, 0]]
    out['small'] = (x, f)
    # tails
    hi = np.round(10 ** rng.uniform(1, 5.2, size=400))
    lo = np.round(hi * 10 ** rng.uniform(-5, -0.05, size=400))
    x = np.stack([hi, lo], axis=1)
    flip = rng.random(400) < 0.5
    x[flip] = x[flip][:, ::-1]
    f = np.exp(rng.normal(0, 0.1, size=(400, 2)))
    out['tails'] = (x.astype(np.int64), f)
    # grid
    n = 4000
    mu = 10 ** rng.uniform(np.log10(0.05), np.log10(2e4), size=n)
    f = np.exp(rng.normal(0, 1.0, size=(n, 2)))
    disp = 10 ** rng.uniform(-3, 0.5, size=n)
    lam = rng.gamma(1 / disp[:, None], mu[:, None] * f * disp[:, None])
    out['grid'] = (rng.poisson(lam).astype(np.int64), f)
    # three replicates
    n = 1500
    mu = 10 ** rng.uniform(-0.5, 3, size=n)
    f = np.exp(rng.normal(0, 0.4, size=(n, 3)))
    lam = rng.gamma(1 / 0.05, mu[:, None] * f * 0.05)
    out['three'] = (rng.poisson(lam).astype(np.int64), f)
    for k, (x, f) in list(out.items()):
        keep = x.sum(axis=1) > 0
        out[k] = (x[keep], f[keep])
    return out


def main():
    ref = refrun.reference_modules()
    out = {'alphas': np.array(ALPHAS)}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for name, (x, f) in cases().items():
            out['%s_x' % name], out['%s_f' % name] = x, f
            for a in ALPHAS:
                out['%s_%g' % (name, a)] = ref.scaled_nb.equalize(
                    x, f.copy(), a)
    np.savez_compressed(os.path.join(HERE, 'ref_equalize_edges.npz'), **out)
    for k in sorted(out):
        if k.split('_')[-1] not in ('x', 'f') and k != 'alphas':
            v = out[k]
            print('%-14s n=%5d  zeros %5d  inf %4d  nan %3d  max finite %.3g' % (
                k, v.size, int((v == 0).sum()), int(np.isinf(v).sum()),
                int(np.isnan(v).sum()), np.nanmax(v[np.isfinite(v)])))


if __name__ == '__main__':
    main()
