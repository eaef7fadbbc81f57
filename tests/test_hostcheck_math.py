"""CPU checks of the scalar FP64 building blocks in
hic3defdr_b200/csrc/h3d_math.cuh, compiled as host code by
tests/hostcheck/hostcheck.cpp (a test tool, never used by the product):
incomplete gamma pair and inverses vs scipy, q2q vs the oracle, fit_mu vs the
oracle, the Brent state machine vs scipy's bounded minimiser."""
import ctypes
import os
import subprocess
import warnings

import numpy as np
import pytest
import scipy.special as sp
from scipy.optimize import minimize_scalar

from oracle import pipeline as op
from tests.helpers import load_stage_golden

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'hostcheck', 'hostcheck.cpp')
LIB = os.path.join(HERE, 'hostcheck', 'libhostcheck.so')
dp = ctypes.POINTER(ctypes.c_double)


def P(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope='module')
def L():
    hdr = os.path.join(os.path.dirname(HERE), 'hic3defdr_b200', 'csrc',
                       'h3d_math.cuh')
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(
            os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.run(['g++', '-O2', '-shared', '-fPIC', '-o', LIB, SRC,
                        '-lm'], check=True)
    lib = ctypes.CDLL(LIB)
    lib.hc_brent.restype = ctypes.c_double
    return lib


def test_incomplete_gamma_pair(L):
    rng = np.random.default_rng(1)
    n = 100000
    a = 10 ** rng.uniform(-1, 3, n)
    x = a * 10 ** rng.normal(0, 0.5, n)
    x[:1000] = a[:1000] * (1 + rng.normal(0, 0.01, 1000))
    p, q = np.zeros(n), np.zeros(n)
    L.hc_gamma_pq(P(a), P(x), n, P(p), P(q))
    pr, qr = sp.gammainc(a, x), sp.gammaincc(a, x)
    small = np.where(pr < qr, np.abs(p - pr) / np.maximum(pr, 1e-300),
                     np.abs(q - qr) / np.maximum(qr, 1e-300))
    ok = np.minimum(pr, qr) > 1e-290
    # scipy itself is only good to ~1e-11 for a ~ 1e3 (checked with mpmath)
    assert small[ok].max() < 5e-11
    # values down to 1e-290: exponents of ~700 carry ~1e-13 in both codes
    assert small[ok & (a < 100)].max() < 1e-12


@pytest.mark.parametrize('upper', [0, 1])
def test_incomplete_gamma_inverse(L, upper):
    rng = np.random.default_rng(2 + upper)
    n = 100000
    a = 10 ** rng.uniform(-1, 3, n)
    t = 10 ** rng.uniform(-30, -0.01, n)
    y = np.zeros(n)
    L.hc_gamma_inv(P(a), P(t), n, upper, P(y))
    yr = sp.gammainccinv(a, t) if upper else sp.gammaincinv(a, t)
    m = np.isfinite(yr) & (yr > 1e-290)
    # the Halley iteration stops after a step below 1e-3 distribution widths
    # (error left ~ step^3; h3d_math.cuh lists accuracy against tolerance):
    # measured 1.7e-10 from the crude start y = a used here, 4.5e-11 through
    # q2q, which starts from the Wilson-Hilferty map (test_q2q_vs_oracle)
    assert (np.abs(y[m] - yr[m]) / yr[m]).max() < 5e-10


@pytest.mark.parametrize('alpha', [0.01, 0.0005, 0.3])
def test_q2q_vs_oracle(L, alpha):
    rng = np.random.default_rng(3)
    n = 100000
    mu_in = 10 ** rng.uniform(-1.2, 2.7, n)
    mu_out = mu_in * np.exp(rng.normal(0, 0.3, n))
    x = rng.poisson(rng.gamma(1 / alpha, mu_in * alpha)).astype(float)
    x[:100] = 0
    x[100:200] = np.round(mu_in[100:200] * 8 + 20)
    mi, mo = mu_in.copy(), mu_out.copy()
    ref = op.q2q(x, mi, mo, alpha)          # clamps mi / mo in place
    out = np.zeros(n)
    L.hc_q2q(P(x), P(mi), P(mo), ctypes.c_double(alpha), n, P(out))
    assert np.array_equal(np.isfinite(out), np.isfinite(ref))
    m = np.isfinite(ref)
    # relative to max(value, 1e-3): tiny outputs are cancellation residues;
    # 4.5e-11 at the Halley stopping tolerance of 1e-3 (h3d_math.cuh), 20x
    # inside the 1e-9 budget of the dispersions the pseudo-data feed
    assert (np.abs(out[m] - ref[m]) / np.maximum(ref[m], 1e-3)).max() < 1e-10


def test_fit_mu_vs_oracle(L):
    rng = np.random.default_rng(4)
    n, R = 50000, 4
    b = rng.lognormal(0, 0.4, (n, R))
    mu = 10 ** rng.uniform(-1, 3, n)
    alpha = 10 ** rng.uniform(-3, 0.5, (n, R))
    x = rng.poisson(rng.gamma(1 / alpha, mu[:, None] * b * alpha)).astype(float)
    keep = x.sum(1) > 0
    x, b, alpha = x[keep], b[keep], alpha[keep]
    n = len(x)
    out = np.zeros(n)
    st = np.zeros(n, dtype=np.int32)
    L.hc_fit_mu(P(x), P(b), P(alpha), n, R, P(out),
                st.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert (st == 0).all()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = op.fit_mu_hat(x, b, alpha)
    assert (np.abs(out - ref) / ref).max() < 1e-10
    g = np.sum((x - out[:, None] * b) /
               (out[:, None] + alpha * out[:, None] ** 2 * b), axis=1)
    assert np.abs(g).max() < 1e-12      # tighter than the reference's 1e-5


def test_brent_state_machine_matches_scipy(L):
    s = load_stage_golden()
    CB = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_double)
    for data in (s['equalize_0.01'], s['equalize_0.2'], s['equalize4_0.05']):
        f = lambda t: float(op.cml_nll(data, t))
        res = minimize_scalar(f, bounds=(1e-4, 100. / 101), method='bounded')
        nfev, flag = ctypes.c_int(), ctypes.c_int()
        xf = L.hc_brent(CB(f), ctypes.c_double(1e-4),
                        ctypes.c_double(100. / 101), ctypes.c_double(1e-5),
                        500, ctypes.byref(nfev), ctypes.byref(flag))
        assert xf == res.x and nfev.value == res.nfev and flag.value == 0


def test_table_log_and_stirling_core(L):
    """the table-driven log of the likelihood kernel and the shifted Stirling
    log-gamma built on it, against libm / scipy in the ranges the kernel uses
    (arguments y + r with r in [0.01, 1e4], y in [0, 1e6])"""
    rng = np.random.default_rng(7)
    n = 400000
    x = 10 ** rng.uniform(np.log10(3.0), 7, n)
    x[:2000] = np.nextafter(2.0 ** rng.integers(2, 20, 2000), 0)   # just below 2^k
    x[2000:4000] = 2.0 ** rng.integers(2, 20, 2000)
    out = np.zeros(n)
    L.hc_fast_log(P(x), n, P(out))
    ref = np.log(x)
    ulp = np.abs(out - ref) / np.spacing(ref)
    assert ulp.max() <= 2.0, ulp.max()
    # the general-purpose entry (m_log): absolute accuracy over the whole
    # positive range, including arguments next to 1, subnormals, zero and inf
    xg = np.concatenate([10 ** rng.uniform(-300, 300, 200000),
                         1 + rng.normal(0, 1e-3, 100000),
                         np.exp(rng.normal(0, 0.3, 100000)),
                         [5e-324, 1e-310, 2.2250738585072014e-308, 0.0, np.inf]])
    og = np.zeros(len(xg))
    L.hc_fast_log(P(xg), len(xg), P(og))
    with np.errstate(divide='ignore'):
        rg = np.log(xg)
    fin = np.isfinite(rg)
    assert np.array_equal(og[~fin], rg[~fin])
    assert (np.abs(og[fin] - rg[fin]) /
            np.maximum(1.0, np.abs(rg[fin]))).max() < 2.5e-16
    xs = np.concatenate([10 ** rng.uniform(-2, 1, 200000),
                         10 ** rng.uniform(1, 6, 200000)])
    out = np.zeros(len(xs))
    L.hc_lgamma_core(P(xs), len(xs), P(out))
    ref = sp.gammaln(xs)
    # absolute error in units of the largest intermediate: arguments below 10
    # are shifted to [10, 11), where (x - .5) ln x ~ 25
    scale = np.maximum(np.abs(ref), 25.0)
    # a handful of roundings at that magnitude: under 5 ulp
    assert (np.abs(out - ref) / scale).max() < 1e-15


def host_equalize(L, x, f, alpha):
    """equalize (scaled_nb.py:207-214) from the host build of fit_mu and
    q2q_one, with the replicate-ordered clamp of equalize_kernel"""
    n, r = x.shape
    xs = np.ascontiguousarray(x, dtype=float)
    fs = np.ascontiguousarray(f, dtype=float)
    mu = np.zeros(n)
    st = np.zeros(n, dtype=np.int32)
    L.hc_fit_mu(P(xs), P(fs), P(np.full((n, r), float(alpha))), n, r, P(mu),
                st.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    assert not st.any()
    mu_out = mu * np.exp(np.log(fs).sum(axis=1) / r)
    out = np.zeros((n, r))
    for k in range(r):
        mi = mu * fs[:, k]
        low = ~((mi >= 0.25) & (mu_out >= 0.25))    # scaled_nb.py:240-242:
        mi[low] = 0.25                              # mu_out is shared by the
        mu_out[low] = 0.25                          # later replicates
        col = np.zeros(n)
        L.hc_q2q(P(xs[:, k].copy()), P(mi), P(mu_out.copy()),
                 ctypes.c_double(alpha), n, P(col))
        out[:, k] = col
    return out


def check_pseudo(got, want, rtol=1e-9):
    """pseudo-data parity: the same entries are infinite (the reference's
    tail-probability underflow), the finite ones agree to ``rtol`` relative to
    max(value, 1e-3) -- tiny outputs are differences of the normal and the
    gamma map that nearly cancel"""
    assert got.shape == want.shape
    assert np.array_equal(np.isposinf(got), np.isposinf(want))
    assert not np.isnan(got).any() and not np.isnan(want).any()
    m = np.isfinite(want)
    err = np.abs(got[m] - want[m]) / np.maximum(want[m], 1e-3)
    assert err.max() < rtol, err.max()
    return float(err.max())


@pytest.mark.parametrize('case', ['small', 'tails', 'grid', 'three'])
def test_equalize_edges_vs_recorded_reference(L, case):
    """host build of the device code on the reference's recorded edge cases
    (tests/golden/make_golden_equalize.py); the device run of the same fixture
    is tests/test_gpu_disp.py::test_equalize_device_vs_recorded_reference"""
    g = np.load(os.path.join(HERE, 'golden', 'ref_equalize_edges.npz'))
    x, f = g['%s_x' % case], g['%s_f' % case]
    for a in g['alphas']:
        check_pseudo(host_equalize(L, x, f, a), g['%s_%g' % (case, a)])


def _random_equalize_case(rng):
    """pixels x replicates far beyond the fixtures: means 0.1 .. 3e4 with a
    log-normal spread, combined factors spread up to sigma 1.2, dispersions
    1e-4 .. 20, a third of the cases with 30 % zero counts"""
    n, r = 200, int(rng.integers(2, 5))
    scale = 10 ** rng.uniform(-1, 4.5)
    disp = max(10 ** rng.uniform(-4, 1.3), 1e-3)
    f = np.exp(rng.normal(0, rng.uniform(0.05, 1.2), size=(n, r)))
    mu = scale * np.exp(rng.normal(0, 1, size=(n, 1)))
    x = rng.poisson(np.minimum(rng.gamma(1 / disp, mu * f * disp), 1e9)
                    ).astype(float)
    if rng.random() < 0.3:
        x[rng.random(x.shape) < 0.3] = 0
    x[x.sum(axis=1) == 0, 0] = 1.0     # an all-zero pixel has no root (the
    return x, f, 10 ** rng.uniform(-4, 1.3)     # reference raises): never tested


def test_equalize_random_sweep_vs_oracle(L):
    """host build of fit_mu + q2q_one against the oracle (= the reference's
    scipy arithmetic, bit for bit) on 16 k random elements.  Infinite entries
    coincide, 99.9 % of the finite ones agree to 1e-9 and all to 1e-7: the
    rest sits where the REFERENCE loses digits (scipy's gammaincinv in far
    tails, see test_far_tail_quantile_map_is_exact, and the cancellation of
    the normal map at x = 0 under a huge mean)."""
    rng = np.random.default_rng(20261019)
    errs = []
    for _ in range(25):
        x, f, alpha = _random_equalize_case(rng)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            want = op.equalize(x, f, alpha)
        got = host_equalize(L, x, f, alpha)
        assert np.array_equal(np.isposinf(got), np.isposinf(want))
        assert not np.isnan(got).any()
        m = np.isfinite(want)
        errs.append(np.abs(got[m] - want[m]) / np.maximum(want[m], 1e-3))
    err = np.concatenate(errs)
    print('equalize sweep: %d elements, max %.2e, %d above 1e-9'
          % (len(err), err.max(), int((err > 1e-9).sum())))
    assert err.max() < 1e-7
    assert (err > 1e-9).mean() < 1e-3


def test_far_tail_quantile_map_is_exact(L):
    """One element of the sweep where the two disagree by 7.4e-9, settled in
    80-digit arithmetic (mpmath series of the incomplete gamma function +
    Newton): x = 493 under a mean of 10093 (17.6 sigma into the lower tail,
    ln P = -715.6).  The kernel's value is exact to 3e-15; the reference's
    (scipy gammaincinv) carries the 7.4e-9."""
    mp = pytest.importorskip('mpmath')
    x, mu_in, mu_out, alpha = 493.0, 10092.906731309311, 4568.4825211657217, \
        0.0028062597011992690
    got = np.zeros(1)
    L.hc_q2q(P(np.array([x])), P(np.array([mu_in])), P(np.array([mu_out])),
             ctypes.c_double(alpha), 1, P(got))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = op.q2q(np.array([x]), np.array([mu_in]), np.array([mu_out]),
                     alpha)[0]
    with mp.workdps(80):
        def log_p(a, y):          # ln of the regularised lower incomplete gamma
            term = tot = mp.mpf(1)
            n = 0
            while term > tot * mp.mpf(10) ** -70:
                n += 1
                term *= y / (a + n)
                tot += term
            return a * mp.log(y) - y - mp.loggamma(a + 1) + mp.log(tot)
        m_in, m_out, al = mp.mpf(mu_in), mp.mpf(mu_out), mp.mpf(alpha)
        r_in, r_out = 1 + al * m_in, 1 + al * m_out
        qn = m_out + (mp.mpf(x) - m_in) * mp.sqrt(m_out * r_out / (m_in * r_in))
        target = log_p(m_in / r_in, mp.mpf(x) / r_in)
        a_out = m_out / r_out
        y = (2 * mp.mpf(float(ref)) - qn) / r_out
        for _ in range(40):
            g = log_p(a_out, y) - target
            h = y * mp.mpf(10) ** -25
            step = g * h / (log_p(a_out, y + h) - log_p(a_out, y))
            y -= step
            if abs(step) < abs(y) * mp.mpf(10) ** -40:
                break
        exact = (qn + y * r_out) / 2
        err_ours = float(abs(mp.mpf(float(got[0])) - exact) / exact)
        err_ref = float(abs(mp.mpf(float(ref)) - exact) / exact)
    print('far tail: ours %.15g (%.1e), reference %.15g (%.1e)'
          % (got[0], err_ours, ref, err_ref))
    assert err_ours < 1e-13
    assert 1e-9 < err_ref < 1e-7           # the reference's own error


@pytest.mark.parametrize('df', [1, 2, 3])
def test_chi2_survival_function_vs_scipy(L, df):
    """chi2_sf (the LRT's p-value, util/lrt.py:47: ``stats.chi2(df).sf``) over
    15 decades of the statistic, df = C - 1 for 2 to 4 conditions: 1e-12
    relative down to p = 1e-290 (the argument of exp(-x/2) carries
    |x| 1e-16 either way); where scipy flushes to zero the kernel's value is
    zero or still a subnormal (the true value is one); sf(0) = 1,
    sf(x >= 1500) = sf(inf) = 0."""
    import scipy.stats as st
    rng = np.random.default_rng(df)
    x = np.concatenate([10 ** rng.uniform(-14, 3.3, 100000),
                        [0.0, 1e-300, 1e-8, 745.0, 1400.0, 1500.0, 5000.0,
                         np.inf]])
    out = np.zeros_like(x)
    L.hc_chi2_sf(P(x), len(x), df, P(out))
    want = st.chi2.sf(x, df)
    m = want > 1e-290
    assert (np.abs(out[m] - want[m]) / want[m]).max() < 1e-12
    assert (out[want == 0] < 2.3e-308).all() and (out[x == 0] == 1).all()
    assert (out[x >= 1500.0] == 0).all()
    sub = ~m & (want > 0)                     # subnormal results: absolute
    assert (np.abs(out[sub] - want[sub]) <= 1e-12 * want[sub] + 1e-320).all()


def test_lgamma_pos_vs_scipy(L):
    """lgamma_pos (constants of the conditional likelihood, lgamma(n r) and
    lgamma(r), util/dispersion.py:72-75) for 1e-3 <= x <= 1e7"""
    rng = np.random.default_rng(4)
    xs = np.concatenate([10 ** rng.uniform(-3, 7, 100000),
                         [1.0, 2.0, 0.5, 9.999999, 10.0, 10.000001]])
    out = np.zeros_like(xs)
    L.hc_lgamma_pos(P(xs), len(xs), P(out))
    want = sp.gammaln(xs)
    assert (np.abs(out - want) / np.maximum(np.abs(want), 1.0)).max() < 3e-14
