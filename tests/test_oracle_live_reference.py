"""CPU, build container only: the oracle restatement against the UNMODIFIED
reference run live on freshly generated datasets (tests/live_reference_check.py
in subprocesses), beyond the recorded fixtures of tests/golden/: six
replicates, three chromosomes, loop_idx, conditional_scaling, unweighted
lowess, refit_mu=False, a fixed lowess fraction.  Every saved array of
``run_to_qvalues`` and the pickled dispersion functions must be bit-identical.
Skipped where no reference tree exists (/root/reference or oracle/_ref)."""
import os
import subprocess
import sys

import pytest

from oracle import refrun
from tests.live_reference_check import VARIANTS

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(
    not refrun.available(), reason='no reference tree in this environment')


@pytest.fixture(scope='module')
def runs():
    procs = {v: subprocess.Popen(
        [sys.executable, os.path.join(REPO, 'tests', 'live_reference_check.py'),
         v], cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
        text=True) for v in VARIANTS}
    out = {}
    for v, p in procs.items():
        try:
            so, se = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            p.kill()
            so, se = p.communicate()
        out[v] = (p.returncode, so, se)
    return out


@pytest.mark.parametrize('variant', sorted(VARIANTS))
def test_oracle_equals_live_reference(runs, variant):
    rc, so, se = runs[variant]
    assert rc == 0, se[-2000:]
    assert 'live reference check %s: ok' % variant in so
