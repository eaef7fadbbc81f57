"""
TEST INFRASTRUCTURE -- imports the UNMODIFIED reference from /root/reference in
the build container (on the GPU box: from the copy under oracle/_ref that
``__graft_entry__.build()`` makes; it is git-ignored, never committed) so that
tests/golden/make_golden.py can record its outputs as fixtures and so that the
restatement in oracle/pipeline.py can be validated against the real thing.

Recipe:
  * oracle/lib5c_shim provides the four lib5c functions on the hot path;
  * a meta-path finder answers MagicMock modules for the plotting-only imports
    (matplotlib, seaborn, mpl_scatter_density, lib5c.plotters, ...), needed
    because hic3defdr/analysis/constructor.py:6-9 imports the plotting and
    simulation mixins unconditionally;
  * ONE documented normalisation: hic3defdr.util.binning.equal_bin is
    monkey-patched to use stable argsorts, because the shipped unstable sort
    (util/binning.py:25) makes the reference's own size factors depend on the
    sort implementation (up to 4.9e-3 relative).
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference():
    """/root/reference in the build container; on the GPU box the copy of the
    (pure-Python) package that ``__graft_entry__.build()`` placed under
    oracle/_ref (git-ignored, travels like the built .so files)."""
    env = os.environ.get('H3D_REFERENCE_ROOT')
    for root in ([env] if env else []) + ['/root/reference',
                                          os.path.join(_HERE, '_ref')]:
        if os.path.isdir(os.path.join(root, 'hic3defdr')):
            return root
    return env or '/root/reference'


REFERENCE_ROOT = _find_reference()

_STUB_PREFIXES = (
    'matplotlib', 'seaborn', 'mpl_scatter_density', 'lib5c.plotters',
    'lib5c.algorithms', 'lib5c.util.plotting', 'lib5c.util.distributions',
    'lib5c.util.bed', 'lib5c.util.primers', 'lib5c.parsers', 'lib5c.writers',
    'lib5c.util.counts', 'lib5c.util.bedgraph', 'lib5c.util.ast_eval',
)


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path=None, target=None):
        if any(name == p or name.startswith(p + '.') for p in _STUB_PREFIXES):
            return importlib.machinery.ModuleSpec(name, _StubLoader(),
                                                  is_package=True)
        return None


_installed = False


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'hic3defdr'))


def install(stable_equal_bin=True):
    """Makes ``import hic3defdr`` resolve to the unmodified reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(here)
    for p in (repo, os.path.join(here, 'lib5c_shim'), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.meta_path.insert(0, _StubFinder())
    import hic3defdr.util.binning as binning
    import hic3defdr.util.scaling as scaling
    if stable_equal_bin:
        import numpy as np

        def equal_bin(data, n_bins):
            idx = np.linspace(0, n_bins, data.size, endpoint=0, dtype=int)
            return idx[np.argsort(np.argsort(data, kind='stable'),
                                  kind='stable')]
        binning.equal_bin = equal_bin
        scaling.equal_bin = equal_bin
    _installed = True


def reference_class():
    install()
    from hic3defdr.analysis.constructor import HiC3DeFDR
    return HiC3DeFDR


def reference_modules():
    """Returns the reference's util modules on the hot path."""
    install()
    import hic3defdr.util.matrices as matrices
    import hic3defdr.util.scaling as scaling
    import hic3defdr.util.scaled_nb as scaled_nb
    import hic3defdr.util.dispersion as dispersion
    import hic3defdr.util.lowess as lowess
    import hic3defdr.util.lrt as lrt
    return types.SimpleNamespace(
        matrices=matrices, scaling=scaling, scaled_nb=scaled_nb,
        dispersion=dispersion, lowess=lowess, lrt=lrt)
