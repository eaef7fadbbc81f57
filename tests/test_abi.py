"""CPU checks of the C-ABI library: it loads, exports every symbol that
include/h3d.h declares, and the Python binding table matches the header."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, 'include', 'h3d.h')
LIB = os.path.join(REPO, 'hic3defdr_b200', 'libh3d.so')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(h3d_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_binding_table():
    from hic3defdr_b200._native import SIGNATURES
    assert sorted(SIGNATURES) == declared_symbols()


def test_library_loads_and_exports_every_symbol():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    lib.h3d_version.restype = ctypes.c_int
    assert lib.h3d_version() >= 100
    lib.h3d_launch_count.restype = ctypes.c_ulonglong
    assert lib.h3d_launch_count() == 0


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from hic3defdr_b200 import ops
    from hic3defdr_b200._native import H3DError
    import numpy as np
    with pytest.raises(H3DError):
        ops.fit_mu_hat(np.ones((2, 2)), np.ones((2, 2)), 0.1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, 'hic3defdr_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(root, f)).read()
                assert 'oracle' not in re.sub(r'#.*', '', src).replace(
                    'DESIGN.md', ''), f
