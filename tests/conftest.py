import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    _ensure_built()  # fixtures.load() parses the bundled JSON with the product's host parser


def _ensure_built():
    """Build libp2v.so in-tree when it is missing or stale (nvcc cross-compiles without a GPU; ~1 min once)."""
    import plonky2_verifier_b200 as m

    m.build()
    return m


@pytest.fixture(scope="session")
def p2v():
    return _ensure_built()


@pytest.fixture(scope="session")
def orc():
    import oracle_lib

    return oracle_lib.load()


@pytest.fixture(scope="session")
def ctx(p2v):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    c = p2v.Context(0)
    yield c
    c.close()
