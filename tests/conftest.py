import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    lib = os.path.join(ROOT, "trace.jl_b200", "csrc", "libtrace_cuda.so")
    ref = os.path.join(ROOT, "oracle", "libtrace_ref.so")
    if not (os.path.exists(lib) and os.path.exists(ref)):
        import __graft_entry__ as g
        g.build()


@pytest.fixture(scope="session")
def T():
    import trace_jl_b200
    return trace_jl_b200


@pytest.fixture(scope="session")
def ctx(T):
    c = T.Context(0)
    yield c
    c.close()
