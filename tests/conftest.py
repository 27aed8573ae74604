import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """The in-tree CUDA library (built on demand; nvcc cross-compiles without a GPU)."""
    from nest_b200 import build
    build.build()
    from nest_b200 import capi
    capi.lib()
    return capi


@pytest.fixture(scope="session")
def ctx(built):
    c = built.Context(0)
    yield c
    c.close()


@pytest.fixture()
def ctx_small_chunks(built):
    """A context whose internal chunk is 2^22 events (the default is 2^24), so that tests of a few million events
    cross chunk boundaries."""
    import os
    os.environ["NB200_CHUNK_LOG2"] = "22"
    try:
        c = built.Context(0)
    finally:
        del os.environ["NB200_CHUNK_LOG2"]
    yield c
    c.close()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref/libnestprobe.so); prebuilt files travel to the GPU box."""
    import refprobe
    r = refprobe.reference()
    if r is None:
        pytest.skip("oracle/_ref/libnestprobe.so not built (needs /root/reference; see oracle/Makefile)")
    return r


@pytest.fixture(scope="session")
def orc():
    """The CPU restatement oracle/liboracle.so."""
    import refprobe
    o = refprobe.oracle()
    if o is None:
        pytest.skip("oracle/liboracle.so not built (make -C oracle oracle)")
    return o
