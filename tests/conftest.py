import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import refvpic
    return refvpic.load_oracle()


@pytest.fixture(scope="session")
def ref_scalar():
    """The unmodified reference, scalar build (oracle/_ref/libvpic_ref_scalar.so)."""
    import refvpic
    if not refvpic.have_ref("scalar"):
        pytest.skip("oracle/_ref/libvpic_ref_scalar.so not built (needs /root/reference: make -C oracle ref)")
    return refvpic.load_ref("scalar", tpp=1)
