"""pytest configuration: the `gpu` marker, repo paths, and builders for the checkers.

The oracle (oracle/) is test infrastructure: it is built here, on demand, with gcc.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible, and vice versa nothing
    else needs one."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_build():
    """Compile oracle/_build/{libntsm_oracle.so,ntsm_oracle} if sources are newer."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return os.path.join(ROOT, "oracle", "_build")


@pytest.fixture(scope="session")
def oracle(oracle_build):
    import oracle_lib
    return oracle_lib.load(os.path.join(oracle_build, "libntsm_oracle.so"))


@pytest.fixture(scope="session")
def ref_bin():
    """The real reference binary, when it has been built (make -C oracle ref)."""
    p = os.path.join(ROOT, "oracle", "_ref", "ntsmCount")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/ntsmCount not built")
    return p


def golden_cases():
    d = os.path.join(GOLDEN, "cases")
    return sorted(os.listdir(d)) if os.path.isdir(d) else []
