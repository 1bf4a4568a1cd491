import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    """Path of libastrild_pk.so, building it (nvcc cross-compiles without a GPU) if absent."""
    path = os.path.join(ROOT, "astrild_b200", "lib", "libastrild_pk.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "astrild_b200", "csrc"), "-j8"],
                              stdout=subprocess.DEVNULL)
    return path


@pytest.fixture(scope="session")
def oracle_fast():
    from oracle import pk_oracle_fast
    pk_oracle_fast.build()
    return pk_oracle_fast
