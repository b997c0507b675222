import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_tests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fixtures():
    return np.load(os.path.join(ROOT, "tests", "golden", "oracle_fixtures.npz"))


@pytest.fixture(scope="session")
def oracle_build():
    """Make sure the C oracle is built (test infrastructure)."""
    import subprocess

    so = os.path.join(ROOT, "oracle", "_build", "libflomat_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return so


@pytest.fixture(scope="session")
def providers(oracle_build):
    from oracle.flomat_oracle import CProvider, OpenBLASProvider

    out = [CProvider()]
    try:
        out.append(OpenBLASProvider())
    except Exception:  # scipy wheel without the bundled library
        pass
    return out


@pytest.fixture(scope="session")
def ref(providers):
    """The oracle used as the checker in GPU parity tests: OpenBLAS (the reference's library class) when present."""
    from oracle.flomat_oracle import Flomat

    return Flomat(providers[-1])


@pytest.fixture(scope="session")
def fm():
    """The product: sci_b200 over libflomat_cuda, initialised on cuda:0.  Raises (never skips to a CPU path)."""
    import sci_b200

    sci_b200.init(0)
    return sci_b200
