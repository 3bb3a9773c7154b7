import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session")
def ref_drivers():
    """The reference's own driver functions (full_test, compare, main, ...) from the patched
    copy in oracle/_ref, bound to the GPU-backed classes by the package's loader."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref is not built")
    from active_matrix_factorization_b200 import drivers

    def load(name):
        return drivers.load(name, ref_loader.REF_DIR)
    return load
