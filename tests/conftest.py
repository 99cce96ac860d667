import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The shared libraries are build artefacts (git-ignored).  On a fresh checkout build them once
    # (nvcc cross-compiles sm_100a without a GPU); on the GPU box they travel prebuilt.
    lib = os.path.join(ROOT, "duckdb_imputation_b200", "lib")
    need = [os.path.join(lib, "libcofactor_b200.so"), os.path.join(lib, "libduckdb_imputation_b200.so"),
            os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def goldens():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")))
