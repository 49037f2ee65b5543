import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kat():
    return json.load(open(os.path.join(GOLDEN, "kat.json")))


@pytest.fixture(scope="session")
def refvec():
    return np.load(os.path.join(GOLDEN, "ref_vectors.npz"))


@pytest.fixture(scope="session")
def refvec2():
    """Round-2 vectors from the compiled reference (tools/make_golden_r2.py): leave-one-out ExomeCount, envelope, a1 ~ -1."""
    return np.load(os.path.join(GOLDEN, "ref_vectors_r2.npz"))


@pytest.fixture(scope="session")
def exomecount():
    return np.load(os.path.join(GOLDEN, "exomecount.npz"))


@pytest.fixture(scope="session")
def geometry():
    return np.load(os.path.join(GOLDEN, "exons_hg19_geometry.npz"))


@pytest.fixture(scope="session")
def port():
    from oracle import port as p
    p.lib()
    return p


@pytest.fixture(scope="session")
def ref():
    """The compiled reference, when its prebuilt .so is present (built in the dev container)."""
    from oracle import ref as r
    if not r.available():
        pytest.skip("oracle/_ref/libexomedepth_ref.so not built (needs /root/reference)")
    r.api().quiet(True)
    return r


def hmm_cases(refvec):
    n = int(refvec["hmm_n"][0])
    for i in range(n):
        yield (refvec[f"hmm{i}_T"], refvec[f"hmm{i}_ll"], refvec[f"hmm{i}_pos"], float(refvec[f"hmm{i}_L"][0]),
               refvec[f"hmm{i}_path"].astype(np.int32), refvec[f"hmm{i}_calls"].astype(np.int64))
