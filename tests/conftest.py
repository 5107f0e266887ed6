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


class Cases:
    """Lazy view of tests/golden/reference_cases.npz: cases['c1']['T'] etc."""

    def __init__(self, path):
        self._z = np.load(path)
        self._names = {}
        for key in self._z.files:
            case, name = key.split("/", 1)
            self._names.setdefault(case, []).append(name)

    def names(self):
        return sorted(self._names)

    def __getitem__(self, case):
        return {name: self._z[f"{case}/{name}"] for name in self._names[case]}


@pytest.fixture(scope="session")
def ref_cases():
    return Cases(os.path.join(GOLDEN, "reference_cases.npz"))


@pytest.fixture(scope="session")
def tutorial_goldens():
    z = np.load(os.path.join(GOLDEN, "tutorial_goldens.npz"))
    return {k: z[k] for k in z.files}
