import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ila():
    import ila_b200
    return ila_b200


def bullhead_grid(nx, ny):
    """λ grid of BASELINE config C3 (window of examples/toeplitz.jl:72)."""
    import numpy as np
    x = np.linspace(-10.0, 7.0, nx)
    y = np.linspace(-7.0, 8.0, ny)
    return x[None, :] + 1j * y[:, None]


BULLHEAD_COLUMN = [2j, 0.0, 0.0, 1.0, 0.7]   # data column of BandedMatrix(-3=>7/10, -2=>1, 1=>2im): +1, 0, -1, -2, -3
