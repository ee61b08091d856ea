import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc

    orc.build()
    return orc


@pytest.fixture
def get_assemblers(request):
    """The reference's backend switch (python/tests/conftest.py:4-22, options "C++" and "numba") with the third
    option this repository adds: "cuda" = dolfinx_mpc_b200 (device kernels behind the same two callables)."""
    if request.param == "cuda":
        from dolfinx_mpc_b200 import assemble_matrix, assemble_vector

        return (assemble_matrix, assemble_vector)
    raise RuntimeError(f"Undefined assembler type: {request.param}.\nOptions here are 'cuda'")
