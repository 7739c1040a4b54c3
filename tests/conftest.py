import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200; run with `-m gpu` on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle: plain-C restatement + (when built) the reference's own templates compiled in place."""
    import itsolv_oracle_lib
    return itsolv_oracle_lib.load()


@pytest.fixture(scope="session")
def ctx():
    """One context on cuda:0 for the whole session. Fails (does not skip) when the CUDA path is unavailable."""
    import torch

    import iterative_solver_b200 as pkg
    # issue the kernels on torch's current stream so that they are ordered with the test's tensor copies
    c = pkg.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    c.init_comm(0, 1, b"\0" * 128)
    yield c
    c.close()
