import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "latticeboltzmann.jl_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices")


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    n = None
    for item in items:
        if "multigpu" in item.keywords:
            n = _gpu_count() if n is None else n
            if n < 2:
                item.add_marker(pytest.mark.skip(reason="needs >= 2 GPUs"))


@pytest.fixture(scope="session")
def oracle():
    import oracle.lbm_oracle as O
    return O


def to_host_layout(f_qyx):
    """oracle f[Q, NY, NX] (C order) -> host-API f[NX, NY, Q] (Fortran order); same bytes."""
    return np.asfortranarray(np.transpose(f_qyx, (2, 1, 0)))


def to_oracle_layout(f_xyq):
    return np.ascontiguousarray(np.transpose(f_xyq, (2, 1, 0)))


def random_populations(q_oracle, nx, ny, seed=1234, amp=0.01):
    rng = np.random.default_rng(seed)
    return np.stack([q_oracle.w[i] * (1 + amp * rng.uniform(-1, 1, (ny, nx))) for i in range(q_oracle.Q)])


def rel_max(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(np.asarray(b))), 1e-300))
