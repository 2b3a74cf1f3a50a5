"""Shared fixtures.  `-m "not gpu"` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI load/exports.
`-m gpu` runs on a B200: the parity tests proper, all through the C ABI (include/go2policy.h)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device; run with -m gpu under gpurun")


@pytest.fixture(scope="session")
def model_path():
    from go2_onnx_controller_b200 import DEFAULT_MODEL
    return DEFAULT_MODEL


@pytest.fixture(scope="session")
def policy(model_path):
    from oracle import oracle
    return oracle.load_policy(model_path)


@pytest.fixture(scope="session")
def cmodel(model_path):
    from oracle import coracle
    coracle.build()
    return coracle.CModel(model_path)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "go2_policy_golden.npz"))


@pytest.fixture(scope="session")
def golden_loop():
    return np.load(os.path.join(GOLDEN, "go2_closed_loop_golden.npz"))


@pytest.fixture(scope="session")
def wide_model_path(tmp_path_factory):
    """BASELINE.json configs[4]: synthetic 245-1024-512-256-12 ELU policy written by the test-side ONNX writer."""
    from oracle import onnx_mini
    ws, bs = onnx_mini.make_wide_policy(seed=5)
    p = tmp_path_factory.mktemp("wide") / "wide.onnx"
    p.write_bytes(onnx_mini.write_mlp_onnx(ws, bs, 1.0, batch="batch"))
    return str(p)


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    return torch
