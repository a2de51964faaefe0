import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The reference is fp32 end to end.  PyTorch's cuDNN convolutions default to TF32 (1e-3 relative), which
    # the stock-PyTorch ModifiedGATLayer (two Conv1d, train.py:83-84) would silently pick up on CUDA.
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="session")
def lib_built():
    """libmgs.so must exist (built by __graft_entry__.build()); build it if the toolchain is here."""
    from m_gat_graphsage_b200 import _build
    if not _build.is_current():
        _build.build()
    return _build.LIB_PATH


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def pytest_sessionfinish(session, exitstatus):
    """GPU runs leave the measured parity errors (both metrics of tests/_parity.py) in gpurun_out/."""
    try:
        import _parity
        _parity.dump_report()
    except Exception:
        pass
