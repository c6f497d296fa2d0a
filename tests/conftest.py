import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    return np.load(path)


@pytest.fixture(autouse=True)
def _seed():
    # every test builds its own throw-away rotation tensors: do not let the "caller builds R per call" heuristic of
    # qutlass_b200._rotation_hint (stop synchronising after 64 dead tensors) leak from one test into the next
    qb = sys.modules.get("qutlass_b200")
    if qb is not None and hasattr(qb, "_ROT_DEAD"):
        qb._ROT_DEAD[0] = 0
    np.random.seed(0)
    try:
        import torch
        torch.manual_seed(0)
    except Exception:
        pass


@pytest.fixture
def b200q_env(monkeypatch):
    """Set / unset a B200Q_* switch of libb200q for one test.  The library caches its environment at first use
    (include/b200q.h: b200q_reload_env), so every change is followed by a reload, and so is the clean-up."""
    from qutlass_b200 import _lib

    def set_(name, value):
        if value is None:
            monkeypatch.delenv(name, raising=False)
        else:
            monkeypatch.setenv(name, str(value))
        _lib.reload_env()

    yield set_
    monkeypatch.undo()
    _lib.reload_env()
