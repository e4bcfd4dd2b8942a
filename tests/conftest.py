import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """-> (meta dict, state_dict, inputs, outputs) of tests/golden/<name>.npz"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    pick = lambda p: {k[len(p):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(p)}  # noqa: E731
    return json.loads(str(z["meta"])), pick("sd/"), pick("in/"), pick("out/")


def rel_row_err(ref, got):
    """max over rows of ||d||_inf / max(1, ||ref||_inf): the sample tolerance metric."""
    ref, got = ref.detach().float().cpu(), got.detach().float().cpu()
    if ref.dim() == 1:
        ref, got = ref[:, None], got[:, None]
    num = (ref - got).abs().amax(dim=1)
    den = ref.abs().amax(dim=1).clamp(min=1.0)
    return float((num / den).max())


@pytest.fixture
def cuda_dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
