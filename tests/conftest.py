import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """tests marked gpu skip (instead of failing) on a machine without a GPU, whatever -m selects"""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a B200 (no CUDA device here)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


from cosmopp_b200.synthetic import synthetic_cl  # noqa: E402,F401


@pytest.fixture(scope="session")
def oracle_api():
    from oracle import api
    api.lib()
    return api


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import cosmopp_b200 as cb
    ctx = cb.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    yield ctx
    ctx.close()
