import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def F():
    import frcnn_b200
    return frcnn_b200


@pytest.fixture(scope="session")
def small_model(F):
    """vgg_small + duplo config with seeded, fully randomised weights on cuda:0 (session-wide)."""
    import torch
    from oracle import model as OM
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    m = F.vgg_small(F.duplo_cfg)
    p = OM.init_params(OM.VGG_SMALL, OM.CFG_DUPLO, seed=0, randomize_aux=True)
    m.load_params(p)
    m.oracle_params = p
    return m
