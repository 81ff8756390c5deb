import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def synth_sd():
    from oracle import synth_checkpoint as S
    return S.make_checkpoint(0)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_outputs_seed0.npz"))


@pytest.fixture(scope="session")
def golden_crops():
    from oracle import synth_checkpoint as S
    from oracle.make_golden import GOLDEN_LENGTHS, SEED
    return S.synth_crops(SEED + 7, len(GOLDEN_LENGTHS), list(GOLDEN_LENGTHS))
