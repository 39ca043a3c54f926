import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_iou():
    return load_golden("iou_family.npz")


@pytest.fixture(scope="session")
def golden_yolo():
    return load_golden("yolo_small.npz")


@pytest.fixture(scope="session")
def golden_nms():
    return load_golden("tv_nms.npz")


@pytest.fixture(scope="session")
def golden_map():
    return load_golden("map_small.npz")


@pytest.fixture(scope="session")
def golden_rpn():
    return load_golden("rpn_small.npz")


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


@pytest.fixture(scope="session")
def golden_grads():
    return load_golden("grads_small.npz")


@pytest.fixture(scope="session")
def golden_demo_loss():
    return load_golden("demo_loss.npz")


@pytest.fixture(scope="session")
def golden_postprocess():
    return load_golden("postprocess.npz")


@pytest.fixture(scope="session")
def golden_anchor():
    return load_golden("anchor_kmeans.npz")
