import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def small_cases():
    return dict(np.load(os.path.join(GOLDEN_DIR, "small_cases.npz")))


@pytest.fixture(scope="session")
def natural():
    """Crops of the reference's own sample photographs + what the unmodified reference computes on them
    (tests/golden/make_golden_natural.py).  -> (records, arrays)"""
    with open(os.path.join(GOLDEN_DIR, "natural.json")) as f:
        meta = json.load(f)
    return meta, dict(np.load(os.path.join(GOLDEN_DIR, "natural.npz")))


def natural_input(arrays, name):
    """u8 HWC crop -> [1,3,H,W] f32 exactly as torchvision's ToTensor() feeds the reference (u8 / 255 in fp32)."""
    u8 = arrays[f"{name}_u8"]
    return np.ascontiguousarray((u8.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)[None])


def mirror_tile(x, h, w):
    """Deterministic large frame with natural content: the crop x [1,3,h0,w0] mirrored and tiled up to h x w."""
    _, _, h0, w0 = x.shape
    row = np.concatenate([x, x[..., ::-1]], axis=3)
    blk = np.concatenate([row, row[:, :, ::-1]], axis=2)
    ry, rx = -(-h // (2 * h0)), -(-w // (2 * w0))
    return np.ascontiguousarray(np.tile(blk, (1, 1, ry, rx))[:, :, :h, :w])
