import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "blender-ngp_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def small_scene():
    """8 cameras, 64x64: the Lego-shaped synthetic scene at a size the CPU oracle handles in seconds."""
    import synthetic
    return synthetic.make_lego_scene(8, 64, device="cpu", seed=0)


def scene_occupancy_bitfield(orc, margin=1):
    """Occupancy bitfield with the cells overlapping the synthetic boxes (dilated) set: grid -> bitfield through the oracle."""
    import synthetic
    grid = np.zeros(128 ** 3, np.float32)
    idx = np.arange(128 ** 3, dtype=np.uint32)

    def inv(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xc30c30c3
        x = (x | (x >> 4)) & 0x0f00f00f
        x = (x | (x >> 8)) & 0xff0000ff
        x = (x | (x >> 16)) & 0x0000ffff
        return x
    cx, cy, cz = inv(idx), inv(idx >> 1), inv(idx >> 2)
    for (bmin, bmax, _) in synthetic.lego_boxes():
        lo = [max(0, int(np.floor(bmin[d] * 128)) - margin) for d in range(3)]
        hi = [min(127, int(np.floor(bmax[d] * 128)) + margin) for d in range(3)]
        m = (cx >= lo[0]) & (cx <= hi[0]) & (cy >= lo[1]) & (cy <= hi[1]) & (cz >= lo[2]) & (cz <= hi[2])
        grid[m] = 1.0
    return grid, orc.bitfield(1, grid, 0.5)


@pytest.fixture(autouse=True)
def _release_device_temporaries(request):
    yield
    if request.node.get_closest_marker("gpu") is not None:
        import gpu_util
        gpu_util.release()
