import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure): builds oracle/libuvo_oracle.so on demand."""
    from oracle import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def ctx():
    """GPU context through the C ABI; fails (does not skip) if the CUDA library cannot run."""
    import ergo_uvo_b200 as U
    c = U.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def small_stereo():
    """2 stereo frames at 640x512 (fast enough for CPU tests)."""
    from tools import synth
    return synth.StereoSequence(640, 512, n_frames=2, tex_size=1024, velocity=(0.02, 0.004, 0.002))


@pytest.fixture(scope="session")
def full_stereo():
    """3 stereo frames at BASELINE config B size, 1280x1024."""
    from tools import synth
    return synth.StereoSequence(1280, 1024, n_frames=3, tex_size=2048)


def noise_image(h, w, seed=0, channels=1):
    rs = np.random.RandomState(seed)
    from scipy import ndimage
    a = ndimage.gaussian_filter(rs.rand(h, w).astype(np.float32), 1.5)
    a = (a - a.min()) / (a.max() - a.min()) * 255
    g = a.astype(np.uint8)
    if channels == 1:
        return g
    return np.stack([np.clip(g * 0.9, 0, 255).astype(np.uint8), g, np.clip(g * 0.8 + 10, 0, 255).astype(np.uint8)], -1)
