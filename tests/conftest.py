import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def teapot():
    import numpy as np
    d = np.load(os.path.join(ROOT, "tests", "golden", "teapot.npz"))
    return d["vertices"], d["faces"]


@pytest.fixture(scope="session")
def golden_images():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_images.npz"))
    sil = np.unpackbits(g["silhouette"]).reshape(256, 256).astype(np.float32)
    return {"silhouette": sil, "depth_u8": g["depth_u8"], "rasterize1_u8": g["rasterize1_u8"],
            "rasterize2_u8": g["rasterize2_u8"]}


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
        return json.load(f)
