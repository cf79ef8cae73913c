import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the reference itself (tests/golden/make_golden.py)."""
    out = {}
    for name in ("golden_v1.npz", "golden_v2.npz"):  # v2: transposed weights (make_golden_v2.py)
        with np.load(ROOT / "tests" / "golden" / name) as z:
            out.update({k: z[k] for k in z.files})
    return out


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc

    orc.build()
    orc.set_mode("jit")
    return orc
