import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden3d():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_step_3d.npz"))


@pytest.fixture(scope="session")
def golden2d():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_step_2d.npz"))


@pytest.fixture(scope="session")
def golden_pml3d():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_pml_3d.npz"))


@pytest.fixture(scope="session")
def golden_pml2d():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_pml_2d.npz"))


@pytest.fixture(scope="session")
def golden_laser3d():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_laser_3d.npz"))


@pytest.fixture(scope="session")
def golden_laser2d():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_laser_2d.npz"))


@pytest.fixture(scope="session")
def golden_mw2d():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_mw_2d.npz"))


@pytest.fixture(scope="session")
def golden_mw3d():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_mw_3d.npz"))
