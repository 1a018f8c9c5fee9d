import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pi_mesh():
    from fesom2_b200 import mesh as M
    return M.load_npz_mesh(os.path.join(GOLDEN, "mesh_pi.npz"))


@pytest.fixture(scope="session")
def cav_mesh():
    """the reference's test/meshes/pi_cavity: 170 elements / 95 nodes start below layer 1 (use_cavity)"""
    from fesom2_b200 import mesh as M
    return M.load_npz_mesh(os.path.join(GOLDEN, "mesh_pi_cavity.npz"))


@pytest.fixture(scope="session")
def nw2_mesh():
    """the reference's test/meshes/neverworld2: a 60-degree periodic sector, 8578 nodes, 15 layers (many short columns per CTA)"""
    from fesom2_b200 import mesh as M
    return M.load_npz_mesh(os.path.join(GOLDEN, "mesh_neverworld2.npz"))


@pytest.fixture(scope="session")
def souf_mesh():
    from fesom2_b200 import mesh as M
    return M.load_npz_mesh(os.path.join(GOLDEN, "mesh_soufflet.npz"))


@pytest.fixture(scope="session")
def small_mesh():
    from fesom2_b200 import mesh as M
    return M.synth_mesh(31, 27, nl=20, min_layers=4)
