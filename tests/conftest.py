import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def uv50():
    from daisyriot_b200 import scenes
    return scenes.msvc_sample_pattern(1)


@pytest.fixture(scope="session")
def cornell512():
    from daisyriot_b200 import scenes
    return scenes.cornell_box(512)


@pytest.fixture(scope="session")
def cornell2048():
    from daisyriot_b200 import scenes
    return scenes.cornell_box(2048)


@pytest.fixture(scope="session")
def fixture_scenes():
    from daisyriot_b200 import scenes
    return {n: scenes.load_scene_npz(os.path.join(GOLDEN, n + ".npz")) for n in ("cornellbox_blacklight", "colorballs")}


@pytest.fixture(scope="session")
def coeff_model(tmp_path_factory):
    from daisyriot_b200 import rgb2spec
    d = tmp_path_factory.mktemp("coeff")
    os.makedirs(d / "color_tables")
    p = str(d / "color_tables" / "srgb.coeff")
    rgb2spec.write_surrogate_table(p, 16)
    return rgb2spec.RGB2Spec.load(p), str(d)


def two_triangle_scene():
    """Two unit right triangles facing each other one unit apart (analytic known-answer case)."""
    from daisyriot_b200.scenes import Scene
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 1, 1], [1, 0, 1]], np.float32)
    VN = np.array([[0, 0, 1], [0, 0, -1]], np.float32)
    T = np.array([[0, 1, 2, 0, 0, 0], [3, 4, 5, 1, 1, 1]], np.int32)
    return Scene(V, VN, T, np.zeros(2, np.int32), [{"name": "w", "Kd": np.ones(3, np.float32), "Ke": np.zeros(3, np.float32), "Ks": np.zeros(3, np.float32)}], "two")


# ---- gather tolerance (north_star: 1e-5 relative) ------------------------------------------------------------------
REL_TOL = 1e-5
REL_FLOOR = 1e-6  # fraction of the array's largest magnitude below which the error is judged absolutely


def rel_err(got, want, floor_frac=REL_FLOOR):
    """(max TRUE relative error over entries with |want| > floor, max absolute error at or below it, floor);
    floor = floor_frac x the largest magnitude of `want`."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    floor = floor_frac * float(np.abs(want).max()) if want.size else 0.0
    big = np.abs(want) > floor
    rel = float((np.abs(got - want)[big] / np.abs(want)[big]).max()) if big.any() else 0.0
    small = float(np.abs(got - want)[~big].max()) if (~big).any() else 0.0
    return rel, small, floor


def assert_rel(got, want, what, tol=REL_TOL):
    rel, small, floor = rel_err(got, want)
    assert rel <= tol, (what, "max relative error", rel, "floor", floor)
    assert small <= tol * floor, (what, "absolute error below the floor", small, "floor", floor)
    return rel
