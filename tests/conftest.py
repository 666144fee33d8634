import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import oracle_lib  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import offline_raytracer_b200 as ort
        return ort.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a GPU test on a box without a usable device FAILS loudly (no silent fallback);
    # it is only skipped when the user did not ask for GPU tests explicitly.
    markexpr = config.getoption("-m") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        return
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here; run with -m gpu on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """builds the oracles (and the reference shim when /root/reference exists) and the sim library"""
    oracle_lib.build_oracles()
    sim = os.path.join(ROOT, "tests", "sim", "libort_sim.so")
    srcs = [os.path.join(ROOT, "tests", "sim", "sim.cpp"),
            os.path.join(ROOT, "offline_raytracer_b200", "csrc", "scene_flatten.cpp")]
    hdrs = [os.path.join(ROOT, "offline_raytracer_b200", "csrc", h) for h in ("bvh.h", "path.h", "core_math.h", "scene_flatten.h")]
    newest = max(os.path.getmtime(p) for p in srcs + hdrs)
    if not os.path.exists(sim) or os.path.getmtime(sim) < newest:
        subprocess.run(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                        "-Wno-unknown-pragmas", "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "offline_raytracer_b200", "csrc"), "-o", sim] + srcs, check=True)
    return True


@pytest.fixture(scope="session")
def oracle(built):
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def ref(built):
    if not oracle_lib.have_ref():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference at build time)")
    return oracle_lib.Ref()


@pytest.fixture(scope="session")
def data_dir(built):
    if not oracle_lib.have_data():
        pytest.skip("reference data not staged under oracle/_ref/data (needs /root/reference at build time)")
    return oracle_lib.DATA_DIR


@pytest.fixture(scope="session")
def ort():
    import offline_raytracer_b200 as m
    m.lib()     # raises loudly if the CUDA library is not built
    return m


@pytest.fixture(scope="session")
def testscene_host(ort, data_dir):
    """testscene.scn assembled by the product's own host code at 480x270 (BASELINE config 1)"""
    return ort.HostScene.load(os.path.join(data_dir, "testscene.scn"), data_dir, 480, 270)


@pytest.fixture(scope="session")
def testscene_oracle(oracle, testscene_host):
    return oracle.scene(testscene_host.world, testscene_host.root)
