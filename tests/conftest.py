import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_pkg():
    """Import the package directory `gie-mapping_b200/` under the importable name gie_mapping_b200."""
    if "gie_mapping_b200" in sys.modules:
        return sys.modules["gie_mapping_b200"]
    pkg_dir = os.path.join(ROOT, "gie-mapping_b200")
    spec = importlib.util.spec_from_file_location("gie_mapping_b200", os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["gie_mapping_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gie():
    return load_pkg()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py
