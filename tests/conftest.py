import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds of CPU reference time (still part of -m gpu)")


def load_package():
    """The package directory is named admm-elastic_b200 (not an identifier): load it by path."""
    name = "admm_elastic_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = os.path.join(ROOT, "admm-elastic_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    g.build()
    return load_package()


@pytest.fixture(scope="session")
def cpu():
    """The checkers: oracle (C restatement) and, when present, the compiled reference."""
    import __graft_entry__ as g
    g.build()
    import checkers
    return checkers
