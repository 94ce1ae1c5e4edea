import importlib.util
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def load_package():
    """Import kmer-extension_b200/ (hyphenated directory) as module ``kmer_extension_b200``."""
    name = "kmer_extension_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = ROOT / "kmer-extension_b200"
    spec = importlib.util.spec_from_file_location(name, pkg_dir / "__init__.py",
                                                  submodule_search_locations=[str(pkg_dir)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle as O
    if not O.REF_SO.exists():
        O.build(ref=True)
    if not O.REF_SO.exists():
        pytest.skip("oracle/_ref/libkmer_ref.so not built and /root/reference absent")
    return O.Ref()


@pytest.fixture(scope="session")
def corc():
    from oracle import oracle as O
    O.build(ref=False)
    return O.COracle()
