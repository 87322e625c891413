import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(g.PKG_DIR, "libucgb200.so")):
        g.build()
    return g.load_package()


@pytest.fixture(scope="session")
def fixtures(tmp_path_factory, pkg):
    """table / state files shared by the session"""
    from lammps_ucg_dev_b200 import synth
    d = tmp_path_factory.mktemp("ucgfix")
    out = dict(dir=str(d))
    out["table4096"] = synth.write_table_file(str(d / "ucg4096.table"), npts=4096)
    out["table1024"] = synth.write_table_file(str(d / "ucg1024.table"), npts=1024)
    out["tableR"] = synth.write_table_file(str(d / "ucgR.table"), npts=600, style="R")
    out["state"] = synth.write_state_file(str(d / "ucg.conf"))
    return out
