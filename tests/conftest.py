import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


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
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_sessionfinish(session, exitstatus):
    """Every single-tick parity test records {fixture: total, loose, worst_tight, worst_loose} (common.PARITY_LOG); a session that
    ran on a GPU writes them to profiles/parity_r02_gpu.json (and gpurun_out/, which is what travels back from the box), a
    CPU session (the host build of the same headers) to profiles/parity_r02_host.json."""
    import json

    try:
        import common
    except Exception:
        return
    if not common.PARITY_LOG:
        return
    kind = "gpu" if (_has_gpu() and any("gpu" in i.keywords for i in session.items)) else "host"
    doc = dict(kind=kind, tol_tight=common.TOL_TIGHT, tol_contact=common.TOL_CONTACT, allow_contact_frac=common.ALLOW_CONTACT_FRAC,
               fixtures=common.PARITY_LOG,
               total=sum(v["total"] for v in common.PARITY_LOG.values()), loose=sum(v["loose"] for v in common.PARITY_LOG.values()))
    for d in (os.path.join(ROOT, "profiles"), os.path.join(ROOT, "gpurun_out")):
        try:
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, f"parity_r02_{kind}.json"), "w") as f:
                json.dump(doc, f, indent=1)
        except OSError:
            pass
