import os
import shutil
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def workload_root():
    """Scratch asset roots for the procedural workloads, one directory per (name, size)."""
    base = tempfile.mkdtemp(prefix="ptb_tests_")
    cache = {}

    def get(name, **kw):
        from pathtracerwithcuda_b200 import procedural as pr
        key = (name, tuple(sorted(kw.items())))
        if key not in cache:
            root = os.path.join(base, "%s_%d" % (name, len(cache)))
            cache[key] = (root, pr.make_workload(root, name, **kw))
        return cache[key]

    yield get
    shutil.rmtree(base, ignore_errors=True)


def has_gpu():
    import pathtracerwithcuda_b200 as ptb
    return ptb.device_count() > 0
