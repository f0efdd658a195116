import os
import sys


def _effective_cpus() -> int:
    """CPUs this process may really use: affinity mask and cgroup quota (os.cpu_count() sees neither)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


# Before NumPy / the oracle load their thread pools: an OpenMP or OpenBLAS team larger than the CPU quota of the box
# spin-waits on descheduled threads (round 1: the same suite took 21 s on one box and 913 s on another).
_n = str(min(8, _effective_cpus()))
for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "SRB_UPLOAD_THREADS"):
    os.environ.setdefault(_v, _n)
os.environ.setdefault("OMP_WAIT_POLICY", "passive")

import pytest  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "kat_4x5.json")) as f:
        return json.load(f)
