import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_sessionfinish(session, exitstatus):
    """Measured parity statistics of every comparison of the session (tests/util.py: REPORT)."""
    try:
        import json
        import util
        if not util.REPORT:
            return
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        name = os.environ.get("SVGF_PARITY_REPORT", "parity_report.json")
        with open(os.path.join(out, name), "w") as f:
            json.dump({"exitstatus": int(exitstatus), "comparisons": util.REPORT}, f, indent=1)
    except Exception as e:      # a report must never fail a run
        print("parity report not written: %s" % e)
