"""Randomised adversarial parity sweep (tests/fuzz_parity.py) as a test: exact score ties, coincident / degenerate boxes,
clusters and class gaps for the three NMS front-ends, the segmented NMS and the mAP matcher, 60 seeds."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fuzz_parity_sweep():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_parity.py"), "--seeds", "60"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "fuzz ok" in r.stdout
