"""Handle churn: extractors of changing geometry are created and destroyed so the allocator hands the same device
addresses (workspace, tensor maps, zero-initialised fresh buffers) to different handles; every configuration must give the
identical result each time it comes round.  (Regression: fresh buffers were zeroed on the legacy stream, which the
handles' non-blocking streams are not ordered after -- a late memset wiped tensor maps / partial results.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_handle_churn_is_deterministic():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_handles.py"), "250"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " 0 mismatches" in r.stdout
