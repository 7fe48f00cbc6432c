"""The bit-sliced FAST primitives (hyslam_b200/csrc/fast_bitslice.cuh) are host+device code: compile them with g++ and check
the 32-pixels-per-word corner test against a per-pixel scalar restatement (no GPU involved)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bitsliced_corner_test_matches_scalar(tmp_path):
    exe = str(tmp_path / "bitslice_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "bitslice_check.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches 0" in r.stdout
