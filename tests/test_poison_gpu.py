"""
The engine, scale and JSON-API parity tests (second order, certified mode included) once more with PLF_POISON=1: every fresh device allocation is filled with
0xFF (NaN as a double, -1 as an int), so that a kernel reading a cell nobody wrote fails the parity checks
instead of passing by luck on zero-initialised memory (this caught the tile kernels reading the exponents of
tips that the skipped leaf kernel had not written).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_parity_with_poisoned_allocations():
    env = dict(os.environ, PLF_POISON="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_engine_gpu.py", "tests/test_scale_gpu.py",
                        "tests/test_json_api_gpu.py", "-q", "-x",
                        "-m", "gpu", "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
