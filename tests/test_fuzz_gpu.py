"""Randomised differential run on the GPU (tests/fuzz_gpu.py): random shapes, radii across the 8-bit ring, the 16-bit
ring and the wide path, 1..400 biomes, five map kinds, batches of 1-3 neighbourhoods -- all bit for bit against the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [7, 8])
def test_randomised_differential(seed):
    run = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_gpu.py"), "150", str(seed)], cwd=ROOT,
                         capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    assert "fuzz ok" in run.stdout
