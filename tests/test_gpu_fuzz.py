"""A short run of the randomised parity fuzz (tests/fuzz_parity.py) inside the GPU suite; longer runs by hand."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_parity_short(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_parity.py"), "--seconds", "12", "--seed", str(seed)],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " 0 failures" in r.stdout


def test_fuzz_parity_short_radix_passes_on_small_sets():
    """LAPPER_NO_SMALL_SORT=1: sets of <= 4096 intervals sorted by the radix passes instead of the shared-memory rank sort
    (csrc/build.cu) -- the path every small set took before, kept covered."""
    env = dict(os.environ, LAPPER_NO_SMALL_SORT="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_parity.py"), "--seconds", "8", "--seed", "13"],
                       capture_output=True, text=True, cwd=ROOT, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " 0 failures" in r.stdout
