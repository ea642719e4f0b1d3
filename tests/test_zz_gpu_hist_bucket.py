"""The bucketed site histogram for genome-length alignments on the device.

These kernels were written AFTER this round's GPU budget was spent: their logic is pinned on the CPU
(tests/test_emu_hist_bucket.py executes the kernel source under an emulator and ThreadSanitizer) and
they are OPT-IN (cgpu_set_option("hist_bucket", 1); the default path is untouched), but the first run
on real hardware is the driver's.  Hence: its own process (a fault cannot poison the CUDA context of
the other tests), a time limit, last in the collection order, and xfail(strict=False) -- XPASS means
the device agrees with the verified kernels bit for bit, XFAIL that it does not (yet)."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.xfail(strict=False, reason="first execution on hardware (opt-in path, written without GPU access)")
def test_bucketed_histogram_matches_the_verified_kernels():
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "gpu_hist_bucket_check.py")], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
