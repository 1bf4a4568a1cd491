"""The CPU-fiber kernel suites once more under AddressSanitizer (tests/simt/build_simt.py, APK_SIMT_ASAN=1).

The device code of deposit_sorted.cu, deposit_atomic.cu, route.cu, mesh_ops.cu and bin_kmu.cu, compiled by g++ with
-fsanitize=address: an out-of-bounds access of a kernel -- a static shared-memory array, a particle column, a mesh, a
table, an output buffer -- aborts the run instead of corrupting a neighbour silently.  (Found the need for it in round 2:
route_group_kernel wrote past its output buffer when the staging buffer had overflowed, and only the SECOND P(k) on the
same plan, on 4 and 8 GPUs, showed it; under this build the old kernel is reported at the offending line.)
"""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _libasan():
    gcc = shutil.which("gcc")
    if gcc is None:
        return None
    path = subprocess.run([gcc, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    return path if os.path.isabs(path) and os.path.exists(path) else None


@pytest.mark.skipif(os.environ.get("APK_SIMT_ASAN", "0") not in ("", "0"), reason="already inside the sanitized run")
def test_kernel_sources_are_clean_under_address_sanitizer():
    lib = _libasan()
    if lib is None:
        pytest.skip("libasan not found")
    env = dict(os.environ, APK_SIMT_ASAN="1", LD_PRELOAD=lib, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    suites = ["tests/test_simt_misc_kernels.py", "tests/test_simt_deposit.py", "tests/test_simt_bin_kmu.py"]
    out = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider", *suites], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=1500)
    tail = (out.stdout + out.stderr)[-4000:]
    assert "AddressSanitizer" not in tail, tail
    assert out.returncode == 0, tail
