"""GPU: the library's own bounds check (guard zones around every device allocation, CALB2_GUARD=1) over a tour of the
entry points -- the stand-in for compute-sanitizer memcheck, which is closed on this GPU pool."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_kernel_writes_outside_its_buffers(native_built):
    env = dict(os.environ, CALB2_GUARD="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "guards_check.py")], capture_output=True, text=True,
                         timeout=900, cwd=ROOT, env=env)
    print(out.stdout[-3000:])
    assert out.returncode == 0 and "PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
