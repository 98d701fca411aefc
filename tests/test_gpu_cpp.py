"""GPU: the header-only C++ mirror of the reference's Column / RowIndex / Groupby / group()
(include/dtb200.hpp) over the C-ABI, run as a native binary."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_mirror():
    exe = os.path.join(ROOT, "datatable_b200", "lib", "host_mirror_test")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
