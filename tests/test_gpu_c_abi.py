"""The C ABI without Python in the loop: examples/c_abi_demo.c (scans -> preprocess -> FPS -> Chamfer matrix ->
MMD/COV/1-NNA through raw cudaMalloc'ed pointers) checks its Chamfer entries against a host loop."""
import subprocess

import pytest

from test_abi import _build_c_demo

pytestmark = pytest.mark.gpu


def test_c_program_runs_the_path(tmp_path):
    out = subprocess.run([_build_c_demo(tmp_path)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "c_abi_demo ok" in out.stdout
