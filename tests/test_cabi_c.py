"""The C ABI from C: gcc compiles tests/cabi/cabi_check.c against include/ila_b200.h, links libila_b200.so and runs it —
header and library agree without ctypes' hand-written signature table in between.  On a box without a GPU the program
checks that the compute entries refuse to run (ILA_ERR_NOGPU); on the GPU box (-m gpu) it checks values."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import ila_b200
    so_dir = os.path.dirname(ila_b200.SO_PATH)
    exe = str(tmp_path / "cabi_check")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cabi", "cabi_check.c"), "-o", exe, "-L", so_dir, "-lila_b200", "-lm",
           f"-Wl,-rpath,{so_dir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _run(exe):
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    return r.returncode, r.stdout + r.stderr


def test_c_program_links_and_refuses_without_gpu(tmp_path, ila):
    exe = _build(tmp_path)
    rc, out = _run(exe)
    assert rc == 0, out
    assert "all checks passed" in out
    if ila.lib().ila_device_count() == 0:
        assert "ILA_ERR_NOGPU" in out


@pytest.mark.gpu
def test_c_program_values_on_gpu(tmp_path, ila):
    assert ila.lib().ila_device_count() > 0
    exe = _build(tmp_path)
    rc, out = _run(exe)
    assert rc == 0, out
    assert "K3 L[1,1] at shift 3" in out and "Bessel: x[0..5]" in out and "cholesky(S)" in out and "FAIL" not in out
