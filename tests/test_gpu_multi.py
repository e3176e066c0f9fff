"""In-process multi-GPU sharding (ngpus = G) of every host entry point reproduces ngpus = 1 bit for bit.
Needs >= 2 GPUs (skipped on a single-GPU box); the checks live in scripts/multi_gpu_check.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_all_host_entries_shard_identically(ila):
    from ila_b200 import _lib
    if _lib.lib().ila_device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "multi_gpu_check.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
