"""Trainer shell + text-image loss kernel + device-side IS/FID on the GPU (bodies in tools/bringup_shell.py).  Sorted
last on purpose: the kernel and network parity cases of test_gpu_parity.py run first."""
import pytest

from tools import bringup_shell

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", [n for n in bringup_shell.CASES if n not in bringup_shell.UNVERIFIED])
def test_shell(name):
    assert bringup_shell.CASES[name](), f"{name} failed"
