"""Trainer shell + text-image loss kernel + device-side IS/FID on the GPU (bodies in tools/bringup_shell.py).  Sorted
last on purpose: the kernel and network parity cases of test_gpu_parity.py run first."""
import pytest

from tools import bringup_shell

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(bringup_shell.CASES))
def test_shell(name):
    assert bringup_shell.CASES[name](), f"{name} failed"
