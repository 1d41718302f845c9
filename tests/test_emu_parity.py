"""CPU-only logic check: the kernel bodies (compiled as the tests-only emulator, one thread per block)
against the oracle. The GPU parity tests proper are in test_gpu_parity.py."""
import pytest

from parity_cases import CASES


@pytest.mark.parametrize("name", sorted(CASES))
def test_emu_case(emu_lib, name):
    CASES[name](emu_lib)
