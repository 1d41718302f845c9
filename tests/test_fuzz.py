"""Randomised parity: 30 seeded random configurations per tier (tools/fuzz_parity.py runs more)."""
import pytest

from fuzz_cases import random_case


@pytest.mark.parametrize("seed", range(9000, 9030))
def test_fuzz_emu(emu_lib, seed):
    random_case(emu_lib, seed)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(9100, 9160))
def test_fuzz_gpu(cuda_lib, seed):
    random_case(cuda_lib, seed)
