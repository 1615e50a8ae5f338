"""The fixed-point FFTs regenerate their twiddle tables instead of copying arm_common_tables.c (q15: floor (cos * 2^15), q31:
floor (cos * 2^31 + 0.05), clipped at +1.0). Checked entry for entry against the tables inside the reference build."""
import ctypes as C

import numpy as np
import pytest

from selenite_lite_b200 import _lib


@pytest.mark.parametrize("N", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_regenerated_twiddles_equal_the_reference_tables(ref, N):
    n = 3 * N // 4 * 2
    lib = _lib.load()
    t15 = np.zeros(n, np.int16); t31 = np.zeros(n, np.int32)
    assert lib.slb_design_twiddle_q15(N, t15.ctypes.data) == 0 and lib.slb_design_twiddle_q31(N, t31.ctypes.data) == 0
    r15 = np.ctypeslib.as_array((C.c_int16 * n).in_dll(ref.lib, "twiddleCoef_%d_q15" % N))
    r31 = np.ctypeslib.as_array((C.c_int32 * n).in_dll(ref.lib, "twiddleCoef_%d_q31" % N))
    assert np.array_equal(t15, r15) and np.array_equal(t31, r31)
