"""The persistent kernels hand work between the roles of a CTA (bulk-copy producer, converters, MMA issuer, epilogue sets; FFT and
recurrence warps; channelizer producer and AGC warps) through mbarriers and named barriers, and between CTAs through look-back words. A protocol that only holds for the usual timing passes every functional test: round 2 found three such
faults in the q15 kernel (a hang at 8192 channels, and two that compute-sanitizer's timing exposed). The stress build
(`-DSL_TC_STRESS`, selenite_lite_b200/lib/libselenite_b200_stress.so, built by `__graft_entry__.build()`) sleeps a pseudo-random time
of up to ~16 us at every arrive / wait / named barrier / look-back publication (csrc/sl_stress.cuh), so roles and warps drift apart
by several tiles' worth of time. The parity tests of every chain must pass unchanged on it (the pre-fix q15 kernel fails 9 of 17 there)."""
import os
import subprocess
import sys

import pytest

from selenite_lite_b200 import build as _build

pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["tests/test_gpu_rx_ssb_tc.py", "tests/test_gpu_rx_ssb_f32.py", "tests/test_gpu_tx_ssb_f32.py", "tests/test_gpu_rx_ssb_q15.py",
         "tests/test_gpu_rx_fm_f32.py", "tests/test_gpu_chan64_f32.py", "tests/test_gpu_ring.py", "tests/test_gpu_scale.py"]


@pytest.mark.timeout(900, method="thread")
def test_parity_tests_pass_on_the_stress_build():
    if not os.path.exists(_build.STRESS_LIB_PATH):
        pytest.skip("stress build not present (python __graft_entry__.py builds it)")
    env = dict(os.environ, SELENITE_B200_LIB=_build.STRESS_LIB_PATH)
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider", "--timeout", "600"] + FILES,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=850)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
