"""Host logic of the firmware ring (no GPU): the product's index arithmetic (slb_ring_plan_*) against the UNMODIFIED
reference dsp_if.c built for the host, against the plain-C port, and against the committed golden trace."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib
from selenite_lite_b200 import _lib

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class ProductRingModel:
    """numpy storage + the product's pointer logic = what the ring kernels do on the device."""

    def __init__(self, fs):
        self.lib = _lib.load()
        self.N = fs // 1000 * 8
        self.i = np.zeros(self.N, np.int16); self.q = np.zeros(self.N, np.int16)
        self.st = (C.c_uint32 * 3)(0, 0, 0)

    def write(self, is_out, hw):
        hw = np.asarray(hw, np.int16); frames = hw.size // 2
        first = self.lib.slb_ring_plan_write(self.N, is_out, C.byref(self.st), frames)
        fi, fq = hw[0::2], hw[1::2]
        for k in range(frames + 1):                       # ring_write_kernel: block, then last frame again
            s = (first + k) % self.N; src = min(k, frames - 1)
            self.i[s] = fi[src]; self.q[s] = fq[src]

    def read(self, is_out, nhw):
        frames = nhw // 2
        first = self.lib.slb_ring_plan_read(self.N, is_out, C.byref(self.st), frames)
        idx = (first + np.arange(frames)) % self.N
        out = np.empty(nhw, np.int16); out[0::2] = self.i[idx]; out[1::2] = self.q[idx]
        return out

    def ptrs(self): return tuple(self.st)


def drive(ring_w, ring_r, ptrs, sched, blocks, hw):
    outs = []
    k = 0
    for op in sched:
        if op == "w":
            ring_w(blocks[k % len(blocks)]); k += 1
        else:
            outs.append(ring_r(hw))
        outs.append(np.array(ptrs(), np.int64))
    return outs


@pytest.mark.parametrize("fs", [48000, 96000, 192000])
@pytest.mark.parametrize("direction", ["rx", "tx"])
def test_ring_logic_matches_reference_under_drift(fs, direction, rng):
    """Random producer/consumer interleavings exercise both slip branches (dsp_if.c:145-163, :266-284)."""
    try:
        ref = oracle_lib.RefRing(fs)
    except FileNotFoundError:
        ref = None
    port = oracle_lib.PortRing(fs)
    prod = ProductRingModel(fs)
    hw = fs // 1000 * 2
    blocks = rng.integers(-32768, 32768, (64, hw)).astype(np.int16)
    # matched clocks, then a fast producer, then a fast consumer
    sched = ["w", "r"] * 40 + ["w", "w", "r"] * 40 + ["w", "r", "r"] * 40 + list(rng.choice(["w", "r"], 300))
    is_out = direction == "tx"
    if is_out:
        a = drive(lambda b: prod.write(1, b), lambda n: prod.read(1, n), prod.ptrs, sched, blocks, hw)
        b = drive(port.out_write, port.out_read, port.ptrs, sched, blocks, hw)
        c = drive(ref.out_write, ref.out_read, lambda: ref.ptrs(1), sched, blocks, hw) if ref else None
    else:
        a = drive(lambda b: prod.write(0, b), lambda n: prod.read(0, n), prod.ptrs, sched, blocks, hw)
        b = drive(port.in_write, lambda n: port.in_read(n * 2), port.ptrs, sched, blocks, hw)
        c = drive(ref.in_write, lambda n: ref.in_read(n * 2), lambda: ref.ptrs(0), sched, blocks, hw) if ref else None
    for k, (x, y) in enumerate(zip(a, b)):
        assert np.array_equal(x, y), ("product vs port", k)
    if c is not None:
        for k, (x, y) in enumerate(zip(a, c)):
            assert np.array_equal(x, y), ("product vs reference", k)
        i_r, q_r = ref.iq(1 if is_out else 0)
        assert np.array_equal(prod.i, i_r) and np.array_equal(prod.q, q_r)


@pytest.mark.parametrize("fs", [48000, 96000])
def test_ring_golden_trace(fs):
    g = np.load(os.path.join(GOLD, "ring.npz"))
    blocks, rx, tx, ptrs = g["in_%d" % fs], g["rx_%d" % fs], g["tx_%d" % fs], g["ptrs_%d" % fs]
    hw = blocks.shape[1]
    for impl in ("product", "port"):
        if impl == "product":
            rin, rout = ProductRingModel(fs), ProductRingModel(fs)
            ops = (lambda b: rin.write(0, b), lambda: rin.read(0, hw), lambda b: rout.write(1, b), lambda: rout.read(1, hw),
                   lambda: rin.ptrs() + rout.ptrs())
        else:
            pin, pout = oracle_lib.PortRing(fs), oracle_lib.PortRing(fs)
            ops = (pin.in_write, lambda: pin.in_read(hw * 2), pout.out_write, lambda: pout.out_read(hw), lambda: pin.ptrs() + pout.ptrs())
        for b in range(blocks.shape[0]):
            ops[0](blocks[b]); assert np.array_equal(ops[1](), rx[b]), (impl, "rx", b)
            ops[2](blocks[b]); assert np.array_equal(ops[3](), tx[b]), (impl, "tx", b)
            assert tuple(ptrs[b]) == tuple(ops[4]()), (impl, "ptrs", b)


def test_steady_state_is_identity_with_fixed_delay():
    """matched clocks -> two start-up repeats, then a locked gap and a pure delay."""
    fs = 48000; r = ProductRingModel(fs); hw = 96
    n = np.arange(400 * 48, dtype=np.int64)
    ramp = ((n % 20000) + 1).astype(np.int16)
    outs = []
    for b in range(400):
        blk = np.empty(hw, np.int16); blk[0::2] = ramp[b * 48:(b + 1) * 48]; blk[1::2] = -ramp[b * 48:(b + 1) * 48]
        r.write(0, blk); outs.append(r.read(0, hw))
    o = np.concatenate(outs)
    i, q = o[0::2].astype(np.int64), o[1::2].astype(np.int64)
    assert np.array_equal(i, -q)                      # I/Q pairing preserved
    tail = i[2000:]
    d = np.diff(tail); d[d < 0] += 20000
    assert np.all(d == 1)                             # every later delta is exactly +1
    en, rd, wr = r.ptrs()
    assert (wr - rd) % r.N == 144                     # locked gap measured on the reference
