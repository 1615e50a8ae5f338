"""N > 1 host logic on the CPU: two gloo ranks each own a contiguous channel shard (no data-path collective), process
it independently, and the optional gather reassembles the 1-rank result byte for byte. The per-shard processor here is
the oracle (this is a test of the plumbing, not of the kernels)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, C, T, ret):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    import oracle_lib
    import selenite_lite_b200 as slb
    from test_golden import GOLD, rx_params
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = slb.shard.shard_range(C, rank, world)
    x = slb.synth_iq(hi - lo, T, first_channel=lo)                # each rank synthesises only its own channels
    orc = oracle_lib.Oracle("port")
    g = np.load(os.path.join(GOLD, "rx_ssb_f32.npz"))
    y, _ = orc.rx_ssb_f32_batch(rx_params(g, "usb"), x)
    full = slb.shard.gather_audio(torch.from_numpy(y), C)
    # spectra of the first 256 frames of the shard's channels (what slb_rx_spectrum_device computes), gathered the same way
    z = x[:, :256, 0].astype(np.float32) / 32768.0 + 1j * (x[:, :256, 1].astype(np.float32) / 32768.0)
    spec = slb.shard.gather_spectra(torch.from_numpy((np.abs(np.fft.fft(z, axis=1)) ** 2).astype(np.float32)), C)
    if rank == 0:
        ret["full"] = full.numpy().copy(); ret["spec"] = spec.numpy().copy()
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.parametrize("C", [5, 6])      # 5: shards of 2 and 3 (padded gather); 6: equal shards (one all-gather into the result)
def test_two_rank_shards_reassemble(C):
    sys.path.insert(0, HERE)
    import oracle_lib
    import selenite_lite_b200 as slb
    from test_golden import GOLD, rx_params
    T = 384 * 3
    assert [slb.shard.shard_range(5, r, 2) for r in range(2)] == [(0, 2), (2, 5)]
    assert [slb.shard.shard_range(65536, r, 8) for r in range(8)][-1] == (57344, 65536)
    oracle_lib.build_oracles(want_ref=False)
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, C, T, ret), nprocs=2, join=True)
    g = np.load(os.path.join(GOLD, "rx_ssb_f32.npz"))
    whole, _ = oracle_lib.Oracle("port").rx_ssb_f32_batch(rx_params(g, "usb"), slb.synth_iq(C, T))
    assert np.array_equal(ret["full"], whole)
    xw = slb.synth_iq(C, T)
    zw = xw[:, :256, 0].astype(np.float32) / 32768.0 + 1j * (xw[:, :256, 1].astype(np.float32) / 32768.0)
    assert np.array_equal(ret["spec"], (np.abs(np.fft.fft(zw, axis=1)) ** 2).astype(np.float32))
