#!/usr/bin/env python
"""Live feeder (slb_live_*) against the blocking replay (slb_feeder_run): sustained real-time factor and the added latency per chunk.
One JSON line per configuration. Profiling aid, not the bench contract.
   real_time_factor = audio seconds processed per wall second (all channels in parallel); latency = push -> results in host memory."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import selenite_lite_b200 as slb

ap = argparse.ArgumentParser()
ap.add_argument("--channels", type=int, default=1024)
ap.add_argument("--chunks", type=int, default=60)
args = ap.parse_args()
C = args.channels
for chain, name in ((slb.CHAIN_RX_SSB_F32, "rx_ssb_f32"), (slb.CHAIN_RX_SSB_Q15, "rx_ssb_q15"), (slb.CHAIN_PASS, "pass")):
    for ticks in (8, 40, 200):
        B = 48; n = ticks * B
        x = slb.synth_iq(C, n); pc = np.zeros((C, n, 2), np.int16)
        d = slb.DspIf(C, chain=chain)
        # blocking replay
        for _ in range(3): d.feeder_run(x, pc)
        t0 = time.perf_counter()
        for _ in range(args.chunks): d.feeder_run(x, pc)
        t_block = (time.perf_counter() - t0) / args.chunks
        # pipeline, depth 3
        lv = d.live_open(ticks, 3)
        lats = []
        for _ in range(3):
            lv.push(x, pc)
        t0 = time.perf_counter()
        for _ in range(args.chunks):
            u, v, lat = lv.pop(); lats.append(lat); lv.push(x, pc)
        t_live = (time.perf_counter() - t0) / args.chunks
        while lv.in_flight(): lv.pop()
        lv.close()
        # latency of a lone chunk (nothing else in flight)
        lv = d.live_open(ticks, 2); lone = []
        for _ in range(10):
            lv.push(x, pc); lone.append(lv.pop()[2])
        lv.close()
        print(json.dumps({"chain": name, "channels": C, "chunk_ms": ticks, "blocking_ms_per_chunk": round(t_block * 1e3, 3), "live_ms_per_chunk": round(t_live * 1e3, 3),
                          "real_time_factor_blocking": round(ticks * 1e-3 / t_block, 2), "real_time_factor_live": round(ticks * 1e-3 / t_live, 2),
                          "latency_us_lone_chunk_median": round(float(np.median(lone)), 1), "latency_us_in_pipeline_median": round(float(np.median(lats)), 1),
                          "Msamples_per_s_live": round(C * n / t_live / 1e6, 1)}), flush=True)
