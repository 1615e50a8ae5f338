#!/usr/bin/env python
"""bench.py — headline metric of BASELINE.json: complex Msamples/s of the fused RX chain, whole box, + HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU; channels shard with no data-path collective -> weak scaling)

A step = one pass of the hot path (slb_rx_process_device, RX-SSB-f32 chain) over one batch of synthetic I/Q:
BASELINE configs[1], 1024 independent 48 kHz channels per GPU x 10 s (480 000 frames) = 1.97 GB in + 1.97 GB out per
GPU per step, far larger than the 126 MB L2. `value` has inputs resident in HBM; `e2e` is the same batch through the
host-buffer C-ABI call (slb_rx_process_host: pinned host memory -> H2D -> kernel -> D2H, chunked and overlapped).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "complex Msamples/s (whole box) RX chain"
UNIT = "Msamples/s"
CHANNELS_PER_GPU = 1024
FS = 48000
SECONDS = 10
FRAMES = FS * SECONDS          # 480 000 = 1250 hops of 384
BYTES_PER_SAMPLE = 8           # 4 B int16 I/Q in + 4 B int16 L/R out
WORKLOAD = "configs[1]: 1024 independent 48 kHz I/Q channels x 10 s per GPU, RX-SSB-f32 chain (overlap-save SSB demod + 2-stage biquad + AGC; filter and biquad block response evaluated as one exact integer contraction on tcgen05)"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for k, n in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def synth_on_gpu(torch, channels, frames, device, first_channel):
    """Tone + complex Gaussian noise per channel (SURVEY.md §8d amplitudes and tone placement), generated on the device."""
    import selenite_lite_b200 as slb
    g = torch.Generator(device=device); g.manual_seed(slb.signals.SEED + first_channel)
    f0 = torch.tensor([slb.channel_tone_hz(first_channel + c) for c in range(channels)], device=device, dtype=torch.float64)
    out = torch.empty((channels, frames, 2), dtype=torch.int16, device=device)
    step = 64
    n = torch.arange(frames, device=device, dtype=torch.float64)
    for c0 in range(0, channels, step):
        c1 = min(channels, c0 + step)
        ph = (2.0 * torch.pi / FS) * f0[c0:c1, None] * n[None, :]
        iq = torch.stack([torch.cos(ph), torch.sin(ph)], -1).float() * 0.25
        iq += 0.01 * torch.randn(iq.shape, device=device, generator=g)
        out[c0:c1] = torch.clamp(torch.round(iq * 32768.0), -32768, 32767).to(torch.int16)
    return out


def cpu_reference_run(threads, budget_s):
    """The reference's own CPU implementation of the path (oracle/_ref: CMSIS-DSP V1.5.3 compiled for this host; the
    port when that library is absent), channels split statically over `threads` pthreads, on a bounded sample."""
    import numpy as np
    import oracle_lib
    import selenite_lite_b200 as slb
    oracle_lib.build_oracles(want_ref=False)
    try:
        orc = oracle_lib.Oracle("ref"); kind = "reference"
    except (FileNotFoundError, OSError):
        orc = oracle_lib.Oracle("port"); kind = "port"
    p = slb.default_rx_f32_params(FS)
    prm = slb.dsp_if.params_to_dict(p, slb.default_mask(FS, p.fft_len, slb.MODE_USB))
    frames = 384 * 125                                       # 1 s per channel
    channels = threads * 8
    x = slb.synth_iq(min(channels, 8), frames)
    x = np.ascontiguousarray(np.tile(x, ((channels + x.shape[0] - 1) // x.shape[0], 1, 1))[:channels])
    orc.rx_ssb_f32_batch(prm, x[:threads], nthreads=threads)            # warm the caches / page in the library
    states, n_calls = None, 0
    t0 = time.perf_counter()
    while True:                                              # a continuing stream: state carries from call to call
        _, states = orc.rx_ssb_f32_batch(prm, x, states, nthreads=threads)
        n_calls += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s:
            break
    return {"value": n_calls * channels * frames / dt / 1e6, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "%d calls x %d channels x %d frames (1 s each) of the same chain and signal model, gcc -O2 -ffp-contract=off, %d pthreads, %.1f s"
                      % (n_calls, channels, frames, threads, dt)}, n_calls * channels * frames, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    times = []
    base = None
    budget = max(1.0, min(10.0, 60.0 / (args.warmup + args.steps)))     # whole run ends within about a minute
    for i in range(args.warmup + args.steps):
        base, n, dt = cpu_reference_run(threads, budget_s=budget)
        if i >= args.warmup:
            times.append((n, dt))
    tot_s = sum(t for _, t in times); tot_n = sum(n for n, _ in times)
    v = tot_n / tot_s / 1e6
    base["value"] = v
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / len(times), "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": WORKLOAD, "note": "CPU arm: each step is a bounded sample (%s)" % base["sample"]},
                      "cpu_baseline": base, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def run_ours(args):
    import torch
    import selenite_lite_b200 as slb
    rank, world, local = slb.shard.env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    C, T = CHANNELS_PER_GPU, FS * args.seconds
    d = slb.DspIf(C, fs=FS, chain=slb.CHAIN_RX_SSB_F32, device=local)
    x = synth_on_gpu(torch, C, T, dev, first_channel=rank * C)
    y = torch.empty_like(x)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        d.rx_process(x, y)
    barrier()
    sampler = ClockSampler(local)
    if not os.environ.get("BENCH_NO_SAMPLER"):
        sampler.start()
    time.sleep(0.3)
    launches0 = d.kernel_launches()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    torch.cuda.profiler.start()                              # ncu --profile-from-start off sees exactly the timed region
    evs[0].record()
    for i in range(args.steps):
        d.rx_process(x, y)
        evs[i + 1].record()
    barrier()
    torch.cuda.profiler.stop()
    clocks = sampler.stop()
    launches = d.kernel_launches() - launches0
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    if os.environ.get("BENCH_DEBUG"):
        print("step_ms", [round(v, 3) for v in step_ms], file=sys.stderr)
    total_ms = evs[0].elapsed_time(evs[-1])
    if dist is not None:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX); total_ms = float(t.item())
        ln = torch.tensor([launches], device=dev, dtype=torch.int64); dist.all_reduce(ln); launches = int(ln.item())
    samples_per_step = C * T * world
    value = samples_per_step * args.steps / (total_ms * 1e-3) / 1e6

    # roofline of the dominant (only) kernel: algorithmic bytes per launch / mean launch duration on this rank
    peak, peak_src = load_peaks()
    kern_ms = sum(step_ms) / len(step_ms)
    achieved = C * T * BYTES_PER_SAMPLE / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": peak_src, "kernel": "rx_ssb_tc_kernel (sl_rx_ssb_tc.cu: tcgen05.mma kind::i8 FIR + biquad/AGC epilogue)", "algorithmic_bytes_per_launch": C * T * BYTES_PER_SAMPLE,
                "note": "1 launch per step; frac vs the nominal 8 TB/s = %.3f; bound by the A-operand fetch of the SS-mode MMAs and by the epilogue's FP32 work, not by HBM (DESIGN.md §4A)" % (achieved / 8000.0)}
    traffic_file = os.path.join(ROOT, "profiles", "traffic_per_launch.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch_bench")
        except Exception:
            pass

    # end to end through the host-buffer C-ABI call
    e2e = None
    if not args.no_e2e:
        xh = torch.empty((C, T, 2), dtype=torch.int16).pin_memory(); yh = torch.empty((C, T, 2), dtype=torch.int16).pin_memory()
        xh.copy_(x); torch.cuda.synchronize()
        d2 = slb.DspIf(C, fs=FS, chain=slb.CHAIN_RX_SSB_F32, device=local)
        for _ in range(2):
            d2.rx_process_pinned(xh, yh)
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(3, min(args.steps, 5))
        for _ in range(n_e2e):
            d2.rx_process_pinned(xh, yh)                     # returns after the D2H copy of the result has landed
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        e2e = {"value": samples_per_step * n_e2e / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": C * T * 4 * world, "d2h_bytes_per_step": C * T * 4 * world,
               "steps": n_e2e, "path": "slb_rx_process_host: pinned host -> strided H2D -> rx_ssb_tc_kernel -> strided D2H, 64 MB time slices of all channels, copies and kernels on 3 streams"}
        del xh, yh, d2

    # the other chains of the library at the same width, device-resident, for context (not the headline metric)
    other = None
    if rank == 0 and world == 1 and not args.no_other:
        other = {}
        del x, y
        torch.cuda.empty_cache()
        g = torch.Generator(device=dev); g.manual_seed(1)

        def timed(fn, n=5):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        xi = torch.randint(-8000, 8000, (C, T, 2), dtype=torch.int16, device=dev, generator=g); yo = torch.empty_like(xi)
        for name, chain, call in (("rx_ssb_q15", slb.CHAIN_RX_SSB_Q15, "rx_process"), ("tx_ssb_f32", slb.CHAIN_TX_SSB_F32, "tx_process")):
            dd = slb.DspIf(C, fs=FS, chain=chain, device=local)
            ms = timed(lambda: getattr(dd, call)(xi, yo))
            other[name] = {"Msamples_per_s": C * T / (ms * 1e-3) / 1e6, "hbm_frac": C * T * BYTES_PER_SAMPLE / (ms * 1e-3) / 1e9 / peak,
                           "note": "bit-exact integer chain: rx_q15_tc_kernel (tcgen05 kind::i8 FIRs, integer epilogue)" if name == "rx_ssb_q15" else "config 3: tx_ssb_tc_kernel (tcgen05 FIR, two rails)"}
            del dd
        del xi, yo
        S, Tw = 64, 192000 * args.seconds // 768 * 768
        xw = torch.randint(-3000, 3000, (S, Tw, 2), dtype=torch.int16, device=dev, generator=g)
        yw = torch.empty((S, 64, Tw // 64, 2), dtype=torch.int16, device=dev)
        dd = slb.DspIf(S, fs=192000, chain=slb.CHAIN_CHAN64_F32, device=local)
        ms = timed(lambda: dd.chan_process(xw, yw))
        other["chan64_f32"] = {"Msamples_per_s": S * Tw / (ms * 1e-3) / 1e6, "hbm_frac": S * Tw * BYTES_PER_SAMPLE / (ms * 1e-3) / 1e9 / peak,
                               "note": "config 4: 64 x 192 kHz wideband streams -> 4096 narrowband channels"}
        del dd, xw, yw

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_baseline = cpu_reference_run(os.cpu_count() or 1, budget_s=12.0)[0]
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    if rank != 0:
        return
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "channels_per_gpu": C, "frames_per_step": T, "fs": FS, "chain": "rx_ssb_f32",
                   "l2_policy": "inputs larger than L2 (3.9 GB touched per step per GPU vs 126 MB L2)", "parallelism": "channels sharded, no collective"},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "other_chains": other}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the context measurements of the other chains")
    ap.add_argument("--seconds", type=int, default=SECONDS, help="signal seconds per channel per step (profiling runs only; the headline uses the default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
