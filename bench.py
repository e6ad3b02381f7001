#!/usr/bin/env python3
"""bench.py -- channel-samples/s of the 4-slot EAX reverb chain (BASELINE.json metric).

One "step" = one 1024-frame block of the hot path (oalsfx_engine_mix: source encode + dry mix +
equalizer + chorus + echo + EAX reverb + output interleave, fused) over ALL streams of the rank.

Workload at N=1 = BASELINE.json configs[4] on one GPU ("cfg4": 65 536 independent stereo 48 kHz
streams, static default parameters) -- the configuration the metric and the north_star target are
quoted on.  With N>1 every rank runs the same 65 536 streams on its own GPU (streams shard with no
data-path collective: "weak" scaling); `value` is the whole-job aggregate.

  value  device-resident I/O (stream-major [stream][frame][channel] fp32 in HBM), CUDA events on the
         launching stream, max over ranks
  e2e    the same call with HOST buffers (pinned): H2D + kernel + D2H inside the timed region
  roofline      algorithmic HBM bytes (236 B/frame, SURVEY.md 8d / DESIGN.md) / event time of the
                fused kernel, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, compiled from the unmodified
                sources) on all host cores, bounded sample

`--impl reference` times that CPU implementation instead (bench.py's only other use of oracle/).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STREAMS_PER_GPU = 65536
CHANNEL_FORMAT = 2  # stereo
CHANNELS = 2
RATE = 48000
BLOCK = 1024
CHAIN = (7, 1, 6, 11)  # equalizer, chorus, echo, eax_reverb (oalsfxpp::EffectType values)
BYTES_PER_FRAME = 236  # I/O 16 + reverb rings 192 + echo 12 + chorus 16
METRIC = "channel-samples/s, 4-slot EAX reverb chain @48 kHz"
UNIT = "channel-samples/s"


def workload_config(streams):
    return {
        "workload": "cfg4: 4-slot chain equalizer+chorus+echo+eax_reverb (defaults), "
                    f"{streams} independent stereo 48 kHz streams per GPU, {BLOCK}-frame blocks",
        "streams_per_gpu": streams, "channels": CHANNELS, "rate": RATE, "block_frames": BLOCK,
        "layout": "stream-major [stream][frame][channel] fp32",
        "l2": "per-step working set (>10 GB of delay lines + 1 GB of I/O) exceeds the 126 MB L2; no flush needed",
        "parallelism": "streams sharded per GPU, no data-path collective",
    }


# ---- CPU reference leg (test infrastructure under oracle/) ------------------------------------------
def _load_cpu_reference():
    """oracle/_ref (the compiled, unmodified reference) if present, else the restatement."""
    for kind, rel in (("reference", "oracle/_ref/liboalsfx_ref_fast.so"), ("reference", "oracle/_ref/liboalsfx_ref.so"),
                      ("port", "oracle/_build/liboalsfx_oracle.so")):
        path = os.path.join(ROOT, rel)
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.orc_bench.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_int)] + [ctypes.c_int] * 3 + \
                [ctypes.c_uint32, ctypes.POINTER(ctypes.c_double)]
            lib.orc_bench.restype = ctypes.c_double
            return kind, lib
    if os.path.isdir("/root/reference/src") or os.path.exists(os.path.join(ROOT, "oracle", "oalsfx_oracle.cpp")):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=False)
        return _load_cpu_reference() if os.path.exists(os.path.join(ROOT, "oracle/_build/liboalsfx_oracle.so")) else (None, None)
    return None, None


def cpu_run(lib, threads, streams, blocks):
    types = (ctypes.c_int * 4)(*CHAIN)
    checksum = ctypes.c_double()
    secs = lib.orc_bench(threads, streams, CHANNEL_FORMAT, RATE, types, 4, BLOCK, blocks, 0x0A15F00D,
                         ctypes.byref(checksum))
    return streams * CHANNELS * BLOCK * blocks / secs, secs


def cpu_baseline(target_seconds=12.0):
    kind, lib = _load_cpu_reference()
    if lib is None:
        return None
    threads = os.cpu_count() or 1
    per_stream_blocks = 64
    rate, _ = cpu_run(lib, threads, 2 * threads, per_stream_blocks)  # probe (same shape) to size the sample
    streams = max(threads, int(rate * target_seconds / (CHANNELS * BLOCK * per_stream_blocks)) // threads * threads)
    value, secs = cpu_run(lib, threads, streams, per_stream_blocks)
    return {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{streams} streams x {per_stream_blocks} blocks of {BLOCK} frames of the same 4-slot stereo chain, "
                      f"one Api instance per thread at a time, {secs:.1f} s wall"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, lib = _load_cpu_reference()
    if lib is None:
        print(json.dumps({"impl": "reference", "unavailable": "no CPU checker library could be built"}))
        return 0
    threads = os.cpu_count() or 1
    blocks = 32
    rate, _ = cpu_run(lib, threads, 2 * threads, blocks)
    # each step: a bounded sample sized for ~2 s so K steps + W warm-ups stay within minutes
    streams = max(threads, int(rate * 2.0 / (CHANNELS * BLOCK * blocks)) // threads * threads)
    for _ in range(args.warmup):
        cpu_run(lib, threads, streams, blocks)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_run(lib, threads, streams, blocks)
    secs = time.perf_counter() - t0
    value = args.steps * streams * CHANNELS * BLOCK * blocks / secs
    sample = f"per step: {streams} streams x {blocks} blocks of {BLOCK} frames, {threads} threads"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference", "config": workload_config(STREAMS_PER_GPU),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---- clocks ----------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled every few ms WHILE the timed region runs (NVML; falls back
    to the nvidia-smi query of B200_PROFILING.md when pynvml is unavailable)."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.sm, self.max_sm, self.reasons = index, [], None, set()
        self._stop, self._thread, self._nvml = threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)))
        bits = n.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
        for name, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                           ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                           ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                           ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
            if bits & getattr(n, attr, 0):
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(",")]
        if len(parts) >= 6:
            self.sm.append(float(parts[0]))
            self.max_sm = float(parts[1])
            for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
                if parts[2 + i].lower().startswith("active"):
                    self.reasons.add(name)

    def _loop(self):
        while not self._stop.is_set():
            try:
                if self._nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.004 if self._nvml else 0.1)

    def __enter__(self):
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no clock samples"]}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "nvml" if self._nvml else "nvidia-smi"}


# ---- our arm ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU, help="streams per GPU (default: the cfg4 size)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import oalsfxpp_b200 as ox

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # Rank 0 prints ONE JSON line on stdout and nothing else: anything a library writes to file descriptor 1
    # (NCCL's version banner does, whatever NCCL_DEBUG says) is sent to stderr instead.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    S, C, F = args.streams, CHANNELS, BLOCK

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = ox.Engine(S, CHANNEL_FORMAT, RATE, 4, device=local_rank)
    for slot, fx in enumerate(CHAIN):
        eng.set_effect(slot, fx)
    gen = torch.Generator(device=dev)
    gen.manual_seed(0x0A15F00D + rank)
    x = torch.rand((S, F, C), device=dev, generator=gen, dtype=torch.float32) - 0.5  # white noise in [-0.5, 0.5)
    y = torch.empty_like(x)
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        eng.mix(x, y, frames=F, stream=stream)

    # -- device-resident throughput ---------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = eng.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clocks:
        ev[0].record()
        for k in range(args.steps):
            step_device()
            ev[k + 1].record()
        barrier()
    launches = eng.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * S * C * F * args.steps / (total_ms_max * 1e-3)
    assert bool(torch.isfinite(y).all()), "non-finite output"

    # -- the optional all-streams output bus (BASELINE config 4): per-GPU deterministic partial sum of the block
    #    just mixed + one NCCL all-reduce of the [frames][channels] bus over NVLink.  Reported beside the headline
    #    number, not inside its timed region (the reference has no such bus).
    bus = torch.zeros((F, C), device=dev, dtype=torch.float32)
    bus_steps = max(3, min(args.steps, 10))

    def step_bus():
        eng.reduce_bus(F, y, bus, stream=stream)
        if world > 1:
            dist.all_reduce(bus, op=dist.ReduceOp.SUM)

    step_bus()
    barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for _ in range(bus_steps):
        step_bus()
    b1.record()
    barrier()
    t = torch.tensor([b0.elapsed_time(b1) / bus_steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    bus_info = {"ms_per_block": float(t.item()),
                "what": "oalsfx_engine_reduce_bus (two coalesced passes over the block's output)" +
                        (f" + NCCL all_reduce of [{F}][{C}] fp32 over {world} GPUs" if world > 1 else ""),
                "bytes_read_per_gpu": S * F * C * 4}

    # -- end to end through the C ABI with host buffers ---------------------------------------------------
    e2e = None
    if not args.skip_e2e:
        hx = torch.empty((S, F, C), dtype=torch.float32).pin_memory()
        hx.copy_(x.cpu())
        hy = torch.empty((S, F, C), dtype=torch.float32).pin_memory()
        hxn, hyn = hx.numpy(), hy.numpy()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            eng.mix(hxn, hyn, frames=F)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.mix(hxn, hyn, frames=F)  # H2D + fused kernel + D2H + sync inside the call
        barrier()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * S * C * F * e2e_steps / float(t.item()), "unit": UNIT,
               "h2d_bytes_per_step": S * F * C * 4, "d2h_bytes_per_step": S * F * C * 4, "steps": e2e_steps,
               "note": "oalsfx_engine_mix with pinned host buffers: H2D + kernel + D2H inside the timed region"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    median_ms = kernel_ms[len(kernel_ms) // 2]
    achieved = BYTES_PER_FRAME * S * F / (median_ms * 1e-3) / 1e9
    traffic = None
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        traffic = json.load(open(traffic_path)).get("dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(S),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "duo_kernel<2,FxEqualizer,FxModDelay,FxEcho,FxReverb> (kDuoChainStereo): "
                                                    "one launch = one step = the whole fused path",
                     "algorithmic_bytes_per_launch": BYTES_PER_FRAME * S * F, "kernel_ms_median": median_ms,
                     "peak_source": peak_src},
        "clocks": clocks.summary(),
        "gpu_launches": launches,
        "bus": bus_info,
        "device_bytes": eng.device_bytes,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.skip_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
