#!/usr/bin/env python3
"""Secondary measurements: ms per 1024-frame block of every BASELINE.json configuration on ONE GPU,
device-resident buffers, CUDA events on the launching stream (the headline cfg4 line is bench.py's).

    python profiles/cfg_timings.py > profiles/r01_cfg_timings.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oalsfxpp_b200 as ox  # noqa: E402
from oalsfxpp_b200 import ChannelFormat as F, EffectType as T  # noqa: E402

BLOCK = 1024


def run(name, streams, fmt, rate, chain, bytes_per_frame, schedule=None, blocks=24, warm=4, setup=None):
    channels = ox.channel_count(fmt)
    dev = torch.device("cuda:0")
    x = (torch.rand(streams, BLOCK, channels, device=dev) - 0.5)
    y = torch.empty_like(x)
    stream = torch.cuda.current_stream().cuda_stream
    with ox.Engine(streams, fmt, rate, len(chain)) as eng:
        for i, t in enumerate(chain):
            eng.set_effect(i, t)
        if setup:
            setup(eng)
        host_s = 0.0
        times = []
        for b in range(warm + blocks):
            t0 = time.perf_counter()
            if schedule:
                schedule(eng, b)
            t1 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.mix(x, y, frames=BLOCK, stream=stream)
            e1.record()
            torch.cuda.synchronize()
            if b >= warm:
                host_s += t1 - t0
                times.append(e0.elapsed_time(e1))
        launches = eng.launch_count
    ms = float(np.median(times))
    frames = streams * BLOCK
    return {"config": name, "streams": streams, "channels": channels, "rate": rate, "ms_per_block_device": ms,
            "host_param_update_ms_per_block": host_s / blocks * 1e3,
            "channel_samples_per_s": frames * channels / (ms * 1e-3),
            "algorithmic_GBps": bytes_per_frame * frames / (ms * 1e-3) / 1e9, "launches": launches}


def cfg2_schedule(eng, b):
    rv = ox.default_props(T.eax_reverb, gain_=0.20 + 0.10 * ((b % 4) / 4.0), reflections_delay_=(0.012 if (b // 16) % 2 else 0.007))
    eq = ox.default_props(T.equalizer, mid1_gain_=1.0 + 0.5 * ((b % 8) / 8.0))
    eng.set_effect(3, T.eax_reverb, rv)
    eng.set_effect(0, T.equalizer, eq)


TABLE_CHUNK = [0]  # streams per run of one preset (0: n / 904, i.e. mixed tiles)


def all_presets(eng):
    """Slot 3: stream s gets reverb preset s mod 113 -- 113 parameter classes in one engine (table mode)."""
    names = ox.reverb_preset_names()
    per = [ox.reverb_preset(g, n) for g, n in names]
    for i, p in enumerate(per):
        # streams i, i + 113, ... : set in strided runs of one stream is slow from Python; use blocks of streams instead
        pass
    n = eng.num_streams
    chunk = TABLE_CHUNK[0] or max(1, n // (len(per) * 8))
    s = 0
    k = 0
    while s < n:
        m = min(chunk, n - s)
        eng.set_effect(3, T.eax_reverb, per[k % len(per)], first_stream=s, n_streams=m)
        s += m
        k += 1


_PRESETS = []


def rotate_presets(eng, b):
    """Every tile (32 streams) its own reverb preset, and every one of them a different preset each block: the whole
    property set changes, so every class is re-derived, the class table re-uploaded, and every stream cross-fades."""
    if not _PRESETS:
        _PRESETS.extend(ox.reverb_preset(g, n) for g, n in ox.reverb_preset_names())
    for c in range(eng.num_streams // 32):
        eng.set_effect(3, T.eax_reverb, _PRESETS[(c + b) % len(_PRESETS)], first_stream=c * 32, n_streams=32)


def main():
    out = [
        run("cfg1: EAX reverb, 1024 mono streams", 1024, F.mono, 48000, [T.eax_reverb], 200),
        run("cfg2: EQ+chorus+echo+EAX, 4096 stereo streams, per-block parameter changes", 4096, F.stereo, 48000,
            [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236, schedule=cfg2_schedule, blocks=48),
        run("cfg2 without parameter changes", 4096, F.stereo, 48000, [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236),
        run("cfg3: flanger+ring modulator+distortion+compressor, 16384 mono streams @96 kHz", 16384, F.mono, 96000,
            [T.flanger, T.ring_modulator, T.distortion, T.compressor], 24),
        run("cfg4: EQ+chorus+echo+EAX, 65536 stereo streams", 65536, F.stereo, 48000,
            [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236, blocks=10, warm=3),
        # signatures without a fused kernel of their own: the relay pipeline (one launch, a warp per slot)
        run("relay: echo+EAX, 65536 stereo streams", 65536, F.stereo, 48000, [T.echo, T.eax_reverb], 220, blocks=10, warm=3),
        run("relay: EAX+chorus+reverb+compressor, 4096 stereo streams", 4096, F.stereo, 48000,
            [T.eax_reverb, T.chorus, T.reverb, T.compressor], 416),
        run("relay: distortion+flanger+equalizer, 65536 stereo streams", 65536, F.stereo, 48000,
            [T.distortion, T.flanger, T.equalizer], 32, blocks=10, warm=3),
        # more than two channels / active send shelf filters: the wide and the filter-carrying relay kernels, one launch
        # (OALSFX_KERNEL=generic gives the former path: one exact pass per slot)
        run("relay wide: cfg4 chain on 5.1, 65536 streams", 65536, F.five_point_one, 48000,
            [T.equalizer, T.chorus, T.echo, T.eax_reverb], 268, blocks=8, warm=3),
        run("relay sf: cfg4 chain, stereo, send shelf filters active, 65536 streams", 65536, F.stereo, 48000,
            [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236, blocks=8, warm=3,
            setup=lambda eng: eng.set_sends(direct=(0.8, 0.5, 1.0), aux=[(0.7, 1.0, 0.4), (1.0, 1.0, 1.0), (1.0, 0.25, 0.5), (0.9, 0.3, 0.6)])),
        # every stream its own reverb preset (113 parameter classes): table mode, coefficient blocks from HBM
        run("table mode: cfg4 chain, 113 reverb presets over 16384 stereo streams", 16384, F.stereo, 48000,
            [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236, blocks=10, warm=3, setup=all_presets),
    ]
    # the same with every run of 32 streams (one tile) its own preset: ONE class-per-tile launch (duo_multi_kernel)
    TABLE_CHUNK[0] = 32
    out.append(run("class per tile: cfg4 chain, 113 reverb presets in runs of 32 over 16384 stereo streams", 16384, F.stereo, 48000,
                   [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236, blocks=10, warm=3, setup=all_presets))
    out.append(run("class per tile: cfg4 chain, 113 reverb presets in runs of 32 over 65536 stereo streams", 65536, F.stereo, 48000,
                   [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236, blocks=10, warm=3, setup=all_presets))
    # arbitrary assignment (stream s -> preset s mod 113: every tile would hold 32 classes) ordered by oalsfx_plan_placement:
    # class-pure tiles, one launch; the engine is created with the padded stream count and the padding streams are processed
    for users in (16384, 65536):
        index, total, classes = ox.plan_placement(np.arange(users, dtype=np.int32) % 113)
        per = [ox.reverb_preset(g, n) for g, n in ox.reverb_preset_names()]

        def place(eng, classes=classes, per=per):
            for label, first, span in classes:
                eng.set_effect(3, T.eax_reverb, per[label % len(per)], first_stream=first, n_streams=span)

        r = run("placed: cfg4 chain, 113 reverb presets arbitrarily assigned over %d stereo streams, ordered by oalsfx_plan_placement (%d engine streams)" % (users, total),
                total, F.stereo, 48000, [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236, blocks=10, warm=3, setup=place)
        r["caller_streams"] = users
        out.append(r)
    # ... and all of them changing every block (`host_param_update_ms_per_block` = the set_effect calls from Python; the
    # derivation of the dirty classes and the table upload are inside the mix call, i.e. inside ms_per_block_device)
    out.append(run("class per tile: cfg4 chain, 512 classes over 16384 stereo streams, every class a new preset every block", 16384, F.stereo, 48000,
                   [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236, schedule=rotate_presets, blocks=10, warm=3))
    out.append(run("class per tile: cfg4 chain, 4096 classes over 131072 stereo streams, every class a new preset every block", 131072, F.stereo, 48000,
                   [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236, schedule=rotate_presets, blocks=6, warm=3))
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "block_frames": BLOCK, "results": out}, indent=1))


if __name__ == "__main__":
    main()
