#!/usr/bin/env python3
"""Many parameter classes changing every block: host and device ms per block.
   r02_classes.py STREAMS RUN   (RUN = streams per class, multiple of 32: class-per-tile launch)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import oalsfxpp_b200 as ox
from oalsfxpp_b200 import ChannelFormat as F, EffectType as T
S, RUN = int(sys.argv[1]), int(sys.argv[2])
BLOCK = 1024
names = ox.reverb_preset_names()
presets = [ox.reverb_preset(g, n) for g, n in names]
dev = torch.device("cuda:0")
x = torch.rand(S, BLOCK, 2, device=dev) - 0.5
y = torch.empty_like(x)
stream = torch.cuda.current_stream().cuda_stream
ncls = S // RUN
with ox.Engine(S, F.stereo, 48000, 4) as eng:
    for i, t in enumerate([T.equalizer, T.chorus, T.echo]):
        eng.set_effect(i, t)
    host_set, host_mix, devms = [], [], []
    for b in range(12):
        t0 = time.perf_counter()
        for c in range(ncls):
            eng.set_effect(3, T.eax_reverb, presets[(c + b) % len(presets)], first_stream=c * RUN, n_streams=RUN)
        t1 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.mix(x, y, frames=BLOCK, stream=stream)
        t2 = time.perf_counter()
        e1.record()
        torch.cuda.synchronize()
        if b >= 3:
            host_set.append((t1 - t0) * 1e3); host_mix.append((t2 - t1) * 1e3); devms.append(e0.elapsed_time(e1))
    print({"streams": S, "classes": ncls, "host_set_effect_ms": float(np.median(host_set)), "host_mix_call_ms": float(np.median(host_mix)),
           "device_ms": float(np.median(devms)), "kernel": eng.last_kernel, "launches_per_block": eng.launch_count / 12})
