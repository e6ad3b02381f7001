#!/usr/bin/env python3
"""Timeline of CTA 0 of span_bulk_kernel (a -DOALSFX_SPAN_TRACE build): span_trace.py LIB chain|eax STREAMS"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "profiles"))
os.environ["OALSFX_LIB"] = sys.argv[1]
import torch, cfg_timings as ct
from oalsfxpp_b200 import ChannelFormat as F, EffectType as T
kind, streams = sys.argv[2], int(sys.argv[3])
if kind == "chain":
    r = ct.run("chain", streams, F.stereo, 48000, [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236, blocks=6, warm=3)
else:
    r = ct.run("eax", streams, F.mono, 48000, [T.eax_reverb], 200, blocks=6, warm=3)
print(r["ms_per_block_device"])
lib = ctypes.CDLL(os.path.abspath(sys.argv[1]))
W, I, P = 24, 72, 8
buf = np.zeros(W * I * P, dtype=np.int64)
lib.oalsfx_debug_span_trace(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), buf.size)
t = buf.reshape(W, I, P)
names = {0: "top", 1: "A-ready / B-start", 2: "A-done / B-done", 3: "after P", 4: "C-ready", 5: "before barrier", 6: "after barrier"}
its = slice(10, 60)
for w in range(W):
    if t[w, its, 0].min() == 0:
        continue
    d = {}
    top = t[w, its, 0].astype(np.float64)
    per_iter = np.diff(t[w, 9:61, 0]).mean()
    pts = sorted([p for p in range(P) if t[w, its, p].min() > 0], key=lambda p: np.mean(t[w, its, p] - t[w, its, 0]))
    segs = []
    for a, b in zip(pts[:-1], pts[1:]):
        segs.append(f"{a}->{b}: {np.mean(t[w, its, b] - t[w, its, a]):7.0f}")
    print(f"warp {w:2d} iter {per_iter:7.0f} cyc | " + " | ".join(segs))
