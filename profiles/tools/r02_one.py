#!/usr/bin/env python3
"""One configuration, a few blocks (for ncu): r02_one.py chain|eax STREAMS [blocks]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import cfg_timings as ct
from oalsfxpp_b200 import ChannelFormat as F, EffectType as T
kind, streams = sys.argv[1], int(sys.argv[2])
blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 6
if kind == "chain":
    r = ct.run(f"chain {streams}", streams, F.stereo, 48000, [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236, blocks=blocks, warm=3)
elif kind == "chain51":
    r = ct.run(f"chain 5.1 {streams}", streams, F.five_point_one, 48000, [T.equalizer, T.chorus, T.echo, T.eax_reverb], 268, blocks=blocks, warm=3)
elif kind == "chainsf":
    r = ct.run(f"chain sf {streams}", streams, F.stereo, 48000, [T.equalizer, T.chorus, T.echo, T.eax_reverb], 236, blocks=blocks, warm=3,
               setup=lambda eng: eng.set_sends(direct=(0.8, 0.5, 1.0), aux=[(0.7, 1.0, 0.4), (1.0, 1.0, 1.0), (1.0, 0.25, 0.5), (0.9, 0.3, 0.6)]))
elif kind == "cfg3":
    r = ct.run(f"cfg3 {streams}", streams, F.mono, 96000, [T.flanger, T.ring_modulator, T.distortion, T.compressor], 24, blocks=blocks, warm=3)
else:
    r = ct.run(f"eax mono {streams}", streams, F.mono, 48000, [T.eax_reverb], 200, blocks=blocks, warm=3)
print(r)
