#!/usr/bin/env python3
"""A/B of the scan equalizer against the exact kernel: few streams, long blocks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "profiles"))
import cfg_timings as ct
from oalsfxpp_b200 import ChannelFormat as F, EffectType as T
ct.BLOCK = 2048
for scan in ("0", "1"):
    os.environ["OALSFX_SCAN"] = scan
    for streams in (32, 256, 1024, 4096):
        r = ct.run(f"equalizer {streams} scan={scan}", streams, F.stereo, 48000, [T.equalizer], 16)
        print((r["config"], round(r["ms_per_block_device"], 4)), flush=True)
