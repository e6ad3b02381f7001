#!/usr/bin/env python3
"""A/B of the span kernels against the other families (ms per 1024-frame block, device buffers).
   r02_span.py [family[:bulk]]...   e.g. auto:1 auto:0 quartet duo"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import cfg_timings as ct
from oalsfxpp_b200 import ChannelFormat as F, EffectType as T

CHAIN = [T.equalizer, T.chorus, T.echo, T.eax_reverb]
sizes_chain = [int(v) for v in os.environ.get("CHAIN_SIZES", "1024,2048,4096,8192,16384").split(",") if v]
sizes_eax = [int(v) for v in os.environ.get("EAX_SIZES", "256,1024,2048,4096,8192").split(",") if v]
for spec in sys.argv[1:] or ["auto:1", "auto:0", "quartet", "duo"]:
    fam, _, bulk = spec.partition(":")
    os.environ["OALSFX_KERNEL"] = fam
    os.environ["OALSFX_SPAN_BULK"] = bulk or "1"
    for streams in sizes_chain:
        r = ct.run(f"chain {streams} [{spec}]", streams, F.stereo, 48000, CHAIN, 236)
        print((r["config"], round(r["ms_per_block_device"], 4), round(r["algorithmic_GBps"])), flush=True)
    for streams in sizes_eax:
        r = ct.run(f"eax mono {streams} [{spec}]", streams, F.mono, 48000, [T.eax_reverb], 200)
        print((r["config"], round(r["ms_per_block_device"], 4), round(r["algorithmic_GBps"])), flush=True)
