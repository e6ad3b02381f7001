#!/usr/bin/env python3
"""Generate tests/golden/golden.npz from the COMPILED, UNMODIFIED reference (oracle/_ref).

The reference ships no golden vectors, so these are outputs of the reference itself
run in the build container: for every case of the quick matrix (tests/cases.py) the SHA-256 of the
output bytes, and for a representative subset the full output arrays.  Run:

    python tests/golden/make_golden.py        # needs /root/reference (or a prebuilt oracle/_ref)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
import harness as H  # noqa: E402

FULL = ("echo-stereo", "eax_reverb-mono", "eax_reverb-stereo", "reverb-five_point_one", "chorus-stereo", "flanger-mono",
        "equalizer-stereo", "distortion-mono", "ring_modulator-stereo", "compressor-quad", "dedicated_dialog-seven_point_one",
        "eax_reverb-stereo-ragged", "preset-Default-forest-stereo", "reverb-mod-stereo", "chain-stereo", "chain2-mono96k",
        "chain-stereo-sends", "cfg2-schedule", "reverb-tap-churn", "type-swaps", "send-changes", "deferred-semantics")


def main():
    ref = H.ref_lib()
    assert ref is not None and ref.orc_kind() == b"reference", "the compiled reference is required"
    out = {}
    names = []
    for name, fmt, rate, effect_count, script, x in cases.all_cases(H.emu_lib(), quick=True):
        y = H.run_script_orc(ref, fmt, rate, effect_count, script, x)
        names.append(name)
        out["sha256/" + name] = np.frombuffer(hashlib.sha256(y.tobytes()).digest(), dtype=np.uint8)
        if name in FULL:
            out["full/" + name] = y
    missing = [n for n in FULL if n not in names]
    assert not missing, missing
    path = os.path.join(HERE, "golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(names)} hashes, {len(FULL)} full vectors, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
