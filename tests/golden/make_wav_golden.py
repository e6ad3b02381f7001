#!/usr/bin/env python3
"""Generate tests/golden/wav_golden.npz from the reference's own WAV demo program.

oracle/_ref/oalsfxpp_test is the UNMODIFIED src/oalsfxpp_test.cpp + src/oalsfxpp.cpp of the reference,
compiled by oracle/Makefile.  For a handful of synthetic PCM inputs this script writes a WAV file, runs the
program on it (the effect is chosen on stdin, like a user would) and stores the input samples and the
complete output file image.  tests/test_wav.py replays the same inputs through the batch tool
(oalsfxpp_b200/wavbatch.py: PCM ingest kernel -> engine -> s16 egress kernel) and compares the bytes.

    python tests/golden/make_wav_golden.py        # needs /root/reference (build: make -C oracle ref)
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402

# name, channels, rate, bits, frames, menu number
CASES = [
    ("echo-stereo-s16", 2, 48000, 16, 6000, 8),
    ("null-mono-u8", 1, 44100, 8, 5000, 12),
    ("eax_reverb-stereo-s16", 2, 48000, 16, 6100, 1),
    ("distortion-mono-s16", 1, 22050, 16, 4099, 7),
    ("equalizer-quad-u8", 4, 48000, 8, 3000, 9),
    ("chorus-stereo-s16-loud", 2, 48000, 16, 5001, 3),
]


def pcm_input(index, channels, bits, frames, loud):
    x = H.noise(900 + index, channels, frames) * (1.999 if loud else 1.2)
    if bits == 16:
        return np.clip(np.round(x * 32767.0), -32768, 32767).astype("<i2")
    return np.clip(np.round(x * 127.0) + 128, 0, 255).astype(np.uint8)


def input_wav(channels, rate, bits, pcm):
    raw = pcm.tobytes()
    block_align = channels * (bits // 8)
    body = b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, channels, rate, rate * block_align, block_align, bits) + \
        b"data" + struct.pack("<I", len(raw)) + raw + (b"\0" if len(raw) % 2 else b"")
    return b"RIFF" + struct.pack("<I", len(body)) + body


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "oalsfxpp_test")
    assert os.path.exists(exe), "build the reference demo first: make -C oracle ref"
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for i, (name, channels, rate, bits, frames, menu) in enumerate(CASES):
            pcm = pcm_input(i, channels, bits, frames, name.endswith("loud"))
            src, dst = os.path.join(tmp, "in.wav"), os.path.join(tmp, "out.wav")
            open(src, "wb").write(input_wav(channels, rate, bits, pcm))
            subprocess.run([exe, src, dst], input=f"{menu}\n".encode(), stdout=subprocess.DEVNULL, check=True)
            out[f"{name}/meta"] = np.array([channels, rate, bits, frames, menu], dtype=np.int64)
            out[f"{name}/pcm"] = pcm
            out[f"{name}/wav"] = np.frombuffer(open(dst, "rb").read(), dtype=np.uint8)
    path = os.path.join(HERE, "wav_golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(CASES)} files, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
