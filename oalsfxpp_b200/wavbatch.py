"""Batch equivalent of the reference's WAV demo.

The reference's only executable, ``oalsfxpp_test <src.wav> <dst.wav>`` (reference:
src/oalsfxpp_test.cpp:747-901), reads one 8/16-bit PCM WAV file, runs it through ONE effect slot of one
``Api`` in a single ``mix`` call and writes a peak-normalised 16-bit file.  This module does the same for
MANY files in one engine call: every file is one stream of a batched engine, the PCM -> float ingest
(oalsfxpp_test.cpp:703-740) and the float -> s16 egress (oalsfxpp_test.cpp:602-651) run as CUDA kernels
through the C ABI (``oalsfx_pcm_to_float`` / ``oalsfx_float_to_s16``), so only 1- or 2-byte samples cross
PCIe.  Output files are byte-identical to the reference program's (tests/test_wav.py), including its
data-chunk-size quirk (the size field counts ``channels`` times too many bytes, oalsfxpp_test.cpp:556).

    python -m oalsfxpp_b200.wavbatch --effect 8 out_dir a.wav b.wav ...      # menu numbers as in the demo

There is no signal processing in this file: parsing, padding and file I/O only.
"""
import argparse
import os
import struct

import numpy as np

from .engine import Engine, LAYOUT_STREAM_MAJOR, SPACE_DEVICE, load_library
from .props import EffectType

# The demo's menu (oalsfxpp_test.cpp:791-804, 828-880)
MENU = {1: EffectType.eax_reverb, 2: EffectType.reverb, 3: EffectType.chorus, 4: EffectType.compressor,
        5: EffectType.dedicated_dialog, 6: EffectType.dedicated_low_frequency, 7: EffectType.distortion,
        8: EffectType.echo, 9: EffectType.equalizer, 10: EffectType.flanger, 11: EffectType.ring_modulator,
        12: EffectType.null}


class WavError(ValueError):
    pass


def read_wav(path):
    """(sampling_rate, channels, bit_depth, frames, raw sample bytes) of a PCM WAV file.  Same acceptance
    rules and messages as WavFile::read (oalsfxpp_test.cpp:300-500)."""
    data = open(path, "rb").read()
    if len(data) < 12 or data[0:4] != b"RIFF":
        raise WavError("Not a WAV stream.")
    riff_size = struct.unpack_from("<I", data, 4)[0]
    if riff_size + 8 < len(data):
        raise WavError("Truncated RIFF stream.")
    if data[8:12] != b"WAVE":
        raise WavError("Not a WAV stream.")
    pos = 12
    fmt = None
    raw = None
    while pos + 8 <= len(data) and not (fmt and raw is not None):
        four_cc = data[pos:pos + 4]
        size = struct.unpack_from("<I", data, pos + 4)[0]
        body = pos + 8
        if four_cc == b"fmt ":
            if size < 16:
                raise WavError("Invalid format chunk.")
            tag, channels, rate, _avg, block_align, bits = struct.unpack_from("<HHIIHH", data, body)
            if tag != 1:
                raise WavError("Expected a PCM codec.")
            if bits not in (8, 16):
                raise WavError("Unsupported bit depth.")
            if channels < 1 or block_align != channels * (bits // 8):
                raise WavError("Invalid format chunk.")
            fmt = (rate, channels, bits)
        elif four_cc == b"data":
            if size == 0:
                raise WavError("No data to read.")
            raw = data[body:body + size]
        pos = body + (size + 1) // 2 * 2
    if not fmt:
        raise WavError("No format chunk.")
    if raw is None:
        raise WavError("No data chunk.")
    rate, channels, bits = fmt
    frame_bytes = channels * (bits // 8)
    frames = len(raw) // frame_bytes
    return rate, channels, bits, frames, raw[:frames * frame_bytes]


def wav_bytes(rate, channels, frames, s16):
    """A 16-bit PCM WAV file image exactly as WavFile::write_pcm_s16_le lays it out (oalsfxpp_test.cpp:521-651)."""
    total_samples = frames * channels
    data_chunk_size = 2 * channels * total_samples  # the reference's quirk: channels counted twice
    riff_chunk_size = 4 + (8 + 16) + (8 + data_chunk_size)
    block_align = channels * 2
    header = b"RIFF" + struct.pack("<I", riff_chunk_size & 0xFFFFFFFF) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, channels, rate, block_align * rate, block_align, 16) + b"data" + struct.pack(
        "<I", data_chunk_size & 0xFFFFFFFF)
    return header + np.ascontiguousarray(s16, dtype="<i2").tobytes()


def _channel_format(channels):
    # Api::channel_count_to_channel_format (oalsfxpp.cpp:3858-3885): 1 2 4 6 7 8 channels
    table = {1: 1, 2: 2, 4: 3, 6: 4, 7: 6, 8: 7}
    if channels not in table:
        raise WavError("Unsupported channel count.")
    return table[channels]


def process(raw_inputs, rate, channels, bits, effect_type, lib=None, device=0):
    """raw_inputs: list of PCM byte strings (one per file, same rate / channels / bit depth, any lengths).
    Returns one int16 array [frames_i * channels] per input: what the reference demo writes for that file."""
    lib = lib if lib is not None else load_library()
    on_gpu = lib.oalsfx_build_info().decode().find("cuda") >= 0
    frame_bytes = channels * (bits // 8)
    frames = [len(r) // frame_bytes for r in raw_inputs]
    longest = max(frames)
    n = len(raw_inputs)
    # Shorter files are padded with PCM silence: the effects are causal, so the first frames_i output
    # frames of a stream do not depend on what follows them.
    silence = 128 if bits == 8 else 0
    pcm = np.full((n, longest * channels), silence, dtype=np.uint8 if bits == 8 else np.int16)
    for i, r in enumerate(raw_inputs):
        pcm[i, :frames[i] * channels] = np.frombuffer(r, dtype=np.uint8 if bits == 8 else "<i2")
    outs = []
    with Engine(n, _channel_format(channels), rate, 1, device=device, lib=lib) as eng:
        eng.set_effect(0, effect_type)
        count = n * longest * channels
        if on_gpu:
            import torch
            dev = torch.device("cuda", device)
            pcm_d = torch.from_numpy(pcm).to(dev)
            x = torch.empty(count, dtype=torch.float32, device=dev)
            y = torch.empty_like(x)
            eng.pcm_to_float(pcm_d, bits, x, count)
            eng.mix(x, y, frames=longest, layout=LAYOUT_STREAM_MAJOR, space=SPACE_DEVICE)
            s16 = torch.empty(count, dtype=torch.int16, device=dev)
            row = longest * channels
            if len(set(frames)) == 1:
                eng.float_to_s16(y, s16, n, row)
            else:  # every file is normalised over its own length
                for i in range(n):
                    eng.float_to_s16(y.data_ptr() + 4 * i * row, s16.data_ptr() + 2 * i * row, 1, frames[i] * channels)
            torch.cuda.synchronize(dev)
            host = s16.cpu().numpy().reshape(n, row)
        else:  # CPU test build of the engine: "device" memory is host memory
            x = np.empty(count, dtype=np.float32)
            y = np.empty_like(x)
            eng.pcm_to_float(pcm, bits, x, count)
            eng.mix(x, y, frames=longest, layout=LAYOUT_STREAM_MAJOR, space=SPACE_DEVICE)
            s16 = np.zeros(count, dtype=np.int16)
            row = longest * channels
            for i in range(n):
                eng.float_to_s16(y.ctypes.data + 4 * i * row, s16.ctypes.data + 2 * i * row, 1, frames[i] * channels)
            host = s16.reshape(n, row)
        for i in range(n):
            outs.append(host[i, :frames[i] * channels].copy())
    return outs


def main(argv=None):
    ap = argparse.ArgumentParser(description="Batch version of the reference's oalsfxpp_test WAV demo.")
    ap.add_argument("--effect", type=int, required=True, choices=sorted(MENU), help="effect number of the demo's menu")
    ap.add_argument("out_dir")
    ap.add_argument("inputs", nargs="+")
    args = ap.parse_args(argv)
    parsed = [read_wav(p) for p in args.inputs]
    rate, channels, bits = parsed[0][:3]
    for p, w in zip(args.inputs, parsed):
        if w[:3] != (rate, channels, bits):
            raise WavError(f"{p}: all files of a batch must share sampling rate, channel count and bit depth")
    outs = process([w[4] for w in parsed], rate, channels, bits, MENU[args.effect])
    os.makedirs(args.out_dir, exist_ok=True)
    for p, w, s16 in zip(args.inputs, parsed, outs):
        with open(os.path.join(args.out_dir, os.path.basename(p)), "wb") as f:
            f.write(wav_bytes(rate, channels, w[3], s16))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
