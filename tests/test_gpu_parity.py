"""GPU parity tests proper: the CUDA library, driven through the C ABI (and the drop-in C++ class via
the shared shim), against the CPU checker on identical inputs.

Bar (BASELINE.json north_star): fp32 within 1e-5 max abs error; integer state bit-exact.  What is
actually asserted is stronger: every effect is BIT-exact except the ring modulator's sinusoid
carrier, whose `sinf` is CUDA's on the device and glibc's in the reference (<= 2 ulp apart, i.e.
~1e-7 on the output) -- those cases are held to TOL.
"""
import numpy as np
import pytest

import cases
import harness as H
import oalsfxpp_b200 as ox
from oalsfxpp_b200 import ChannelFormat as F, EffectType as T

pytestmark = pytest.mark.gpu

TOL = 1e-5  # max abs error per sample, the north_star tolerance


def _lib():
    return H.cuda_lib()


def _uses_device_sinf(script):
    """ring modulator with the sinusoid carrier (waveform 0, the default)."""
    active = {}
    hit = False
    for op in script:
        if op[0] == "type":
            active[op[1]] = (op[2], None)
        elif op[0] == "props":
            active[op[1]] = (active.get(op[1], (None, None))[0], op[2])
    for t, p in active.values():
        if t == T.ring_modulator and (p is None or p.ring_modulator_.waveform_ == 0):
            hit = True
    return hit


def _assert_match(expect, got, exact, what=""):
    assert expect.shape == got.shape
    if exact:
        # Bit for bit: the uint32 images must be equal, so +0 and -0 are NOT interchangeable.  One exception: where the
        # reference itself produces NaN (filters designed at / above Nyquist, e.g. at 8 kHz) both sides must be NaN, but
        # the payload is the hardware's default NaN -- 0xFFC00000 from SSE, 0x7FFFFFFF from the GPU -- and is not compared.
        e32 = np.ascontiguousarray(expect, dtype=np.float32).view(np.uint32)
        g32 = np.ascontiguousarray(got, dtype=np.float32).view(np.uint32)
        differ = (e32 != g32) & ~(np.isnan(expect) & np.isnan(got))
        if differ.any():
            bad = np.argwhere(differ)
            raise AssertionError((what, "first mismatch at", tuple(bad[0]), "count", len(bad), "max abs diff", H.max_abs_diff(expect, got)))
    else:
        finite = np.isfinite(expect)
        assert np.array_equal(finite, np.isfinite(got)), what
        assert H.max_abs_diff(expect[finite], got[finite]) <= TOL, (what, H.max_abs_diff(expect[finite], got[finite]))


CASES = list(cases.all_cases(None, quick=False)) if False else None


def _cases():
    global CASES
    if CASES is None:
        CASES = list(cases.all_cases(_lib(), quick=False))
    return CASES


def test_case_matrix_through_the_cpp_api(checker):
    """All ~200 cases (every effect x layout x rate x block partition x property extremes x schedules)
    through oalsfxpp::Api of the CUDA library."""
    shim = H.api_shim("cuda")
    failures = []
    for name, fmt, rate, effect_count, script, x in _cases():
        expect = H.run_script_orc(checker, fmt, rate, effect_count, script, x)
        got = H.run_script_orc(shim, fmt, rate, effect_count, script, x)
        try:
            _assert_match(expect, got, not _uses_device_sinf(script), name)
        except AssertionError as err:
            failures.append(str(err)[:200])
    assert not failures, failures


def test_batched_heterogeneous_streams(checker):
    lib = _lib()
    S, frames, block = 203, 4000, 1024
    blocks = H.blocks_of(frames, block)
    default = lambda t, **kw: ox.default_props(t, lib=lib, **kw)
    configs = [
        [(T.equalizer, None), (T.chorus, None), (T.echo, None), (T.eax_reverb, None)],
        [(T.flanger, default(T.flanger, waveform_=0)), (T.ring_modulator, default(T.ring_modulator, waveform_=1)),
         (T.distortion, None), (T.compressor, None)],
        [(T.eax_reverb, ox.reverb_preset("Default", "forest", lib=lib)), (T.null, None),
         (T.echo, default(T.echo, delay_=0.05)), (T.null, None)],
        [(T.reverb, ox.reverb_preset("Default", "padded_cell", lib=lib)), (T.dedicated_dialog, None),
         (T.equalizer, default(T.equalizer, mid1_gain_=3.0)), (T.chorus, default(T.chorus, waveform_=0, rate_=3.3))],
    ]
    which = [(s * 7 + s // 5) % 4 for s in range(S)]
    x = np.stack([H.noise(100 + s, 2, frames) for s in range(S)])
    with ox.Engine(S, F.stereo, 48000, 4, lib=lib) as eng:
        for s in range(S):
            for slot, (t, p) in enumerate(configs[which[s]]):
                eng.set_effect(slot, t, p, first_stream=s, n_streams=1)
        y = np.empty_like(x)
        pos = 0
        for n in blocks:
            y[:, pos:pos + n] = eng.mix(np.ascontiguousarray(x[:, pos:pos + n]))
            pos += n
    for s in range(0, S, 3):
        script = H.simple_script(configs[which[s]], blocks)
        expect = H.run_script_orc(checker, F.stereo, 48000, 4, script, x[s])
        _assert_match(expect, y[s], True, f"stream {s} config {which[s]}")


@pytest.mark.parametrize("effect", [T.eax_reverb, T.reverb, T.echo, T.chorus, T.flanger, T.equalizer, T.distortion,
                                    T.ring_modulator, T.compressor, T.dedicated_dialog])
def test_ten_seconds_per_effect(checker, effect):
    """north_star: within 1e-5 over 10 s of audio per effect (480 000 frames @48 kHz, 1024-frame blocks;
    32 streams with different noise, every 8th compared with the checker)."""
    lib = _lib()
    S, rate, block, total = 32, 48000, 1024, 480000
    blocks = H.blocks_of(total, block)
    x = np.stack([H.noise(s, 2, total) for s in range(S)])
    y = np.empty_like(x)
    with ox.Engine(S, F.stereo, rate, 1, lib=lib) as eng:
        eng.set_effect(0, effect)
        pos = 0
        for n in blocks:
            y[:, pos:pos + n] = eng.mix(np.ascontiguousarray(x[:, pos:pos + n]))
            pos += n
    script = H.simple_script([(effect, None)], blocks)
    for s in range(0, S, 8):
        expect = H.run_script_orc(checker, F.stereo, rate, 1, script, x[s])
        _assert_match(expect, y[s], effect != T.ring_modulator, f"{effect.name} stream {s}")


def _ten_seconds_at_full_size(checker, streams, fmt, rate, slots, exact, unique=32, check=(0, 5, 17, 31)):
    """10 s of audio at a BASELINE configuration's full stream count, device buffers, block by block.  Stream s is fed
    input s mod `unique` (the inputs are independent noise): every copy must equal the first one bit for bit after every
    block, and the `check` inputs are compared with the checker over the whole 10 s."""
    import torch
    lib = _lib()
    block, total = 1024, 10 * rate
    blocks = H.blocks_of(total, block)
    C = ox.channel_count(fmt)
    base = np.stack([H.noise(9000 + k, C, total) for k in range(unique)])
    base_dev = torch.from_numpy(base).cuda()
    kept = {k: [] for k in check}
    kernels = set()
    assert streams % unique == 0
    with ox.Engine(streams, fmt, rate, len(slots), lib=lib) as eng:
        for i, (t, p) in enumerate(slots):
            eng.set_effect(i, t, p)
        pos = 0
        for n in blocks:
            x = base_dev[:, pos:pos + n].repeat(streams // unique, 1, 1).contiguous()
            y = torch.empty_like(x)
            eng.mix(x, y, frames=n, stream=torch.cuda.current_stream().cuda_stream)
            per = y.view(streams // unique, unique, n, C)
            assert bool((per == per[0:1]).all()), f"streams with identical input diverged in the block at frame {pos}"
            for k in check:
                kept[k].append(per[0, k].cpu().numpy())
            kernels.add(eng.last_kernel)
            pos += n
        ints = [eng.debug_state(s, slot) for s in (0, streams - 1) for slot in range(len(slots))]
    script = H.simple_script(slots, blocks)
    for k in check:
        expect = H.run_script_orc(checker, fmt, rate, len(slots), script, base[k])
        _assert_match(expect, np.concatenate(kept[k], axis=0), exact, f"input {k}")
    return kernels, ints


def test_cfg1_ten_seconds_at_full_size(checker):
    """BASELINE cfg1: EAX reverb on 1024 mono 48 kHz streams, 10 s in 1024-frame blocks (the span kernel serves every
    block but the first), bit-exact; ring offset and cross-fade counter at the end."""
    kernels, ints = _ten_seconds_at_full_size(checker, 1024, F.mono, 48000, [(T.eax_reverb, None)], True)
    assert any(k.startswith("kSpan") for k in kernels), kernels
    for st in ints:
        assert st["offset"] == 480000 and st["fade_count"] == 128


def test_cfg3_ten_seconds_at_full_size(checker):
    """BASELINE cfg3: flanger (sinusoid LFO: its whole period of 355 556 samples is visited 2.7 times) + ring modulator +
    distortion + compressor on 16 384 mono streams at 96 kHz, 10 s; within 1e-5 (the ring modulator's carrier is the one
    device sinf), integer state exact."""
    lib = _lib()
    rate = 96000
    slots = [(T.flanger, ox.default_props(T.flanger, lib=lib, waveform_=0)), (T.ring_modulator, None), (T.distortion, None), (T.compressor, None)]
    kernels, ints = _ten_seconds_at_full_size(checker, 16384, F.mono, rate, slots, False, check=(0, 13, 31))
    total = 10 * rate
    step = int(np.float32(440.0) * np.float32(1 << 24) / np.float32(rate))
    for i, st in enumerate(ints):
        slot = i % 4
        if slot == 0:
            assert st["offset"] == total, st                         # flanger ring offset (oalsfxpp.cpp:5484)
        elif slot == 1:
            assert st["ring_mod_index"] == (total * step) & 0xFFFFFF, st   # oalsfxpp.cpp:5722-5738


@pytest.mark.parametrize("family", ["single", "duo", "quartet", "relay", "span"])
def test_integer_state_agrees_across_kernel_families(family, monkeypatch):
    """Ring offsets, the reverb's cross-fade counter and modulator index, the chorus offset: after every block, whichever
    kernel family ran it, they follow the reference's formulas (offset_ += n, oalsfxpp.cpp:7853 / 4212 / 4962; fade 128
    samples from an update that changes a tap, :6118-6138; modulator index modulo its range, :7443-7470)."""
    monkeypatch.setenv("OALSFX_KERNEL", family)
    lib = _lib()
    rate, S = 48000, 96
    mod_time = 0.25
    chain = [T.equalizer, T.chorus, T.echo, T.eax_reverb]
    blocks = [1024, 100, 2048, 77, 1024, 1024, 6]
    change_at = 4
    with ox.Engine(S, F.stereo, rate, 4, lib=lib) as eng:
        for i, t in enumerate(chain):
            eng.set_effect(i, t)
        eng.set_effect(3, T.eax_reverb, ox.default_props(T.eax_reverb, lib=lib, modulation_depth_=0.3, modulation_time_=mod_time))
        total, since = 0, 0
        for b, n in enumerate(blocks):
            if b == change_at:  # a tap change: the cross-fade counter restarts
                eng.set_effect(3, T.eax_reverb, ox.default_props(T.eax_reverb, lib=lib, modulation_depth_=0.3, modulation_time_=mod_time,
                                                               reflections_delay_=0.02))
                since = 0
            eng.mix(np.zeros((S, n, 2), np.float32))
            total += n
            since += n
            for s in (0, 31, 32, S - 1):
                st = eng.debug_state(s, 3)
                assert st["offset"] == total and st["fade_count"] == min(since, 128), (family, b, s, st)
                assert st["mod_index"] == total % int(mod_time * rate), (family, b, s, st)
                assert eng.debug_state(s, 1)["offset"] == total and eng.debug_state(s, 2)["offset"] == total


def test_cfg2_schedule_on_4096_streams(checker):
    """BASELINE cfg2: 4-slot chain, 4096 stereo streams, per-block parameter changes (reverb gain ramp,
    tap cross-fade every 16th block, EQ step); a sample of streams compared with the checker."""
    lib = _lib()
    S, block, nblocks = 4096, 1024, 36
    chain = [T.equalizer, T.chorus, T.echo, T.eax_reverb]
    x = np.stack([H.noise(s, 2, block * nblocks) for s in range(S)])
    y = np.empty_like(x)
    script = [("type", i, t) for i, t in enumerate(chain)] + [("apply",)]
    with ox.Engine(S, F.stereo, 48000, 4, lib=lib) as eng:
        for i, t in enumerate(chain):
            eng.set_effect(i, t)
        for b in range(nblocks):
            rv = ox.default_props(T.eax_reverb, lib=lib, gain_=0.20 + 0.10 * ((b % 4) / 4.0),
                                  reflections_delay_=(0.012 if (b // 16) % 2 else 0.007))
            eq = ox.default_props(T.equalizer, lib=lib, mid1_gain_=1.0 + 0.5 * ((b % 8) / 8.0))
            # Api::apply_changes only pushes a slot whose properties differ from the active ones
            if b == 0 or (b % 16 == 0) or True:
                eng.set_effect(3, T.eax_reverb, rv)
            eng.set_effect(0, T.equalizer, eq)
            script += [("props", 3, rv), ("props", 0, eq), ("apply",), ("mix", block)]
            y[:, b * block:(b + 1) * block] = eng.mix(np.ascontiguousarray(x[:, b * block:(b + 1) * block]))
    for s in (0, 1, 31, 32, 1000, 2047, 4095):
        expect = H.run_script_orc(checker, F.stereo, 48000, 4, script, x[s])
        _assert_match(expect, y[s], True, f"stream {s}")


@pytest.mark.parametrize("family", ["quartet", "duo", "single", "relay", "span"])
def test_every_kernel_family_on_the_chain(checker, family, monkeypatch):
    """The fused 4-slot stereo chain has five implementations (OALSFX_KERNEL, read when an engine is
    created): the two-stage duo kernel (default), the four-stage quartet pipeline, the relay pipeline, the plain
    thread-per-stream kernel and -- on blocks without a pending update -- the time-parallel span kernel (here: the
    device-side steady-state check and its exact fallback).  Each must equal the checker bit for bit on a
    schedule that exercises ramps (reverb gain), a tap cross-fade (reflections delay change in block 2),
    an equalizer step, a block size that is not a multiple of 4 and a partially filled last tile."""
    monkeypatch.setenv("OALSFX_KERNEL", family)
    lib = _lib()
    S = 160 + 7
    blocks = [1024, 1024, 640, 333, 1024]
    chain = [T.equalizer, T.chorus, T.echo, T.eax_reverb]
    total = sum(blocks)
    x = np.stack([H.noise(s, 2, total) for s in range(S)])
    y = np.empty_like(x)
    script = [("type", i, t) for i, t in enumerate(chain)] + [("apply",)]
    with ox.Engine(S, F.stereo, 48000, 4, lib=lib) as eng:
        for i, t in enumerate(chain):
            eng.set_effect(i, t)
        at = 0
        for b, n in enumerate(blocks):
            rv = ox.default_props(T.eax_reverb, lib=lib, gain_=0.20 + 0.10 * ((b % 4) / 4.0),
                                  reflections_delay_=(0.012 if b >= 2 else 0.007))
            eq = ox.default_props(T.equalizer, lib=lib, mid1_gain_=1.0 + 0.5 * ((b % 8) / 8.0))
            eng.set_effect(3, T.eax_reverb, rv)
            eng.set_effect(0, T.equalizer, eq)
            script += [("props", 3, rv), ("props", 0, eq), ("apply",), ("mix", n)]
            y[:, at:at + n] = eng.mix(np.ascontiguousarray(x[:, at:at + n]))
            at += n
    for s in (0, 31, 32, 100, 159, 160, S - 1):
        expect = H.run_script_orc(checker, F.stereo, 48000, 4, script, x[s])
        _assert_match(expect, y[s], True, f"{family} stream {s}")


@pytest.mark.parametrize("variant", ["standard-reverb-96k", "zero-delays", "density-extremes"])
@pytest.mark.parametrize("family", ["quartet", "duo"])
def test_pipelined_kernels_on_reverb_corner_cases(checker, family, variant, monkeypatch):
    """The warp-specialised kernels hand delay-line data from one warp to the next through HBM.  Corner cases
    of that hand-off: the standard (non-EAX) reverb at 96 kHz (all delays doubled, no high-pass stage);
    reflections_delay = late_reverb_delay = 0 (the early stage reads the sample the input stage wrote in the
    SAME frame, the late stage the one the early stage just fed: no prefetch possible, direct reads across
    warps); smallest / largest density and diffusion (shortest delay lines the presets allow)."""
    monkeypatch.setenv("OALSFX_KERNEL", family)
    lib = _lib()
    rate = 96000 if variant == "standard-reverb-96k" else 48000
    rtype = T.reverb if variant == "standard-reverb-96k" else T.eax_reverb
    if variant == "zero-delays":
        steps = [ox.default_props(rtype, lib=lib, reflections_delay_=0.0, late_reverb_delay_=0.0), None,
                 ox.default_props(rtype, lib=lib, reflections_delay_=0.0, late_reverb_delay_=0.0, decay_time_=0.3)]
    elif variant == "density-extremes":
        steps = [ox.default_props(rtype, lib=lib, density_=0.0, diffusion_=0.0), None,
                 ox.default_props(rtype, lib=lib, density_=1.0, diffusion_=1.0, decay_time_=0.1)]
    else:
        steps = [None, ox.default_props(rtype, lib=lib, gain_=0.5, late_reverb_gain_=2.0), None]
    S, blocks = 64, [1024, 700, 1024]
    chain = [T.equalizer, T.chorus, T.echo, rtype]
    total = sum(blocks)
    x = np.stack([H.noise(400 + s, 2, total) for s in range(S)])
    y = np.empty_like(x)
    script = [("type", i, t) for i, t in enumerate(chain)] + [("apply",)]
    with ox.Engine(S, F.stereo, rate, 4, lib=lib) as eng:
        for i, t in enumerate(chain):
            eng.set_effect(i, t)
        at = 0
        for n, p in zip(blocks, steps):
            if p is not None:
                eng.set_effect(3, rtype, p)
                script += [("props", 3, p), ("apply",)]
            script += [("mix", n)]
            y[:, at:at + n] = eng.mix(np.ascontiguousarray(x[:, at:at + n]))
            at += n
    for s in (0, 31, 32, 63):
        expect = H.run_script_orc(checker, F.stereo, rate, 4, script, x[s])
        _assert_match(expect, y[s], True, f"{family} {variant} stream {s}")


@pytest.mark.parametrize("family", ["quartet", "duo", "relay"])
def test_pipelined_kernels_under_load_are_race_free(checker, family, monkeypatch):
    """16 384 streams (512 tiles: several waves of CTAs, every SM busy) with zero reverb delays, i.e. the
    configuration in which a stage reads, in the same frame, what the previous stage's warp has just written.
    Every stream is fed one of four inputs: all copies must agree bit for bit (a lost hand-off ordering would
    show up as a tile that differs) and equal the checker."""
    import torch
    monkeypatch.setenv("OALSFX_KERNEL", family)
    lib = _lib()
    S, block, nblocks = 16384, 512, 3
    chain = [T.equalizer, T.chorus, T.echo, T.eax_reverb]
    props = ox.default_props(T.eax_reverb, lib=lib, reflections_delay_=0.0, late_reverb_delay_=0.0)
    base = [H.noise(s, 2, block * nblocks) for s in range(4)]
    dev = torch.device("cuda:0")
    outs = []
    with ox.Engine(S, F.stereo, 48000, 4, lib=lib) as eng:
        for i, t in enumerate(chain[:3]):
            eng.set_effect(i, t)
        eng.set_effect(3, T.eax_reverb, props)
        for b in range(nblocks):
            xb = torch.from_numpy(np.stack([v[b * block:(b + 1) * block] for v in base])).to(dev)
            x = xb.repeat(S // 4, 1, 1).contiguous()
            y = torch.empty_like(x)
            eng.mix(x, y, frames=block, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            outs.append(y)
    y = torch.cat(outs, dim=1)
    per = y.view(S // 4, 4, block * nblocks, 2)
    assert bool((per == per[0:1]).all()), "streams with identical input diverged"
    script = [("type", i, t) for i, t in enumerate(chain)] + [("props", 3, props), ("apply",)] + [("mix", block)] * nblocks
    for k in range(4):
        expect = H.run_script_orc(checker, F.stereo, 48000, 4, script, base[k])
        _assert_match(expect, per[0, k].cpu().numpy(), True, f"{family} input {k}")


@pytest.mark.parametrize("fmt", [F.mono, F.stereo])
@pytest.mark.parametrize("family", ["quartet", "single", "span"])
def test_every_kernel_family_on_the_single_reverb_slot(checker, family, fmt, monkeypatch):
    """cfg1's signature (one reverb slot, mono): pipeline (default), quad and plain kernels, with a preset
    change mid-stream (tap cross-fade + gain ramp), a modulated preset and odd block sizes."""
    monkeypatch.setenv("OALSFX_KERNEL", family)
    lib = _lib()
    S = 96
    blocks = [1024, 512, 1023, 1024, 6]
    total = sum(blocks)
    x = np.stack([H.noise(50 + s, ox.channel_count(fmt), total) for s in range(S)])
    y = np.empty_like(x)
    presets = [None, ox.reverb_preset("Default", "forest", lib=lib), None,
               ox.default_props(T.eax_reverb, lib=lib, modulation_depth_=0.8, modulation_time_=0.6), None]
    script = [("type", 0, T.eax_reverb), ("apply",)]
    with ox.Engine(S, fmt, 48000, 1, lib=lib) as eng:
        eng.set_effect(0, T.eax_reverb)
        at = 0
        for n, p in zip(blocks, presets):
            if p is not None:
                eng.set_effect(0, T.eax_reverb, p)
                script += [("props", 0, p), ("apply",)]
            script += [("mix", n)]
            y[:, at:at + n] = eng.mix(np.ascontiguousarray(x[:, at:at + n]))
            at += n
    for s in (0, 31, 32, 95):
        expect = H.run_script_orc(checker, fmt, 48000, 1, script, x[s])
        _assert_match(expect, y[s], True, f"{family} stream {s}")


@pytest.mark.parametrize("case", ["eax-mono", "eax-stereo", "standard-96k", "forest", "short-first-blocks", "dense"])
def test_span_kernel_on_steady_state_blocks(checker, case):
    """span.cuh: blocks without a pending update run the single reverb slot block-parallel in time (spans of
    up to 64 frames, recurrences in dedicated warps).  Steady-state blocks of odd sizes after a preset change,
    a ragged last tile, presets with short delay lines, and blocks so short that the tap cross-fade is still
    running when the next block starts (the device-side check sends those to the exact serial body)."""
    lib = _lib()
    fmt = F.mono if case in ("eax-mono", "short-first-blocks") else F.stereo
    rate = 96000 if case == "standard-96k" else 48000
    rtype = T.reverb if case == "standard-96k" else T.eax_reverb
    first = {"forest": ox.reverb_preset("Default", "forest", lib=lib),
             "dense": ox.default_props(rtype, lib=lib, density_=0.0, diffusion_=1.0, reflections_delay_=0.002, late_reverb_delay_=0.003)}.get(case)
    second = ox.default_props(rtype, lib=lib, gain_=0.5, decay_time_=3.0, reflections_delay_=0.02)
    blocks = [50, 30, 70, 1024, 333] if case == "short-first-blocks" else [1024, 1024, 777, 64, 2, 1500, 1024]
    change_at = 4 if case != "short-first-blocks" else 99
    S = 70
    C = ox.channel_count(fmt)
    total = sum(blocks)
    x = np.stack([H.noise(2000 + s, C, total) for s in range(S)])
    y = np.empty_like(x)
    script = [("type", 0, rtype)] + ([("props", 0, first)] if first is not None else []) + [("apply",)]
    with ox.Engine(S, fmt, rate, 1, lib=lib) as eng:
        eng.set_effect(0, rtype, first)
        at = 0
        for b, n in enumerate(blocks):
            if b == change_at:
                eng.set_effect(0, rtype, second)
                script += [("props", 0, second), ("apply",)]
            script += [("mix", n)]
            y[:, at:at + n] = eng.mix(np.ascontiguousarray(x[:, at:at + n]))
            at += n
    for s in (0, 31, 32, 63, 64, S - 1):
        expect = H.run_script_orc(checker, fmt, rate, 1, script, x[s])
        _assert_match(expect, y[s], True, f"span {case} stream {s}")


@pytest.mark.parametrize("case", ["default", "flanger-96k", "forest-short-echo", "short-first-blocks", "standard-reverb", "eax-mono"])
def test_span_bulk_kernel(checker, case, monkeypatch):
    """span_bulk_kernel: the same pipeline with the ring rows of a span moved by bulk copies (issued ahead of / behind the
    arithmetic, so with its own legality rule).  OALSFX_SPAN_BULK=2 selects it however few tiles the engine has."""
    monkeypatch.setenv("OALSFX_SPAN_BULK", "2")
    if case == "eax-mono":
        test_span_kernel_on_steady_state_blocks(checker, case)
    else:
        test_span_chain_on_steady_state_blocks(checker, case)


@pytest.mark.parametrize("case", ["default", "flanger-96k", "forest-short-echo", "short-first-blocks", "standard-reverb"])
def test_span_chain_on_steady_state_blocks(checker, case):
    """span.cuh on the 4-slot chain: blocks without a pending update run equalizer + chorus / flanger + echo + reverb
    block-parallel in time as a three-stage software pipeline (phase A of span i+1, the recurrences of span i and
    phase C of span i-1 overlap).  Steady-state blocks of odd sizes, a ragged last tile, parameter changes in between
    (the update block and cross-fading blocks take the exact kernels), short echo / flanger delays (shorter spans or no
    span at all), the standard reverb (no high-pass shelf)."""
    lib = _lib()
    rate = 96000 if case == "flanger-96k" else 48000
    rtype = T.reverb if case == "standard-reverb" else T.eax_reverb
    mtype = T.flanger if case == "flanger-96k" else T.chorus
    chain = [T.equalizer, mtype, T.echo, rtype]
    first = {3: ox.reverb_preset("Default", "forest", lib=lib), 2: ox.default_props(T.echo, lib=lib, delay_=0.002, lr_delay_=0.003)} if case == "forest-short-echo" else {}
    if case == "flanger-96k":  # the default flanger sweeps its delay down to zero (depth 1): no span is legal with that
        first = {1: ox.default_props(T.flanger, lib=lib, delay_=0.004, depth_=0.25, rate_=1.3, waveform_=0)}
    second = {3: ox.default_props(rtype, lib=lib, gain_=0.5, decay_time_=3.0, reflections_delay_=0.02),
              0: ox.default_props(T.equalizer, lib=lib, mid1_gain_=1.7, low_gain_=0.5),
              2: ox.default_props(T.echo, lib=lib, delay_=0.05, feedback_=0.7)}
    blocks = [50, 30, 70, 1024, 333] if case == "short-first-blocks" else [1024, 1024, 777, 64, 2, 1500, 1024, 31]
    change_at = 4 if case != "short-first-blocks" else 99
    S = 70
    total = sum(blocks)
    x = np.stack([H.noise(3000 + s, 2, total) for s in range(S)])
    y = np.empty_like(x)
    script = [("type", i, t) for i, t in enumerate(chain)] + [("props", i, p) for i, p in first.items()] + [("apply",)]
    with ox.Engine(S, F.stereo, rate, 4, lib=lib) as eng:
        for i, t in enumerate(chain):
            eng.set_effect(i, t, first.get(i))
        at = 0
        for b, n in enumerate(blocks):
            if b == change_at:
                for i, p in second.items():
                    eng.set_effect(i, chain[i], p)
                    script += [("props", i, p)]
                script += [("apply",)]
            script += [("mix", n)]
            y[:, at:at + n] = eng.mix(np.ascontiguousarray(x[:, at:at + n]))
            at += n
    for s in (0, 31, 32, 63, 64, S - 1):
        expect = H.run_script_orc(checker, F.stereo, rate, 4, script, x[s])
        _assert_match(expect, y[s], True, f"span chain {case} stream {s}")


@pytest.mark.parametrize("tiles", [8, 50, 100])
def test_span_chain_every_tile_share(checker, tiles):
    """The chain's span kernel with a tile shared by 4 / 2 / 1 CTAs (8 / 16 / 32 streams per CTA)."""
    import torch
    lib = _lib()
    S = tiles * 32
    blocks = [256, 1024, 333, 1024]
    total = sum(blocks)
    chain = [T.equalizer, T.chorus, T.echo, T.eax_reverb]
    base = [H.noise(7100 + s, 2, total) for s in range(4)]
    dev = torch.device("cuda:0")
    outs = []
    with ox.Engine(S, F.stereo, 48000, 4, lib=lib) as eng:
        for i, t in enumerate(chain):
            eng.set_effect(i, t)
        at = 0
        for n in blocks:
            xb = torch.from_numpy(np.stack([v[at:at + n] for v in base])).to(dev)
            x = xb.repeat(S // 4, 1, 1).contiguous()
            y = torch.empty_like(x)
            eng.mix(x, y, frames=n, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            outs.append(y)
            at += n
    y = torch.cat(outs, dim=1)
    per = y.view(S // 4, 4, total, 2)
    assert bool((per == per[0:1]).all()), "streams with identical input diverged"
    script = [("type", i, t) for i, t in enumerate(chain)] + [("apply",)] + [("mix", n) for n in blocks]
    for k in range(4):
        expect = H.run_script_orc(checker, F.stereo, 48000, 4, script, base[k])
        _assert_match(expect, per[0, k].cpu().numpy(), True, f"span chain share, {tiles} tiles, input {k}")


@pytest.mark.parametrize("tiles", [8, 50, 100])
def test_span_kernel_every_tile_share(checker, tiles):
    """The span kernel shares a tile among 4 / 2 / 1 CTAs depending on the tile count (8 / 16 / 32 streams per CTA):
    8, 50 and 100 tiles pick the three variants.  Every stream is fed one of four inputs; all copies must agree bit
    for bit and equal the checker, over steady-state blocks of odd sizes (the first block carries the update)."""
    import torch
    lib = _lib()
    S = tiles * 32
    blocks = [256, 1024, 333, 1024]
    total = sum(blocks)
    base = [H.noise(7000 + s, 1, total) for s in range(4)]
    dev = torch.device("cuda:0")
    outs = []
    with ox.Engine(S, F.mono, 48000, 1, lib=lib) as eng:
        eng.set_effect(0, T.eax_reverb)
        at = 0
        for n in blocks:
            xb = torch.from_numpy(np.stack([v[at:at + n] for v in base])).to(dev)
            x = xb.repeat(S // 4, 1, 1).contiguous()
            y = torch.empty_like(x)
            eng.mix(x, y, frames=n, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            outs.append(y)
            at += n
    y = torch.cat(outs, dim=1)
    per = y.view(S // 4, 4, total, 1)
    assert bool((per == per[0:1]).all()), "streams with identical input diverged"
    script = [("type", 0, T.eax_reverb), ("apply",)] + [("mix", n) for n in blocks]
    for k in range(4):
        expect = H.run_script_orc(checker, F.mono, 48000, 1, script, base[k])
        _assert_match(expect, per[0, k].cpu().numpy(), True, f"span share, {tiles} tiles, input {k}")


def test_placement_keeps_arbitrary_presets_on_the_class_per_tile_launch(checker):
    """f1: every voice its own preset in no particular order.  Left in the caller's order the tiles hold mixed classes
    (table mode); ordered by oalsfx_plan_placement every tile is class-pure, one set_effect call per class covers its
    contiguous range, and a block is ONE class-per-tile launch.  Bit-exact per caller stream; the padding streams are
    silent."""
    lib = _lib()
    names = ox.reverb_preset_names(lib=lib)
    n_user, n_presets, blocks = 700, 23, [1024, 333]
    labels = np.array([(s * 7 + s // 50) % n_presets for s in range(n_user)], dtype=np.int32)
    index, total, classes = ox.plan_placement(labels, lib=lib)
    assert len(classes) == n_presets and total % 32 == 0 and total >= n_user
    assert len(set(index.tolist())) == n_user and index.max() < total
    for label, first, span in classes:
        members = np.nonzero(labels == label)[0]
        assert span == (len(members) + 31) // 32 * 32 and first % 32 == 0
        assert np.array_equal(index[members], first + np.arange(len(members)))   # contiguous, caller order kept
    frames = sum(blocks)
    x_user = np.stack([H.noise(7000 + s, 2, frames) for s in range(n_user)])
    x = np.zeros((total, frames, 2), dtype=np.float32)
    x[index] = x_user
    y = np.empty_like(x)
    chain = [T.equalizer, T.chorus, T.echo]
    props = {label: ox.reverb_preset(*names[(label * 5) % len(names)], lib=lib) for label, _, _ in classes}
    with ox.Engine(total, F.stereo, 48000, 4, lib=lib) as eng:
        for i, t in enumerate(chain):
            eng.set_effect(i, t)
        for label, first, span in classes:                       # (padding streams included: they share their tile's class)
            eng.set_effect(3, T.eax_reverb, props[label], first_stream=first, n_streams=span)
        at = 0
        for n in blocks:
            before = eng.launch_count
            y[:, at:at + n] = eng.mix(np.ascontiguousarray(x[:, at:at + n]))
            assert eng.launch_count - before == 1 and eng.last_kernel == "kMultiChainStereo", (eng.launch_count - before, eng.last_kernel)
            at += n
    for s in (0, 1, 49, 50, 333, n_user - 1):
        script = H.simple_script([(t, None) for t in chain] + [(T.eax_reverb, props[int(labels[s])])], blocks)
        expect = H.run_script_orc(checker, F.stereo, 48000, 4, script, x_user[s])
        _assert_match(expect, y[index[s]], True, f"caller stream {s}")
    used = np.zeros(total, dtype=bool)
    used[index] = True
    assert not y[~used].any()


def test_wide_relay_with_send_filters_over_four_seconds(checker):
    """The 4-slot chain on 5.1 with active send shelf filters (kRelaySfWideHeavy) for 4 s in 2048-frame blocks: every delay
    line wraps several times, the send filter histories and the six-channel pan gains settle; bit-exact per stream."""
    lib = _lib()
    fmt, S, block, nblocks = F.five_point_one, 40, 2048, 94
    C = ox.channel_count(fmt)
    slots = [T.equalizer, T.chorus, T.echo, T.eax_reverb]
    sends = {-1: (0.8, 0.5, 1.0), 0: (0.7, 1.0, 0.4), 2: (1.0, 0.25, 0.5), 3: (0.9, 0.6, 0.8)}
    total = block * nblocks
    x = np.stack([H.noise(9000 + s, C, total) for s in range(S)])
    y = np.empty_like(x)
    with ox.Engine(S, fmt, 48000, 4, lib=lib) as eng:
        for i, t in enumerate(slots):
            eng.set_effect(i, t)
        eng.set_sends(direct=sends[-1], aux=[sends.get(i, (1.0, 1.0, 1.0)) for i in range(4)])
        for b in range(nblocks):
            y[:, b * block:(b + 1) * block] = eng.mix(np.ascontiguousarray(x[:, b * block:(b + 1) * block]))
        assert eng.launch_count == nblocks and eng.last_kernel == "kRelaySfWideHeavy", (eng.launch_count, eng.last_kernel)
    script = [("type", i, t) for i, t in enumerate(slots)] + [("send", i, sends.get(i, (1.0, 1.0, 1.0))) for i in range(-1, 4)] + [("apply",)]
    script += [("mix", block)] * nblocks
    for s in (0, 31, 39):
        expect = H.run_script_orc(checker, fmt, 48000, 4, script, x[s])
        _assert_match(expect, y[s], True, f"stream {s}")


def test_waveshaper_divisions_are_ieee_divisions():
    """The distortion stage runs its twelve divisions per sample (oalsfxpp.cpp:4720-4722) as a batched correctly-rounded
    sequence behind one guard instead of through the `/` operator (fx.cuh, FxDistortion::shape).  Bit for bit against
    the same formula in numpy float32 (IEEE division): samples over 80 decades of amplitude, right at the guard's edges,
    every kind of bit pattern (zeros of both signs, denormals, infinities, NaNs), groups of four that mix fast-path and
    slow-path samples, and edge coefficients inside and outside the guard."""
    rng = np.random.default_rng(20)
    n = 1 << 21
    inputs = []
    inputs.append((rng.standard_normal(n) * np.exp(rng.uniform(-90.0, 30.0, n))).astype(np.float32))
    e = rng.integers(60, 142, n).astype(np.uint32)   # exponents around both ends of the guard (67 .. 134)
    inputs.append(((rng.integers(0, 2, n).astype(np.uint32) << 31) | (e << 23) | rng.integers(0, 1 << 23, n).astype(np.uint32)).view(np.float32))
    inputs.append(rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32))
    special = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-38, 2.0 ** -61, 2.0 ** -60, -(2.0 ** -60), 1.0, -1.0, 255.99, 256.0, -256.0, 1e38,
                        np.inf, -np.inf, np.nan], dtype=np.float32)
    inputs.append(np.concatenate([np.repeat(special, 4), np.tile(special, 4), rng.permutation(np.repeat(special, 64))]))
    mixed = inputs[0].copy()
    mixed[::7] = inputs[2][: len(mixed[::7])]
    inputs.append(mixed)
    inputs.append(np.zeros(4096, dtype=np.float32))
    inputs.append(-np.zeros(4096, dtype=np.float32))
    on_device = _lib() is H.cuda_lib()   # tests/test_emu_parity.py runs this scenario on the CPU backend (host buffers)
    with ox.Engine(32, F.mono, 48000, 1, lib=_lib()) as eng:
        for fc in (0.0, 0.37281, 2.0, 198.0, 200.0, 200.5, 1e6, -0.5, float("inf")):
            f = np.float32(fc)
            for i, s in enumerate(inputs):
                with np.errstate(all="ignore"):
                    one = np.float32(1.0)
                    x = ((one + f) * s / (one + (f * np.abs(s)))).astype(np.float32)
                    x = ((one + f) * x / (one + (f * np.abs(x))) * np.float32(-1.0)).astype(np.float32)
                    x = ((one + f) * x / (one + (f * np.abs(x)))).astype(np.float32)
                if on_device:
                    import torch
                    ts = torch.from_numpy(s.copy()).to("cuda:0")
                    out = torch.empty_like(ts)
                    eng.debug_waveshaper(ts, fc, out, ts.numel(), stream=torch.cuda.current_stream().cuda_stream)
                    torch.cuda.synchronize()
                    got = out.cpu().numpy()
                else:
                    got = np.empty_like(s)
                    eng.debug_waveshaper(np.ascontiguousarray(s), fc, got, s.size)
                _assert_match(x, got, True, f"edge coefficient {fc}, input set {i}")


@pytest.mark.parametrize("case", ["chain-stereo", "echo-mono", "reverb-eq-stereo", "chain-5.1", "flanger-7.1"])
def test_send_shelf_filters_run_in_one_launch(checker, case):
    """Active send shelf filters (apply_filters, oalsfxpp.cpp:3101-3143) used to drop a group to one exact pass per slot;
    the relay pipeline now carries them (relay_sf_kernel: every stage filters the input with its own send's filters, the
    histories are state): one launch per block, bit-exact, also when the send settings change between blocks."""
    lib = _lib()
    fmt = {"echo-mono": F.mono, "chain-5.1": F.five_point_one, "flanger-7.1": F.seven_point_one}.get(case, F.stereo)
    slots = {"chain-stereo": [T.equalizer, T.chorus, T.echo, T.eax_reverb], "echo-mono": [T.echo],
             "reverb-eq-stereo": [T.eax_reverb, T.null, T.equalizer], "chain-5.1": [T.equalizer, T.chorus, T.echo, T.eax_reverb],
             "flanger-7.1": [T.flanger, T.compressor]}[case]
    sends_a = {-1: (0.8, 0.5, 1.0), 0: (0.7, 1.0, 0.4), 2: (1.0, 0.25, 0.5)}
    sends_b = {-1: (1.0, 1.0, 0.3), 0: (0.9, 0.3, 0.6)}
    S, blocks = 70, [1024, 333, 1024, 2, 640]
    C = ox.channel_count(fmt)
    total = sum(blocks)
    x = np.stack([H.noise(5000 + s, C, total) for s in range(S)])
    y = np.empty_like(x)
    script = [("type", i, t) for i, t in enumerate(slots)]

    def apply_sends(eng, sends):
        triples = [sends.get(i, (1.0, 1.0, 1.0)) for i in range(len(slots))]
        eng.set_sends(direct=sends.get(-1, (1.0, 1.0, 1.0)), aux=triples)
        return [("send", i, sends.get(i, (1.0, 1.0, 1.0))) for i in range(-1, len(slots))] + [("apply",)]

    with ox.Engine(S, fmt, 48000, len(slots), lib=lib) as eng:
        for i, t in enumerate(slots):
            eng.set_effect(i, t)
        script += apply_sends(eng, sends_a)
        at = 0
        for b, n in enumerate(blocks):
            if b == 3:
                script += apply_sends(eng, sends_b)
            before = eng.launch_count
            y[:, at:at + n] = eng.mix(np.ascontiguousarray(x[:, at:at + n]))
            assert eng.launch_count - before == 1, (case, b, eng.last_kernel)
            assert eng.last_kernel.startswith("kRelaySf"), eng.last_kernel
            script += [("mix", n)]
            at += n
    for s in (0, 31, 32, 64, S - 1):
        expect = H.run_script_orc(checker, fmt, 48000, len(slots), script, x[s])
        _assert_match(expect, y[s], True, f"{case} stream {s}")


@pytest.mark.parametrize("settings", ["moderate", "extreme"])
@pytest.mark.parametrize("fmt", [F.mono, F.stereo])
def test_equalizer_as_a_linear_recurrence_scan(checker, fmt, settings, monkeypatch):
    """scan.cuh (OALSFX_SCAN=1): a lone equalizer slot with its 32 lanes spread over time -- zero-state pass per chunk,
    fp64 Kogge-Stone stitch of the chunk states, second pass from the true state.  It re-associates the filter sums, and
    the reference's direct-form filters with poles close to the unit circle (a 200 Hz shelf at 48 kHz: radius 0.99; the
    "extreme" case: 18 dB shelves and a 0.01-octave peak, radius 1 - 1e-4) amplify every rounding by 1e2 .. 1e4: the
    reference's own output moves by 1e-5 .. 1e-3 when its input moves by one ulp.  No evaluation order other than the
    reference's can land within 1e-5 of it there, so the bar is: 1e-5 (the north_star tolerance), or 8x the reference's
    own response to a one-ulp nudge of the same input, whichever is larger.  Also: parameter changes between blocks,
    block sizes that do not divide by 32, one too short for the scan (the exact kernel takes it)."""
    monkeypatch.setenv("OALSFX_SCAN", "1")
    lib = _lib()
    S, blocks = 40, [1024, 2048, 777, 64, 33, 1500]
    C = ox.channel_count(fmt)
    total = sum(blocks)
    x = np.stack([H.noise(6000 + s, C, total) for s in range(S)])
    y = np.empty_like(x)
    if settings == "extreme":
        steps = [ox.default_props(T.equalizer, lib=lib, low_gain_=7.943, low_cutoff_=50.0, mid1_gain_=0.126, mid1_width_=0.01, high_gain_=7.943),
                 None, ox.default_props(T.equalizer, lib=lib, mid2_gain_=7.943, mid2_center_=1000.0, mid2_width_=1.0, high_cutoff_=4000.0), None, None, None]
    else:
        steps = [None, None, ox.default_props(T.equalizer, lib=lib, low_gain_=2.0, mid1_gain_=0.5, mid2_gain_=1.5, mid2_width_=0.5, high_gain_=0.7), None, None, None]
    script = [("type", 1, T.equalizer), ("apply",)]
    kernels = []
    with ox.Engine(S, fmt, 48000, 3, lib=lib) as eng:
        eng.set_effect(1, T.equalizer)
        at = 0
        for n, p in zip(blocks, steps):
            if p is not None:
                eng.set_effect(1, T.equalizer, p)
                script += [("props", 1, p), ("apply",)]
            script += [("mix", n)]
            y[:, at:at + n] = eng.mix(np.ascontiguousarray(x[:, at:at + n]))
            kernels.append(eng.last_kernel)
            at += n
    assert kernels[0].startswith("kScanEqualizer") and kernels[4] == "kGenEqualizer", kernels
    worst = bound = 0.0
    for s in (0, 1, 31, 32, S - 1):
        expect = H.run_script_orc(checker, fmt, 48000, 3, script, x[s])
        nudged = H.run_script_orc(checker, fmt, 48000, 3, script, np.nextafter(x[s], np.float32(np.inf)))
        limit = max(TOL, 8.0 * H.max_abs_diff(expect, nudged))   # the reference's own response to a one-ulp nudge
        diff = H.max_abs_diff(expect, y[s])
        worst, bound = max(worst, diff), max(bound, limit)
        assert diff <= limit, (s, diff, limit)
    print(f"scan equalizer [{settings}]: max |diff| = {worst:.3e} (limit {bound:.3e})")


RELAY_SIGNATURES = [
    # (format, rate, slots) -- none of these has a fused kernel of its own: the relay pipeline runs them in one launch
    (F.stereo, 48000, [T.echo, T.eax_reverb]),
    (F.stereo, 48000, [T.eax_reverb, T.chorus, T.reverb, T.compressor]),       # reverb in stage 0, two reverbs
    (F.stereo, 48000, [T.distortion, T.null, T.flanger, T.equalizer]),          # a null slot in the middle
    (F.stereo, 44100, [T.ring_modulator, T.dedicated_dialog, T.echo, T.chorus]),
    (F.mono, 48000, [T.chorus, T.echo]),
    (F.mono, 96000, [T.equalizer, T.distortion, T.compressor, T.dedicated_low_frequency]),
    (F.mono, 48000, [T.null, T.reverb, T.echo, T.flanger]),
    (F.stereo, 48000, [T.compressor, T.equalizer, T.null, T.null]),
    # more than two device channels: the relay kernels with a run-time channel count (index 8 ..)
    (F.five_point_one, 48000, [T.equalizer, T.chorus, T.echo, T.eax_reverb]),    # BASELINE's chain on 5.1
    (F.quad, 44100, [T.echo, T.eax_reverb]),
    (F.seven_point_one, 48000, [T.flanger, T.null, T.distortion, T.compressor]),
    (F.six_point_one, 48000, [T.reverb, T.dedicated_dialog, T.equalizer, T.ring_modulator]),
    (F.five_point_one_rear, 96000, [T.chorus, T.echo]),
]


@pytest.mark.parametrize("sig", range(len(RELAY_SIGNATURES)))
def test_relay_pipeline_runs_any_signature_in_one_launch(checker, sig):
    """relay.cuh: one warp per non-null slot, effect per stage chosen at run time.  Signatures without a fused
    kernel, parameter changes (ramps / cross-fades / filter steps) in block 2, block sizes that are not a
    multiple of the hand-off, a ragged last tile.  One launch per block, bit-exact against the checker
    (ring modulator: device sinf, 1e-5)."""
    fmt, rate, slots = RELAY_SIGNATURES[sig]
    lib = _lib()
    C = ox.channel_count(fmt)
    S, blocks = 70, [1024, 333, 1024, 2, 641]
    total = sum(blocks)
    x = np.stack([H.noise(900 + 10 * sig + s, C, total) for s in range(S)])
    y = np.empty_like(x)

    def changed(t):
        if t in (T.eax_reverb, T.reverb):
            return ox.default_props(t, lib=lib, gain_=0.2, reflections_delay_=0.012, decay_time_=2.5)
        if t == T.equalizer:
            return ox.default_props(t, lib=lib, mid1_gain_=1.7)
        if t == T.echo:
            return ox.default_props(t, lib=lib, delay_=0.05, feedback_=0.7)
        if t in (T.chorus, T.flanger):
            return ox.default_props(t, lib=lib, depth_=0.5, feedback_=-0.4)
        if t == T.distortion:
            return ox.default_props(t, lib=lib, edge_=0.7)
        if t == T.ring_modulator:
            return ox.default_props(t, lib=lib, frequency_=1000.0)
        return None

    script = [("type", i, t) for i, t in enumerate(slots)] + [("apply",)]
    with ox.Engine(S, fmt, rate, len(slots), lib=lib) as eng:
        for i, t in enumerate(slots):
            eng.set_effect(i, t)
        at = 0
        for b, n in enumerate(blocks):
            if b == 2:
                for i, t in enumerate(slots):
                    p = changed(t)
                    if p is not None:
                        eng.set_effect(i, t, p)
                        script += [("props", i, p)]
                script += [("apply",)]
            script += [("mix", n)]
            y[:, at:at + n] = eng.mix(np.ascontiguousarray(x[:, at:at + n]))
            at += n
        assert eng.launch_count == len(blocks), eng.launch_count
        assert eng.last_kernel.startswith("kRelayWide" if C > 2 else "kRelay"), eng.last_kernel
    exact = T.ring_modulator not in slots
    for s in (0, 31, 32, 63, 64, S - 1):
        expect = H.run_script_orc(checker, fmt, rate, len(slots), script, x[s])
        _assert_match(expect, y[s], exact, f"relay signature {sig} stream {s}")


def test_relay_pipeline_on_the_cfg3_chain(checker, monkeypatch):
    """cfg3's signature (flanger + ring modulator + distortion + compressor, mono, 96 kHz) has a fused
    thread-per-stream kernel; OALSFX_KERNEL=relay runs it as a four-warp pipeline instead."""
    monkeypatch.setenv("OALSFX_KERNEL", "relay")
    lib = _lib()
    slots = [T.flanger, T.ring_modulator, T.distortion, T.compressor]
    S, blocks = 96, [1024, 1024, 500]
    total = sum(blocks)
    x = np.stack([H.noise(1300 + s, 1, total) for s in range(S)])
    y = np.empty_like(x)
    script = H.simple_script([(t, None) for t in slots], blocks)
    with ox.Engine(S, F.mono, 96000, 4, lib=lib) as eng:
        for i, t in enumerate(slots):
            eng.set_effect(i, t)
        at = 0
        for n in blocks:
            y[:, at:at + n] = eng.mix(np.ascontiguousarray(x[:, at:at + n]))
            at += n
        assert eng.launch_count == len(blocks)
    for s in (0, 31, 32, 95):
        expect = H.run_script_orc(checker, F.mono, 96000, 4, script, x[s])
        _assert_match(expect, y[s], False, f"relay cfg3 stream {s}")


def test_full_size_chain_device_buffers(checker):
    """BASELINE cfg4 at full size on one GPU: 65 536 stereo streams, 4-slot chain, device-resident
    buffers.  Size-independent properties: (1) streams fed identical input produce identical output
    (checksum of checksums over all tiles), (2) sampled streams equal the checker bit for bit,
    (3) the integer ring offset of every sampled stream equals the frames mixed."""
    import torch
    lib = _lib()
    S, block, nblocks = 65536, 1024, 3
    chain = [T.equalizer, T.chorus, T.echo, T.eax_reverb]
    base = [H.noise(s, 2, block * nblocks) for s in range(4)]
    dev = torch.device("cuda:0")
    with ox.Engine(S, F.stereo, 48000, 4, lib=lib) as eng:
        for i, t in enumerate(chain):
            eng.set_effect(i, t)
        outs = []
        for b in range(nblocks):
            # stream s gets input base[s % 4]
            xb = torch.from_numpy(np.stack([v[b * block:(b + 1) * block] for v in base])).to(dev)
            x = xb.repeat(S // 4, 1, 1).contiguous()
            y = torch.empty_like(x)
            eng.mix(x, y, frames=block, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            outs.append(y)
        y = torch.cat(outs, dim=1)  # [S][F][C]
        assert eng.debug_state(S - 1, 3)["offset"] == block * nblocks
        assert eng.debug_state(12345, 2)["offset"] == block * nblocks
    per = y.view(S // 4, 4, block * nblocks, 2)
    assert bool((per == per[0:1]).all()), "streams with identical input diverged"
    script = H.simple_script([(t, None) for t in chain], [block] * nblocks)
    for k in range(4):
        expect = H.run_script_orc(checker, F.stereo, 48000, 4, script, base[k])
        _assert_match(expect, per[0, k].cpu().numpy(), True, f"input {k}")


def test_tiled_layout_and_bus_reduce():
    import torch
    lib = _lib()
    S, n = 96, 512
    x = np.stack([H.noise(s, 2, n) for s in range(S)])
    with ox.Engine(S, F.stereo, 48000, 2, lib=lib) as eng:
        eng.set_effect(0, T.echo)
        eng.set_effect(1, T.eax_reverb)
        y = eng.mix(x)
    with ox.Engine(S, F.stereo, 48000, 2, lib=lib) as eng:
        eng.set_effect(0, T.echo)
        eng.set_effect(1, T.eax_reverb)
        tiled = np.ascontiguousarray(x.reshape(-1, 32, n, 2).transpose(0, 2, 3, 1))
        xt = torch.from_numpy(tiled).cuda()
        yt = torch.empty_like(xt)
        eng.mix(xt, yt, frames=n, layout=ox.LAYOUT_TILED)
        bus = torch.zeros(n, 2, device="cuda")
        eng.reduce_bus(n, yt, bus, layout=ox.LAYOUT_TILED)
        torch.cuda.synchronize()
        y2 = yt.cpu().numpy().transpose(0, 3, 1, 2).reshape(S, n, 2)
    assert np.array_equal(y, y2)
    want = y.astype(np.float64).sum(axis=0)
    assert np.max(np.abs(bus.cpu().numpy() - want)) <= 1e-5 * np.sqrt(S)


def test_many_parameter_sets_table_mode(checker):
    """1000 streams, every one with its own parameters (table mode: coefficient blocks per stream from HBM),
    two slot signatures, send filters, a parameter change mid-stream.  A handful of launches per block, not
    one per parameter set; sampled streams equal the checker."""
    launches, nblocks = H.many_parameter_sets(_lib(), checker, 1000, 1500, 512, (0, 1, 4, 7, 31, 32, 333, 504, 998, 999),
                                              exact_all=False)
    assert launches <= nblocks * 2 * 4


def test_class_per_tile_launch_many_presets(checker):
    """Every run of 32 streams its own reverb preset / equalizer / echo settings (40 parameter classes, a ragged last
    tile): ONE duo_multi launch per block with per-tile coefficient blocks from HBM.  A preset change on some tiles
    in block 2 (pending bits per class), a 1-frame block (the fused kernel needs >= 2 frames: per-class fallback)."""
    lib = _lib()
    names = ox.reverb_preset_names(lib=lib)
    tiles = 40
    S = tiles * 32 - 5
    blocks = [1024, 500, 1024, 1, 777]
    total = sum(blocks)
    x = np.stack([H.noise(3000 + s, 2, total) for s in range(S)])
    y = np.empty_like(x)

    def config(t, phase):
        g, n = names[(7 * t + 3 + 11 * phase * (t % 3 == 0)) % len(names)]
        return [(T.equalizer, ox.default_props(T.equalizer, lib=lib, mid1_gain_=0.6 + 0.03 * t)),
                (T.chorus, ox.default_props(T.chorus, lib=lib, depth_=0.05 + 0.002 * t)),
                (T.echo, ox.default_props(T.echo, lib=lib, delay_=0.03 + 0.002 * t, feedback_=0.2 + 0.01 * (t % 30))),
                (T.eax_reverb, ox.reverb_preset(g, n, lib=lib))]

    with ox.Engine(S, F.stereo, 48000, 4, lib=lib) as eng:
        for t in range(tiles):
            n = min(32, S - 32 * t)
            for slot, (et, p) in enumerate(config(t, 0)):
                eng.set_effect(slot, et, p, first_stream=32 * t, n_streams=n)
        at = 0
        counts = []
        for b, n in enumerate(blocks):
            if b == 2:
                for t in range(0, tiles, 3):
                    m = min(32, S - 32 * t)
                    for slot, (et, p) in enumerate(config(t, 1)):
                        eng.set_effect(slot, et, p, first_stream=32 * t, n_streams=m)
            before = eng.launch_count
            y[:, at:at + n] = eng.mix(np.ascontiguousarray(x[:, at:at + n]))
            counts.append(eng.launch_count - before)
            at += n
    assert [c for c, n in zip(counts, blocks) if n >= 2] == [1, 1, 1, 1], counts
    for s in (0, 31, 32, 95, 96, 100, 640, 1000, S - 1):
        t = s // 32
        script = []
        for slot, (et, p) in enumerate(config(t, 0)):
            script += [("type", slot, et), ("props", slot, p)]
        script += [("apply",)]
        for b, n in enumerate(blocks):
            if b == 2 and t % 3 == 0:
                for slot, (et, p) in enumerate(config(t, 1)):
                    script += [("type", slot, et), ("props", slot, p)]
                script += [("apply",)]
            script += [("mix", n)]
        expect = H.run_script_orc(checker, F.stereo, 48000, 4, script, x[s])
        _assert_match(expect, y[s], True, f"class-per-tile stream {s}")


@pytest.mark.parametrize("fmt", [F.mono, F.stereo])
def test_class_per_tile_launch_on_a_relay_signature(checker, fmt):
    """Class per tile for a signature without a fused kernel (echo, null, reverb, flanger): relay_multi_kernel, one
    launch per block, slot positions compacted (pending bits follow), 12 classes, ragged last tile."""
    lib = _lib()
    names = ox.reverb_preset_names(lib=lib)
    tiles = 12
    S = tiles * 32 - 9
    C = ox.channel_count(fmt)
    blocks = [1024, 333, 1024, 640]
    total = sum(blocks)
    x = np.stack([H.noise(5000 + s, C, total) for s in range(S)])
    y = np.empty_like(x)

    def config(t, phase):
        g, n = names[(5 * t + 1 + 17 * phase) % len(names)]
        return [(T.echo, ox.default_props(T.echo, lib=lib, delay_=0.02 + 0.004 * t, feedback_=0.3 + 0.02 * t)), (T.null, None),
                (T.eax_reverb, ox.reverb_preset(g, n, lib=lib)),
                (T.flanger, ox.default_props(T.flanger, lib=lib, rate_=0.2 + 0.05 * t, feedback_=-0.5 + 0.05 * t + 0.1 * phase))]

    with ox.Engine(S, fmt, 48000, 4, lib=lib) as eng:
        for t in range(tiles):
            n = min(32, S - 32 * t)
            for slot, (et, p) in enumerate(config(t, 0)):
                eng.set_effect(slot, et, p, first_stream=32 * t, n_streams=n)
        at = 0
        for b, n in enumerate(blocks):
            if b == 2:
                for t in range(0, tiles, 2):
                    m = min(32, S - 32 * t)
                    for slot, (et, p) in enumerate(config(t, 1)):
                        if et in (T.eax_reverb, T.flanger):
                            eng.set_effect(slot, et, p, first_stream=32 * t, n_streams=m)
            y[:, at:at + n] = eng.mix(np.ascontiguousarray(x[:, at:at + n]))
            at += n
        assert eng.launch_count == len(blocks), eng.launch_count
    for s in (0, 31, 32, 64, 100, 200, S - 1):
        t = s // 32
        script = []
        for slot, (et, p) in enumerate(config(t, 0)):
            script += [("type", slot, et)] + ([("props", slot, p)] if p is not None else [])
        script += [("apply",)]
        for b, n in enumerate(blocks):
            if b == 2 and t % 2 == 0:
                for slot, (et, p) in enumerate(config(t, 1)):
                    if et in (T.eax_reverb, T.flanger):
                        script += [("props", slot, p)]
                script += [("apply",)]
            script += [("mix", n)]
        expect = H.run_script_orc(checker, fmt, 48000, 4, script, x[s])
        _assert_match(expect, y[s], True, f"relay class-per-tile stream {s}")


def test_stream_major_bus_reduce_is_exact_enough_and_deterministic():
    """The coalesced two-pass reduction (row groups of 128 streams, then the groups): equals the float64 sum
    of the per-stream outputs within 1e-5 * sqrt(S) and is bit-identical from call to call."""
    import torch
    lib = _lib()
    S, n = 1000, 512  # not a multiple of the row group: the last group is short
    x = np.stack([H.noise(s, 2, n) for s in range(S)])
    with ox.Engine(S, F.stereo, 48000, 1, lib=lib) as eng:
        eng.set_effect(0, T.echo)
        xd = torch.from_numpy(x).cuda()
        yd = torch.empty_like(xd)
        eng.mix(xd, yd, frames=n)
        a = torch.zeros(n, 2, device="cuda")
        b = torch.zeros(n, 2, device="cuda")
        eng.reduce_bus(n, yd, a)
        eng.reduce_bus(n, yd, b)
        torch.cuda.synchronize()
    want = yd.cpu().numpy().astype(np.float64).sum(axis=0)
    assert np.max(np.abs(a.cpu().numpy() - want)) <= 1e-5 * np.sqrt(S)
    assert bool((a == b).all())


@pytest.mark.parametrize("streams", [7000, 96])
def test_mix_bus_epilogue_is_the_sum_of_the_outputs(streams):
    """oalsfx_engine_mix_bus (mix + the deterministic bus reduction in one call): the output rows are bit-identical to a
    plain mix, the bus equals the float64 sum of the rows within 1e-5 * sqrt(S), and it is bit-identical from run to run."""
    import torch
    lib = _lib()
    n = 512
    chain = [T.equalizer, T.chorus, T.echo, T.eax_reverb]
    x = torch.from_numpy(np.stack([H.noise(s, 2, 2 * n) for s in range(streams)])).cuda()  # independent streams: |bus| ~ sqrt(S)
    outs = []
    for use_bus in (False, True, True):
        with ox.Engine(streams, F.stereo, 48000, 4, lib=lib) as eng:
            for i, t in enumerate(chain):
                eng.set_effect(i, t)
            ys, buses = [], []
            for b in range(2):
                xb = x[:, b * n:(b + 1) * n].contiguous()
                y = torch.empty_like(xb)
                bus = torch.zeros(n, 2, device="cuda")
                if use_bus:
                    eng.mix_bus(xb, y, n, bus)
                else:
                    eng.mix(xb, y, frames=n)
                torch.cuda.synchronize()
                ys.append(y)
                buses.append(bus)
            kernel = eng.last_kernel
        outs.append((torch.cat(ys, dim=1), torch.cat(buses, dim=0), kernel))
    plain, first, second = outs
    assert bool((plain[0] == first[0]).all()) and bool((first[0] == second[0]).all())
    assert bool((first[1] == second[1]).all()), "the bus differs from run to run"
    want = first[0].cpu().numpy().astype(np.float64).sum(axis=0)
    assert np.max(np.abs(first[1].cpu().numpy() - want)) <= 1e-5 * np.sqrt(streams)


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()
