"""Shared test plumbing: builds/loads the CPU checkers and drives "scripts" through them.

Checkers (all test infrastructure, see oracle/README.md):
  * ``ref``    oracle/_ref/liboalsfx_ref.so      -- the unmodified reference behind a C shim
  * ``oracle`` oracle/_build/liboalsfx_oracle.so -- the hand-written CPU restatement
Both export the same ``orc_*`` ABI (oracle/ref_shim.cpp).

Systems under test:
  * ``emu``  tests/_build/liboalsfx_emu.so  -- engine host logic + kernel bodies on the CPU (no GPU)
  * ``cuda`` oalsfxpp_b200/liboalsfx_b200.so -- the product
Both export the C ABI of include/oalsfx_engine.h; the drop-in ``oalsfxpp::Api`` class of either is
reached through the very same shim source the reference is wrapped with (oracle/ref_shim.cpp,
compiled against include/oalsfxpp.h instead of the reference's header).

A *script* is a list of ops replayed identically against a checker and a system under test:
  ("type", slot, EffectType)            Api::set_effect_type
  ("props", slot, EffectProps)          Api::set_effect_props
  ("send", index, (gain, hf, lf))       Api::set_send_props   (index < 0: direct)
  ("apply",)                            Api::apply_changes
  ("mix", frames)                       Api::mix on the next `frames` frames of the input
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oalsfxpp_b200 import engine as _eng  # noqa: E402
from oalsfxpp_b200.props import EffectProps, EffectType, channel_count  # noqa: E402

REF_SRC = "/root/reference/src"
SEED = 0x0A15F00D


def _run(cmd, **kw):
    subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, **kw)


class _BuildLock:
    """The lazy builds below may be reached by several pytest-xdist workers at once: one builds, the others wait."""

    def __enter__(self):
        import fcntl
        os.makedirs(os.path.join(ROOT, "tests", "_build"), exist_ok=True)
        self._f = open(os.path.join(ROOT, "tests", "_build", ".lock"), "w")
        fcntl.flock(self._f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self._f, fcntl.LOCK_UN)
        self._f.close()


def build_checkers():
    with _BuildLock():
        _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])


def build_emu():
    with _BuildLock():
        _run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])


def _bind_orc(lib):
    vp, i32 = C.c_void_p, C.c_int
    lib.orc_kind.restype = C.c_char_p
    lib.orc_create.argtypes = [i32, i32, i32]
    lib.orc_create.restype = vp
    lib.orc_destroy.argtypes = [vp]
    lib.orc_destroy.restype = None
    lib.orc_channel_count.argtypes = [vp]
    lib.orc_set_effect_type.argtypes = [vp, i32, i32]
    lib.orc_set_effect_props.argtypes = [vp, i32, vp]
    lib.orc_get_effect.argtypes = [vp, i32, C.POINTER(i32), vp]
    lib.orc_get_deferred_effect.argtypes = [vp, i32, C.POINTER(i32), vp]
    lib.orc_set_send_props.argtypes = [vp, i32, C.POINTER(C.c_float)]
    lib.orc_apply.argtypes = [vp]
    lib.orc_mix.argtypes = [vp, i32, vp, vp]
    lib.orc_noise.argtypes = [C.c_uint32, C.c_uint32, i32, i32, i32, vp]
    lib.orc_noise.restype = None
    lib.orc_bench.argtypes = [i32, i32, i32, i32, C.POINTER(i32), i32, i32, i32, C.c_uint32, C.POINTER(C.c_double)]
    lib.orc_bench.restype = C.c_double
    return lib


_CACHE = {}


def ref_lib(fast=False):
    """The compiled reference, or None when neither the mount nor a prebuilt .so is there."""
    key = "ref_fast" if fast else "ref"
    if key not in _CACHE:
        path = os.path.join(ROOT, "oracle", "_ref", "liboalsfx_ref_fast.so" if fast else "liboalsfx_ref.so")
        if not os.path.exists(path) and os.path.isdir(REF_SRC):
            build_checkers()
        _CACHE[key] = _bind_orc(C.CDLL(path)) if os.path.exists(path) else None
    return _CACHE[key]


def oracle_lib():
    if "oracle" not in _CACHE:
        path = os.path.join(ROOT, "oracle", "_build", "liboalsfx_oracle.so")
        with _BuildLock():
            _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
        _CACHE["oracle"] = _bind_orc(C.CDLL(path))
    return _CACHE["oracle"]


def checker_lib():
    """Strongest checker available: the real reference, else the restatement."""
    return ref_lib() or oracle_lib()


def emu_lib():
    if "emu" not in _CACHE:
        build_emu()
        _CACHE["emu"] = _eng.bind(C.CDLL(os.path.join(ROOT, "tests", "_build", "liboalsfx_emu.so")))
        _CACHE["emu"].oalsfx_emu_span_streams.restype = C.c_longlong
        _CACHE["emu"].oalsfx_emu_span_bulk_streams.restype = C.c_longlong
    return _CACHE["emu"]


def cuda_lib():
    return _eng.load_library()


def api_shim(kind):
    """oracle/ref_shim.cpp compiled against include/oalsfxpp.h and linked to the emu / cuda library:
    the drop-in proof for the C++ class."""
    key = "shim_" + kind
    if key not in _CACHE:
        out_dir = os.path.join(ROOT, "tests", "_build")
        os.makedirs(out_dir, exist_ok=True)
        out = os.path.join(out_dir, f"libapi_shim_{kind}.so")
        if kind == "emu":
            emu_lib()
            libdir, libname = out_dir, "oalsfx_emu"
        else:
            libdir, libname = os.path.join(ROOT, "oalsfxpp_b200"), "oalsfx_b200"
        src = os.path.join(ROOT, "oracle", "ref_shim.cpp")
        dep = os.path.join(libdir, f"lib{libname}.so")
        with _BuildLock():
            if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(dep)):
                tmp = out + f".{os.getpid()}.tmp"
                _run(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-pthread", "-I", os.path.join(ROOT, "include"),
                      "-o", tmp, src, "-L", libdir, "-l" + libname, "-Wl,-rpath," + libdir])
                os.replace(tmp, out)
        _CACHE[key] = _bind_orc(C.CDLL(out))
    return _CACHE[key]


# ---- inputs --------------------------------------------------------------------------------------
def _fmix32(h):
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


def noise(stream, channels, frames, first_frame=0, seed=SEED):
    """SURVEY.md 8d synthetic white noise, [frames][channels] float32 in [-0.5, 0.5)."""
    with np.errstate(over="ignore"):
        n = (np.arange(first_frame, first_frame + frames, dtype=np.uint32)[:, None] * np.uint32(0xC2B2AE35))
        c = (np.arange(channels, dtype=np.uint32)[None, :] * np.uint32(0x85EBCA6B))
        s = np.uint32((stream * 0x9E3779B9) & 0xFFFFFFFF)
        h = _fmix32(np.uint32(seed) ^ s ^ c ^ n)
    return (((h >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 8388608.0)) - np.float32(1.0)) * np.float32(0.5)


def sine(stream, channels, frames, rate):
    f = 110.0 * 2.0 ** ((stream % 48) / 12.0)
    t = np.arange(frames, dtype=np.float64)
    x = (0.5 * np.sin(2.0 * np.pi * f * t / rate)).astype(np.float32)
    return np.repeat(x[:, None], channels, axis=1).copy()


def impulse(channels, frames):
    x = np.zeros((frames, channels), dtype=np.float32)
    x[0, :] = 0.5
    return x


def burst(stream, channels, frames, active):
    x = noise(stream, channels, frames)
    x[active:, :] = 0.0
    return x


# ---- script replay -------------------------------------------------------------------------------
def run_script_orc(lib, channel_format, rate, effect_count, script, x):
    """Replay `script` through an orc_* library (reference, restatement or Api shim).  x: [frames][C]."""
    h = lib.orc_create(int(channel_format), rate, effect_count)
    assert h, "orc_create failed"
    try:
        ch = channel_count(channel_format)
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.zeros_like(x)
        pos = 0
        for op in script:
            if op[0] == "type":
                assert lib.orc_set_effect_type(h, op[1], int(op[2]))
            elif op[0] == "props":
                assert lib.orc_set_effect_props(h, op[1], C.addressof(op[2]))
            elif op[0] == "send":
                g = (C.c_float * 3)(*op[2])
                assert lib.orc_set_send_props(h, op[1], g)
            elif op[0] == "apply":
                assert lib.orc_apply(h)
            elif op[0] == "mix":
                n = op[1]
                assert lib.orc_mix(h, n, x[pos:].ctypes.data, y[pos:].ctypes.data)
                pos += n
            else:
                raise ValueError(op)
        assert pos * ch == x.size, (pos, x.shape)
        return y
    finally:
        lib.orc_destroy(h)


def simple_script(slots, blocks, sends=None):
    """Set the slot effects (type, props-or-None), optional sends, apply, then mix the blocks."""
    script = []
    for i, (etype, props) in enumerate(slots):
        script.append(("type", i, etype))
        if props is not None:
            script.append(("props", i, props))
    for index, triple in (sends or {}).items():
        script.append(("send", index, triple))
    script.append(("apply",))
    # Aux send props written directly only take effect with the next refresh; a second apply makes the
    # first mix see them in the reference as well (it flags the source as changed).
    script.extend(("mix", n) for n in blocks)
    return script


def blocks_of(total, block):
    out = [block] * (total // block)
    if total % block:
        out.append(total % block)
    return out


def max_abs_diff(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0


# ---- shared scenario: every stream its own parameter set (engine table mode) ---------------------------
def many_parameter_sets(lib, checker, S, frames, block, sample, exact_all):
    """Every stream its own parameter set (table mode: > 8 distinct classes, coefficient blocks in HBM), two
    different slot signatures, send filters on some streams, a parameter change in the middle."""
    import oalsfxpp_b200 as ox
    from oalsfxpp_b200 import ChannelFormat as F, EffectType as T
    default = lambda t, **kw: ox.default_props(t, lib=lib, **kw)
    blocks = blocks_of(frames, block)
    x = np.stack([noise(500 + s, 2, frames) for s in range(S)])

    def config(s, phase):
        if s % 5 == 4:  # a different signature in the same engine
            return [(T.flanger, default(T.flanger, rate_=0.1 + 0.01 * s)), (T.ring_modulator, default(T.ring_modulator, frequency_=300.0 + 7 * s)),
                    (T.distortion, default(T.distortion, edge_=0.1 + 0.005 * (s % 100))), (T.compressor, None)]
        return [(T.equalizer, default(T.equalizer, mid1_gain_=0.5 + 0.01 * s + 0.2 * phase)), (T.chorus, default(T.chorus, depth_=0.05 + 0.001 * s)),
                (T.echo, default(T.echo, delay_=0.02 + 0.001 * s, feedback_=0.3)),
                (T.eax_reverb if s % 2 else T.reverb, default(T.eax_reverb if s % 2 else T.reverb, gain_=0.1 + 0.002 * s,
                                                               decay_time_=0.5 + 0.01 * s + 0.3 * phase, reflections_delay_=0.001 * (s % 20)))]

    y = np.empty_like(x)
    with ox.Engine(S, F.stereo, 48000, 4, lib=lib) as eng:
        for s in range(S):
            for slot, (t, p) in enumerate(config(s, 0)):
                eng.set_effect(slot, t, p, first_stream=s, n_streams=1)
            if s % 7 == 0:
                eng.set_sends(direct=(0.8, 0.5, 1.0), aux=[(1.0, 0.25, 1.0)] * 4, first_stream=s, n_streams=1)
        pos = 0
        for b, n in enumerate(blocks):
            if b == 1:
                for s in range(S):
                    for slot, (t, p) in enumerate(config(s, 1)):
                        eng.set_effect(slot, t, p, first_stream=s, n_streams=1)
            y[:, pos:pos + n] = eng.mix(np.ascontiguousarray(x[:, pos:pos + n]))
            pos += n
        launches = eng.launch_count
    for s in sample:
        script = []
        for slot, (t, p) in enumerate(config(s, 0)):
            script += [("type", slot, t)] + ([("props", slot, p)] if p is not None else [])
        if s % 7 == 0:
            script += [("send", -1, (0.8, 0.5, 1.0))] + [("send", i, (1.0, 0.25, 1.0)) for i in range(4)]
        script += [("apply",)]
        for b, n in enumerate(blocks):
            if b == 1:
                for slot, (t, p) in enumerate(config(s, 1)):
                    script += [("type", slot, t)] + ([("props", slot, p)] if p is not None else [])
                script += [("apply",)]
            script += [("mix", n)]
        expect = run_script_orc(checker, F.stereo, 48000, 4, script, x[s])
        exact = s % 5 != 4  # the ring modulator's carrier is the device sinf on the GPU
        if exact or exact_all:
            assert np.array_equal(expect, y[s], equal_nan=True), (s, max_abs_diff(expect, y[s]))
        else:
            assert max_abs_diff(expect, y[s]) <= 1e-5, s
    return launches, len(blocks)

