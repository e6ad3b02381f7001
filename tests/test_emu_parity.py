"""CPU-only parity: the engine's host logic + the __host__ __device__ kernel bodies (compiled for the
CPU into tests/_build/liboalsfx_emu.so, test infrastructure) against the CPU checker.  Bit-exact.

These tests prove the algorithm restatement and all host-side logic without a GPU; the same cases
run on the device in test_gpu_parity.py.
"""
import numpy as np
import pytest

import cases
import harness as H
import oalsfxpp_b200 as ox
from oalsfxpp_b200 import ChannelFormat as F, EffectType as T

CASES = list(cases.all_cases(H.emu_lib(), quick=True))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_api_shell_matches_checker_bit_for_bit(case, checker):
    name, fmt, rate, effect_count, script, x = case
    expect = H.run_script_orc(checker, fmt, rate, effect_count, script, x)
    got = H.run_script_orc(H.api_shim("emu"), fmt, rate, effect_count, script, x)
    # NaN payload/sign bits are not part of the contract (the 8 kHz cases drive the shelf filters
    # unstable in the reference itself): NaNs must sit at the same samples, everything else is bit-equal.
    assert np.array_equal(expect, got, equal_nan=True), H.max_abs_diff(expect, got)


def _expected(checker, fmt, rate, slots, blocks, x):
    script = H.simple_script(slots, blocks)
    return H.run_script_orc(checker, fmt, rate, len(slots), script, x)


@pytest.mark.parametrize("layout", [ox.LAYOUT_STREAM_MAJOR, ox.LAYOUT_TILED])
def test_batched_engine_heterogeneous_streams(checker, layout):
    """70 streams (ragged last tile), three different slot signatures and parameter sets interleaved
    inside tiles; every stream must equal its own single-instance checker run."""
    lib = H.emu_lib()
    S, frames, block = 70, 2500, 1024
    blocks = H.blocks_of(frames, block)
    default = lambda t, **kw: ox.default_props(t, lib=lib, **kw)
    configs = [
        [(T.equalizer, None), (T.chorus, None), (T.echo, None), (T.eax_reverb, None)],
        [(T.flanger, default(T.flanger, waveform_=0)), (T.ring_modulator, None), (T.distortion, None), (T.compressor, None)],
        [(T.eax_reverb, ox.reverb_preset("Default", "forest", lib=lib)), (T.null, None), (T.echo, default(T.echo, delay_=0.05)), (T.null, None)],
    ]
    which = [(s * 7 + s // 5) % 3 for s in range(S)]
    x = np.stack([H.noise(100 + s, 2, frames) for s in range(S)])
    with ox.Engine(S, F.stereo, 48000, 4, lib=lib) as eng:
        for s in range(S):
            for slot, (t, p) in enumerate(configs[which[s]]):
                eng.set_effect(slot, t, p, first_stream=s, n_streams=1)
        y = np.empty_like(x)
        pos = 0
        for n in blocks:
            xb = np.ascontiguousarray(x[:, pos:pos + n])
            if layout == ox.LAYOUT_TILED:
                pad = np.zeros((eng.padded_streams, n, 2), np.float32)
                pad[:S] = xb
                tiled = np.ascontiguousarray(pad.reshape(-1, 32, n, 2).transpose(0, 2, 3, 1))
                out = eng.mix(tiled, layout=layout)
                yb = out.transpose(0, 3, 1, 2).reshape(-1, n, 2)[:S]
            else:
                yb = eng.mix(xb)
            y[:, pos:pos + n] = yb
            pos += n
    for s in range(S):
        expect = _expected(checker, F.stereo, 48000, configs[which[s]], blocks, x[s])
        assert np.array_equal(expect, y[s]), (s, which[s], H.max_abs_diff(expect, y[s]))


def test_long_call_is_cut_into_2048_frame_blocks(checker):
    """One 5000-frame call = blocks 2048, 2048, 904 (reference: oalsfxpp.cpp:3818-3826)."""
    lib = H.emu_lib()
    x = np.stack([H.noise(s, 1, 5000) for s in range(3)])
    with ox.Engine(3, F.mono, 48000, 1, lib=lib) as eng:
        eng.set_effect(0, T.eax_reverb)
        y = eng.mix(x)
    for s in range(3):
        expect = _expected(checker, F.mono, 48000, [(T.eax_reverb, None)], [5000], x[s])
        assert np.array_equal(expect, y[s])
        different = _expected(checker, F.mono, 48000, [(T.eax_reverb, None)], H.blocks_of(5000, 1024), x[s])
        assert not np.array_equal(different, y[s])  # the partition matters for the reverb


def test_integer_state_is_bit_exact():
    """Ring offsets, reverb fade counter and modulator index follow the reference's formulas
    (offset_ += n, oalsfxpp.cpp:7853; fade 128 samples, :6118-6138; index % range, :7457)."""
    lib = H.emu_lib()
    rate, frames = 48000, [1024, 100, 2048, 77]
    mod_time = 0.25
    with ox.Engine(2, F.stereo, rate, 3, lib=lib) as eng:
        eng.set_effect(0, T.eax_reverb, ox.default_props(T.eax_reverb, lib=lib, modulation_depth_=0.3, modulation_time_=mod_time))
        eng.set_effect(1, T.ring_modulator)
        eng.set_effect(2, T.chorus)
        total = 0
        for n in frames:
            eng.mix(np.zeros((2, n, 2), np.float32))
            total += n
            st = eng.debug_state(1, 0)
            assert st["offset"] == total
            assert st["fade_count"] == min(total, 128)
            assert st["mod_index"] == total % int(mod_time * rate)
            step = int(np.float32(440.0) * np.float32(1 << 24) / np.float32(rate))
            assert eng.debug_state(1, 1)["ring_mod_index"] == (total * step) & 0xFFFFFF
            assert eng.debug_state(0, 2)["offset"] == total


def test_bus_reduce_matches_float64_sum():
    lib = H.emu_lib()
    S, n = 40, 256
    x = np.stack([H.noise(s, 2, n) for s in range(S)])
    with ox.Engine(S, F.stereo, 48000, 1, lib=lib) as eng:
        eng.set_effect(0, T.echo)
        y = eng.mix(x)
        bus = np.zeros((n, 2), np.float32)
        eng.reduce_bus(n, y, bus)
    want = y.astype(np.float64).sum(axis=0)
    assert np.max(np.abs(bus - want)) <= 1e-5 * np.sqrt(S)


def test_engine_argument_errors():
    lib = H.emu_lib()
    with pytest.raises(ox.OalsfxError) as e:
        ox.Engine(1, 0, 48000, 1, lib=lib)
    assert e.value.message == "Invalid channel format."
    with pytest.raises(ox.OalsfxError) as e:
        ox.Engine(1, F.mono, 7999, 1, lib=lib)
    assert e.value.message == "Sampling rate is out of range."
    with pytest.raises(ox.OalsfxError) as e:
        ox.Engine(1, F.mono, 48000, 5, lib=lib)
    assert e.value.message == "Effect count is out of range."
    with ox.Engine(4, F.mono, 48000, 2, lib=lib) as eng:
        with pytest.raises(ox.OalsfxError):
            eng.set_effect(2, T.echo)
        with pytest.raises(ox.OalsfxError):
            eng.set_effect(0, T.echo, first_stream=3, n_streams=2)
        assert eng.mix(np.zeros((4, 0, 1), np.float32), frames=0).size == 0  # zero frames is a no-op


def test_host_buffer_mix_is_sliced_without_changing_results(checker):
    """>= 4096 streams in one uniform group: the host-buffer path uploads / mixes / downloads tile slices
    on three overlapped streams (engine.cpp).  Results must not depend on the slicing, including the
    ragged last tile and a parameter change (pending update) between blocks."""
    lib = H.emu_lib()
    S, block = 4100, 300
    x = np.stack([H.noise(s % 7, 2, 2 * block) for s in range(S)])
    with ox.Engine(S, F.stereo, 48000, 2, lib=lib) as eng:
        eng.set_effect(0, T.echo)
        eng.set_effect(1, T.eax_reverb)
        y0 = eng.mix(np.ascontiguousarray(x[:, :block]))
        eng.set_effect(1, T.eax_reverb, ox.default_props(T.eax_reverb, lib=lib, gain_=0.2, reflections_delay_=0.01))
        y1 = eng.mix(np.ascontiguousarray(x[:, block:]))
    y = np.concatenate([y0, y1], axis=1)
    script = [("type", 0, T.echo), ("type", 1, T.eax_reverb), ("apply",), ("mix", block),
              ("props", 1, ox.default_props(T.eax_reverb, lib=lib, gain_=0.2, reflections_delay_=0.01)), ("apply",),
              ("mix", block)]
    for s in (0, 1, 6, 2047, 2048, 4095, 4096, 4099):
        expect = H.run_script_orc(checker, F.stereo, 48000, 2, script, x[s])
        assert np.array_equal(expect, y[s]), s


def test_many_parameter_sets_table_mode(checker):
    launches, nblocks = H.many_parameter_sets(H.emu_lib(), checker, 70, 1500, 512, range(70), exact_all=True)
    # two kind signatures -> two groups of (dry+slot, 3 accumulate passes): not one launch per parameter set
    assert launches <= nblocks * 2 * 4


# ---- the device-only kernel families, emulated by the CPU backend (tests/emu/host_backend.cpp): what is checked here
# is the engine's host logic around them -- selection, slot compaction, legality checks, class tables, fallbacks ----
def _gpu_scenarios(monkeypatch):
    import test_gpu_parity as G
    monkeypatch.setattr(G, "_lib", H.emu_lib)
    return G


@pytest.mark.parametrize("sig", [0, 1, 2, 6, 8, 10])
def test_relay_selection_and_slot_compaction(checker, sig, monkeypatch):
    _gpu_scenarios(monkeypatch).test_relay_pipeline_runs_any_signature_in_one_launch(checker, sig)


@pytest.mark.parametrize("case", ["eax-mono", "short-first-blocks", "dense"])
def test_span_selection_and_legality(checker, case, monkeypatch):
    _gpu_scenarios(monkeypatch).test_span_kernel_on_steady_state_blocks(checker, case)


@pytest.mark.parametrize("case", ["default", "flanger-96k", "forest-short-echo", "short-first-blocks", "standard-reverb"])
def test_span_chain_schedule_and_legality(checker, case, monkeypatch):
    """The chain's span kernel: its phase bodies and its three-stage pipeline schedule (emulated serially in an
    adversarial order by the CPU backend) against the checker, and that the schedule was really taken."""
    before = H.emu_lib().oalsfx_emu_span_streams()
    _gpu_scenarios(monkeypatch).test_span_chain_on_steady_state_blocks(checker, case)
    assert H.emu_lib().oalsfx_emu_span_streams() > before


@pytest.mark.parametrize("case", ["default", "flanger-96k", "forest-short-echo", "short-first-blocks", "standard-reverb", "eax-mono"])
def test_span_bulk_schedule_and_legality(checker, case, monkeypatch):
    """span_bulk_kernel's timing (row loads issued one / two iterations ahead, row stores landing one / two behind),
    emulated at both extremes of when the copies may take effect."""
    before = H.emu_lib().oalsfx_emu_span_bulk_streams()
    _gpu_scenarios(monkeypatch).test_span_bulk_kernel(checker, case, monkeypatch)
    assert H.emu_lib().oalsfx_emu_span_bulk_streams() > before


def test_span_schedule_was_taken_for_the_single_reverb_slot(checker, monkeypatch):
    before = H.emu_lib().oalsfx_emu_span_streams()
    _gpu_scenarios(monkeypatch).test_span_kernel_on_steady_state_blocks(checker, "eax-stereo")
    assert H.emu_lib().oalsfx_emu_span_streams() > before


def test_class_per_tile_tables_chain(checker, monkeypatch):
    _gpu_scenarios(monkeypatch).test_class_per_tile_launch_many_presets(checker)


def test_class_per_tile_tables_relay(checker, monkeypatch):
    _gpu_scenarios(monkeypatch).test_class_per_tile_launch_on_a_relay_signature(checker, F.mono)


def test_mix_bus_host_logic():
    """oalsfx_engine_mix_bus on the CPU backend (a call longer than 2048 frames is cut into blocks)."""
    lib = H.emu_lib()
    chain = [T.equalizer, T.chorus, T.echo, T.eax_reverb]
    for streams, n in ((200, 12), (40, 2500)):
        x = np.stack([H.noise(s % 16, 2, n) for s in range(streams)])
        with ox.Engine(streams, F.stereo, 48000, 4, lib=lib) as eng:
            for i, t in enumerate(chain):
                eng.set_effect(i, t)
            y = np.empty_like(x)
            bus = np.zeros((n, 2), dtype=np.float32)
            eng.mix_bus(x, y, n, bus)
            kernel = eng.last_kernel
        with ox.Engine(streams, F.stereo, 48000, 4, lib=lib) as eng:
            for i, t in enumerate(chain):
                eng.set_effect(i, t)
            y2 = eng.mix(x)
        assert np.array_equal(y.view(np.uint32), y2.view(np.uint32))
        want = y.astype(np.float64).sum(axis=0)
        assert np.max(np.abs(bus - want)) <= 1e-5 * np.sqrt(streams)


def test_placement_helper_and_class_per_tile_selection(checker, monkeypatch):
    _gpu_scenarios(monkeypatch).test_placement_keeps_arbitrary_presets_on_the_class_per_tile_launch(checker)


def test_waveshaper_expectation(monkeypatch):
    """The waveshaper scenario of the GPU suite on the CPU backend (the `/` operator): pins the numpy formula the GPU's
    batched division is held against."""
    _gpu_scenarios(monkeypatch).test_waveshaper_divisions_are_ieee_divisions()


@pytest.mark.parametrize("case", ["chain-stereo", "echo-mono", "chain-5.1"])
def test_send_filter_relay_selection(checker, case, monkeypatch):
    _gpu_scenarios(monkeypatch).test_send_shelf_filters_run_in_one_launch(checker, case)
