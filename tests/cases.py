"""The parity case matrix (SURVEY.md section 4 "recommended matrix"), shared by the CPU (emu) and
GPU test modules.  Each case is (name, channel_format, rate, effect_count, script, input).

The reference has no golden vectors of its own; the expected output of every case is
produced by replaying the same script through the compiled reference (oracle/_ref) or, where that
is absent, the restatement (oracle/_build) which is itself pinned to tests/golden/*.npz.
"""
import harness as H
from oalsfxpp_b200.props import ChannelFormat as F, EffectType as T


def _p(lib):
    """props helpers bound to whatever engine library the test session uses for defaults/presets."""
    from oalsfxpp_b200 import props as P

    def default(effect_type, **kw):
        return P.default_props(effect_type, lib=lib, **kw)

    def preset(group, name, **kw):
        return P.reverb_preset(group, name, lib=lib, **kw)

    return default, preset


ALL_TYPES = [t for t in T]
FORMATS = [F.mono, F.stereo, F.quad, F.five_point_one, F.five_point_one_rear, F.six_point_one, F.seven_point_one]


def basic_cases(lib, frames=6000):
    """Every effect type x every channel format, defaults, 1024-frame blocks."""
    for t in ALL_TYPES:
        for fmt in FORMATS:
            ch = H.channel_count(fmt)
            x = H.noise(1, ch, frames)
            yield (f"{t.name}-{fmt.name}", fmt, 48000, 1, H.simple_script([(t, None)], H.blocks_of(frames, 1024)), x)


def rate_and_block_cases(lib, frames=5000):
    """Sampling rates 8 k / 96 k and odd block partitions (the reverb's output depends on them)."""
    for rate in (8000, 96000, 44100):
        for t in (T.eax_reverb, T.reverb, T.echo, T.chorus, T.flanger, T.ring_modulator, T.distortion, T.equalizer,
                  T.compressor):
            x = H.noise(2, 2, frames)
            yield (f"{t.name}-stereo-{rate}", F.stereo, rate, 1,
                   H.simple_script([(t, None)], H.blocks_of(frames, 1024)), x)
    for block in (1, 2, 129, 480, 2048, 5000):
        n = 700 if block == 1 else frames
        for t in (T.eax_reverb, T.echo, T.equalizer):
            x = H.noise(3, 1, n)
            yield (f"{t.name}-mono-block{block}", F.mono, 48000, 1,
                   H.simple_script([(t, None)], H.blocks_of(n, block)), x)
    # ragged partition
    parts = [7, 1024, 300, 2048, 1, 1, 513, 128, 127, 851]
    x = H.noise(4, 2, sum(parts))
    yield ("eax_reverb-stereo-ragged", F.stereo, 48000, 1, H.simple_script([(T.eax_reverb, None)], parts), x)


def property_cases(lib, frames=6000):
    """Minimum / maximum / unusual property values per effect, plus reverb presets."""
    default, preset = _p(lib)
    blocks = H.blocks_of(frames, 1024)
    x2 = H.noise(5, 2, frames)
    x1 = H.noise(6, 1, frames)
    variants = [
        ("chorus-sin", T.chorus, default(T.chorus, waveform_=0)),
        ("chorus-max", T.chorus, default(T.chorus, rate_=10.0, depth_=1.0, feedback_=1.0, phase_=180)),
        ("chorus-negphase", T.chorus, default(T.chorus, phase_=-135, feedback_=-1.0, depth_=0.5)),
        ("chorus-rate0", T.chorus, default(T.chorus, rate_=0.0)),
        ("chorus-delay0", T.chorus, default(T.chorus, delay_=0.0)),
        ("flanger-sin", T.flanger, default(T.flanger, waveform_=0)),
        ("flanger-max", T.flanger, default(T.flanger, rate_=10.0, delay_=0.004, feedback_=1.0, phase_=-180)),
        ("flanger-min", T.flanger, default(T.flanger, rate_=0.05, delay_=0.0, depth_=0.0)),
        ("compressor-off", T.compressor, default(T.compressor, on_off_=False)),
        ("dialog-half", T.dedicated_dialog, default(T.dedicated_dialog, gain_=0.5)),
        ("lfe-half", T.dedicated_low_frequency, default(T.dedicated_low_frequency, gain_=0.5)),
        ("distortion-max", T.distortion, default(T.distortion, edge_=1.0, gain_=1.0, low_pass_cutoff_=24000.0,
                                                 eq_center_=24000.0, eq_bandwidth_=24000.0)),
        ("distortion-min", T.distortion, default(T.distortion, edge_=0.0, gain_=0.01, low_pass_cutoff_=80.0,
                                                 eq_center_=80.0, eq_bandwidth_=80.0)),
        ("echo-max", T.echo, default(T.echo, delay_=0.207, lr_delay_=0.404, damping_=0.99, feedback_=1.0, spread_=1.0)),
        ("echo-min", T.echo, default(T.echo, delay_=0.0, lr_delay_=0.0, damping_=0.0, feedback_=0.0, spread_=0.0)),
        ("equalizer-max", T.equalizer, default(T.equalizer, low_gain_=7.943, mid1_gain_=7.943, mid2_gain_=7.943,
                                               high_gain_=7.943, mid1_width_=0.01, mid2_width_=0.01)),
        ("equalizer-min", T.equalizer, default(T.equalizer, low_gain_=0.126, mid1_gain_=0.126, mid2_gain_=0.126,
                                               high_gain_=0.126, low_cutoff_=50.0, high_cutoff_=16000.0)),
        ("ringmod-saw", T.ring_modulator, default(T.ring_modulator, waveform_=1, frequency_=8000.0)),
        ("ringmod-square", T.ring_modulator, default(T.ring_modulator, waveform_=2, high_pass_cutoff_=0.0)),
        ("ringmod-freq0", T.ring_modulator, default(T.ring_modulator, frequency_=0.0, high_pass_cutoff_=24000.0)),
        ("reverb-density0", T.eax_reverb, default(T.eax_reverb, density_=0.0, diffusion_=0.0)),
        ("reverb-nodelay", T.eax_reverb, default(T.eax_reverb, reflections_delay_=0.0, late_reverb_delay_=0.0)),
        ("reverb-maxdelay", T.eax_reverb, default(T.eax_reverb, reflections_delay_=0.3, late_reverb_delay_=0.1,
                                                  decay_time_=20.0, density_=1.0)),
        ("reverb-mod", T.eax_reverb, default(T.eax_reverb, modulation_depth_=1.0, modulation_time_=0.04)),
        ("reverb-modslow", T.eax_reverb, default(T.eax_reverb, modulation_depth_=0.7, modulation_time_=4.0,
                                                 density_=0.0)),
        ("reverb-echo", T.eax_reverb, default(T.eax_reverb, echo_depth_=1.0, echo_time_=0.075)),
        ("reverb-pan", T.eax_reverb, default(T.eax_reverb, reflections_pan_=(0.3, -0.2, 0.5),
                                             late_reverb_pan_=(-0.9, 0.1, -0.4))),
        ("reverb-lfhf", T.eax_reverb, default(T.eax_reverb, gain_lf_=0.3, gain_hf_=0.2, decay_lf_ratio_=2.0,
                                              decay_hf_ratio_=2.0, decay_hf_limit_=False)),
        ("reverb-lfshort", T.eax_reverb, default(T.eax_reverb, decay_lf_ratio_=0.1, decay_hf_ratio_=0.1)),
        ("reverb-lflong-hfshort", T.reverb, default(T.reverb, decay_lf_ratio_=1.7, decay_hf_ratio_=0.4)),
        ("reverb-hfeq", T.reverb, default(T.reverb, decay_lf_ratio_=0.5, decay_hf_ratio_=1.0, decay_hf_limit_=False)),
    ]
    for group, name in (("Default", "generic"), ("Default", "padded_cell"), ("Default", "forest"), ("Default", "dizzy"),
                        ("Default", "psychotic"), ("Default", "underwater"), ("Castle", "courtyard"),
                        ("Dome", "saint_pauls"), ("Pipe", "resonant"), ("Mood", "heaven"), ("City", "abandoned")):
        variants.append((f"preset-{group}-{name}", T.eax_reverb, preset(group, name)))
    for name, t, props in variants:
        yield (name + "-stereo", F.stereo, 48000, 1, H.simple_script([(t, props)], blocks), x2)
    for name, t, props in variants[::3]:
        yield (name + "-mono96k", F.mono, 96000, 1, H.simple_script([(t, props)], blocks), x1)


def chain_cases(lib, frames=6000):
    """Multi-slot configurations incl. the two BASELINE chains, sends, and other inputs."""
    default, preset = _p(lib)
    blocks = H.blocks_of(frames, 1024)
    chain = [(T.equalizer, None), (T.chorus, None), (T.echo, None), (T.eax_reverb, None)]
    chain2 = [(T.flanger, None), (T.ring_modulator, None), (T.distortion, None), (T.compressor, None)]
    yield ("chain-stereo", F.stereo, 48000, 4, H.simple_script(chain, blocks), H.noise(7, 2, frames))
    yield ("chain-stereo-sine", F.stereo, 48000, 4, H.simple_script(chain, blocks), H.sine(7, 2, frames, 48000))
    yield ("chain-stereo-impulse", F.stereo, 48000, 4, H.simple_script(chain, blocks), H.impulse(2, frames))
    yield ("chain-stereo-burst", F.stereo, 48000, 4, H.simple_script(chain, blocks), H.burst(7, 2, frames, 600))
    # the fused fast kernels' other paths: modulated reverb (no prefetch), tiny taps (direct reads), EAX off
    for tag, rv_type, rv in (("underwater", T.eax_reverb, preset("Default", "underwater")),
                             ("psychotic", T.eax_reverb, preset("Default", "psychotic")),
                             ("tinytaps", T.eax_reverb, default(T.eax_reverb, density_=0.0, reflections_delay_=0.0,
                                                                late_reverb_delay_=0.0)),
                             ("std-reverb", T.reverb, preset("Default", "hangar"))):
        slots = [(T.equalizer, default(T.equalizer, mid2_gain_=2.0)), (T.chorus, default(T.chorus, waveform_=0)),
                 (T.echo, default(T.echo, delay_=0.0001, lr_delay_=0.0)), (rv_type, rv)]
        yield (f"chain-stereo-{tag}", F.stereo, 48000, 4, H.simple_script(slots, blocks), H.noise(30, 2, frames))
    yield ("chain-stereo-flanger", F.stereo, 48000, 4,
           H.simple_script([(T.equalizer, None), (T.flanger, None), (T.echo, None), (T.eax_reverb, None)], blocks),
           H.noise(31, 2, frames))
    yield ("chain-mono", F.mono, 48000, 4, H.simple_script(chain, blocks), H.noise(8, 1, frames))
    yield ("chain-5.1", F.five_point_one, 48000, 4, H.simple_script(chain, blocks), H.noise(9, 6, frames))
    yield ("chain2-mono96k", F.mono, 96000, 4, H.simple_script(chain2, blocks), H.noise(10, 1, frames))
    yield ("chain2-stereo", F.stereo, 48000, 4, H.simple_script(chain2, blocks), H.noise(11, 2, frames))
    yield ("chain2-sin-mono96k", F.mono, 96000, 4,
           H.simple_script([(T.flanger, default(T.flanger, waveform_=0))] + chain2[1:], blocks), H.noise(10, 1, frames))
    # partly filled / mixed slot sets (multi-pass path)
    yield ("null-echo-null-reverb", F.stereo, 48000, 4,
           H.simple_script([(T.null, None), (T.echo, None), (T.null, None), (T.reverb, None)], blocks),
           H.noise(12, 2, frames))
    yield ("reverb-reverb", F.quad, 44100, 2,
           H.simple_script([(T.eax_reverb, preset("Default", "cave")), (T.reverb, preset("Default", "room"))], blocks),
           H.noise(13, 4, frames))
    yield ("dedicated-comp-dist", F.seven_point_one, 48000, 3,
           H.simple_script([(T.dedicated_dialog, None), (T.compressor, None), (T.distortion, None)], blocks),
           H.noise(14, 8, frames))
    # send gains and shelf filters (direct + aux)
    sends = {-1: (0.8, 0.5, 1.0), 0: (0.7, 1.0, 0.4), 2: (1.0, 0.25, 0.5)}
    yield ("chain-stereo-sends", F.stereo, 48000, 4, H.simple_script(chain, blocks, sends) , H.noise(15, 2, frames))
    yield ("echo-mono-direct-lf", F.mono, 48000, 1,
           H.simple_script([(T.echo, None)], blocks, {-1: (1.0, 1.0, 0.3)}), H.noise(16, 1, frames))
    yield ("eq-stereo-aux-bandpass-block1", F.stereo, 48000, 1,
           H.simple_script([(T.equalizer, None)], [1] * 300, {0: (0.9, 0.3, 0.6)}), H.noise(17, 2, 300))


def schedule_cases(lib, blocks_n=24):
    """Parameter changes between blocks: cfg2's schedule, effect-type swaps, send changes."""
    default, preset = _p(lib)
    chain = [(T.equalizer, None), (T.chorus, None), (T.echo, None), (T.eax_reverb, None)]
    script = []
    for i, (t, _) in enumerate(chain):
        script.append(("type", i, t))
    script.append(("apply",))
    for b in range(blocks_n):
        rv = default(T.eax_reverb, gain_=0.20 + 0.10 * ((b % 4) / 4.0),
                     reflections_delay_=(0.012 if (b // 16) % 2 else 0.007))
        eq = default(T.equalizer, mid1_gain_=1.0 + 0.5 * ((b % 8) / 8.0))
        script += [("props", 3, rv), ("props", 0, eq), ("apply",), ("mix", 1024)]
    yield ("cfg2-schedule", F.stereo, 48000, 4, script, H.noise(20, 2, 1024 * blocks_n))

    # tap changes every block (cross-fade restarted while still fading) + modulation time changes
    script = [("type", 0, T.eax_reverb), ("apply",)]
    for b in range(12):
        rv = default(T.eax_reverb, reflections_delay_=0.002 * (b % 5), late_reverb_delay_=0.003 * (b % 3),
                     density_=0.2 + 0.1 * (b % 4), modulation_depth_=0.5, modulation_time_=0.1 + 0.05 * (b % 3))
        script += [("props", 0, rv), ("apply",), ("mix", 100 if b % 2 else 700)]
    yield ("reverb-tap-churn", F.stereo, 48000, 1, script, H.noise(21, 2, 6 * 100 + 6 * 700))

    # effect type swapped mid-stream (state reset), back and forth, plus reverb <-> eax_reverb
    script = [("type", 0, T.echo), ("type", 1, T.reverb), ("apply",), ("mix", 1500)]
    script += [("type", 0, T.chorus), ("apply",), ("mix", 1000)]
    script += [("type", 0, T.echo), ("type", 1, T.eax_reverb), ("apply",), ("mix", 2048)]
    script += [("type", 1, T.null), ("apply",), ("mix", 500)]
    script += [("type", 1, T.eax_reverb), ("apply",), ("mix", 952)]
    yield ("type-swaps", F.stereo, 48000, 2, script, H.noise(22, 2, 6000))

    # same props re-applied (no update must happen), props set without apply (must not take effect)
    script = [("type", 0, T.eax_reverb), ("apply",), ("mix", 300),
              ("props", 0, default(T.eax_reverb, gain_=0.1)), ("mix", 300),
              ("apply",), ("mix", 300), ("apply",), ("mix", 300),
              ("props", 0, default(T.eax_reverb, gain_=5.0, density_=-3.0)), ("apply",), ("mix", 800)]  # clamped
    yield ("deferred-semantics", F.stereo, 48000, 1, script, H.noise(23, 2, 2000))

    # send changes mid-stream incl. the aux quirk (props written directly, seen at the next refresh)
    script = [("type", 0, T.echo), ("type", 1, T.equalizer), ("apply",), ("mix", 512),
              ("send", 0, (0.5, 0.5, 1.0)), ("mix", 512),            # not yet visible
              ("apply",), ("mix", 512),                              # flagged -> visible
              ("send", -1, (0.25, 1.0, 0.5)), ("apply",), ("mix", 512),
              ("send", 1, (2.0, 1.0, 1.0)), ("send", -1, (1.0, 1.0, 1.0)), ("apply",), ("mix", 512),
              ("send", 0, (1.0, 1.0, 1.0)), ("send", 1, (1.0, 1.0, 1.0)), ("apply",), ("mix", 512)]
    yield ("send-changes", F.stereo, 48000, 2, script, H.noise(24, 2, 3072))


def all_cases(lib, quick=False):
    gens = [basic_cases(lib, 3000 if quick else 6000), rate_and_block_cases(lib, 3000 if quick else 5000),
            property_cases(lib, 3000 if quick else 6000), chain_cases(lib, 3000 if quick else 6000),
            schedule_cases(lib, 20 if quick else 24)]
    for g in gens:
        yield from g
