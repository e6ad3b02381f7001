"""ctypes mirror of the reference's property PODs (reference: src/oalsfxpp.h:36-62, 65-530).

Field names and order follow the header so the bytes are exactly an ``oalsfxpp::EffectProps``
(108 bytes); defaults, clamping and presets are NOT re-implemented here, they come from the
library (``oalsfx_effect_defaults`` / ``oalsfx_effect_normalize`` / ``oalsfx_reverb_preset``).
"""
import ctypes as C
import enum


class ChannelFormat(enum.IntEnum):
    none = 0
    mono = 1
    stereo = 2
    quad = 3
    five_point_one = 4
    five_point_one_rear = 5
    six_point_one = 6
    seven_point_one = 7


class EffectType(enum.IntEnum):
    null = 0
    chorus = 1
    compressor = 2
    dedicated_dialog = 3
    dedicated_low_frequency = 4
    distortion = 5
    echo = 6
    equalizer = 7
    flanger = 8
    ring_modulator = 9
    reverb = 10
    eax_reverb = 11


_CHANNELS = {0: 0, 1: 1, 2: 2, 3: 4, 4: 6, 5: 6, 6: 7, 7: 8}


def channel_count(channel_format):
    """reference: Device::channel_format_to_channel_count, src/oalsfxpp.cpp:2593-2624."""
    return _CHANNELS[int(channel_format)]


class Chorus(C.Structure):
    _fields_ = [("waveform_", C.c_int), ("phase_", C.c_int), ("rate_", C.c_float),
                ("depth_", C.c_float), ("feedback_", C.c_float), ("delay_", C.c_float)]


class Compressor(C.Structure):
    _fields_ = [("on_off_", C.c_bool)]


class Dedicated(C.Structure):
    _fields_ = [("gain_", C.c_float)]


class Distortion(C.Structure):
    _fields_ = [("edge_", C.c_float), ("gain_", C.c_float), ("low_pass_cutoff_", C.c_float),
                ("eq_center_", C.c_float), ("eq_bandwidth_", C.c_float)]


class Echo(C.Structure):
    _fields_ = [("delay_", C.c_float), ("lr_delay_", C.c_float), ("damping_", C.c_float),
                ("feedback_", C.c_float), ("spread_", C.c_float)]


class Equalizer(C.Structure):
    _fields_ = [("low_cutoff_", C.c_float), ("low_gain_", C.c_float),
                ("mid1_center_", C.c_float), ("mid1_gain_", C.c_float), ("mid1_width_", C.c_float),
                ("mid2_center_", C.c_float), ("mid2_gain_", C.c_float), ("mid2_width_", C.c_float),
                ("high_cutoff_", C.c_float), ("high_gain_", C.c_float)]


class Flanger(C.Structure):
    _fields_ = Chorus._fields_


class Reverb(C.Structure):
    _fields_ = [("density_", C.c_float), ("diffusion_", C.c_float), ("gain_", C.c_float),
                ("gain_hf_", C.c_float), ("gain_lf_", C.c_float), ("decay_time_", C.c_float),
                ("decay_hf_ratio_", C.c_float), ("decay_lf_ratio_", C.c_float),
                ("reflections_gain_", C.c_float), ("reflections_delay_", C.c_float),
                ("reflections_pan_", C.c_float * 3), ("late_reverb_gain_", C.c_float),
                ("late_reverb_delay_", C.c_float), ("late_reverb_pan_", C.c_float * 3),
                ("echo_time_", C.c_float), ("echo_depth_", C.c_float),
                ("modulation_time_", C.c_float), ("modulation_depth_", C.c_float),
                ("air_absorption_gain_hf_", C.c_float), ("hf_reference_", C.c_float),
                ("lf_reference_", C.c_float), ("room_rolloff_factor_", C.c_float),
                ("decay_hf_limit_", C.c_bool)]


class RingModulator(C.Structure):
    _fields_ = [("frequency_", C.c_float), ("high_pass_cutoff_", C.c_float), ("waveform_", C.c_int)]


class EffectProps(C.Union):
    _fields_ = [("chorus_", Chorus), ("compressor_", Compressor), ("dedicated_", Dedicated),
                ("distortion_", Distortion), ("echo_", Echo), ("equalizer_", Equalizer),
                ("flanger_", Flanger), ("reverb_", Reverb), ("ring_modulator_", RingModulator)]

    def copy(self):
        out = EffectProps()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(EffectProps))
        return out

    def block(self, effect_type):
        """The member of the union that belongs to ``effect_type``."""
        return getattr(self, _BLOCK[EffectType(int(effect_type))])


_BLOCK = {
    EffectType.chorus: "chorus_", EffectType.compressor: "compressor_",
    EffectType.dedicated_dialog: "dedicated_", EffectType.dedicated_low_frequency: "dedicated_",
    EffectType.distortion: "distortion_", EffectType.echo: "echo_",
    EffectType.equalizer: "equalizer_", EffectType.flanger: "flanger_",
    EffectType.reverb: "reverb_", EffectType.eax_reverb: "reverb_",
    EffectType.ring_modulator: "ring_modulator_", EffectType.null: "chorus_",
}

assert C.sizeof(EffectProps) == 108, C.sizeof(EffectProps)


def _lib(lib):
    if lib is not None:
        return lib
    from .engine import load_library
    return load_library()


def default_props(effect_type, lib=None, **overrides):
    """Effect::set_type_and_defaults (reference: src/oalsfxpp.cpp:1782-1788) plus field overrides."""
    props = EffectProps()
    rc = _lib(lib).oalsfx_effect_defaults(int(effect_type), C.byref(props), C.sizeof(props))
    if rc != 0:
        raise ValueError(f"bad effect type {effect_type}")
    block = props.block(effect_type)
    for key, value in overrides.items():
        if key.endswith("pan_"):
            for i in range(3):
                getattr(block, key)[i] = value[i]
        else:
            setattr(block, key, value)
    return props


def normalize_props(effect_type, props, lib=None):
    out = props.copy()
    rc = _lib(lib).oalsfx_effect_normalize(int(effect_type), C.byref(out), C.sizeof(out))
    if rc != 0:
        raise ValueError(f"bad effect type {effect_type}")
    return out


def reverb_preset(group, name, lib=None, **overrides):
    """ReverbPresets::<group>::<name> (reference: src/oalsfxpp.h:583-757)."""
    props = EffectProps()
    rc = _lib(lib).oalsfx_reverb_preset(group.encode(), name.encode(), C.byref(props), C.sizeof(props))
    if rc != 0:
        raise KeyError(f"{group}::{name}")
    for key, value in overrides.items():
        setattr(props.reverb_, key, value)
    return props


def reverb_preset_names(lib=None):
    lib = _lib(lib)
    names, i = [], 0
    while True:
        s = lib.oalsfx_reverb_preset_name(i)
        if not s:
            return names
        names.append(tuple(s.decode().split("::")))
        i += 1
