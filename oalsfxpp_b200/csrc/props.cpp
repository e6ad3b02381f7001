// props.cpp -- the parameter model behind include/oalsfxpp.h: set_defaults / normalize / are_equal
// for every property block, Effect dispatch, SendProps and the reverb preset constants
// (reference: src/oalsfxpp.cpp:1409-1930 and :1938-2191).  Host-only, no arithmetic beyond clamps.
//
// Compiled as C++14 so the out-of-class definitions of the static constexpr members below are
// emitted for C++14 clients that odr-use them (the reference defines them at oalsfxpp.cpp:1158-1407).
#include "oalsfxpp.h"

#include <algorithm>

namespace oalsfxpp {
namespace {

template <typename T>
void clamp_to(T& v, const T lo, const T hi) { v = std::min(hi, std::max(lo, v)); }

} // namespace

// One row per property: FIELD(member, range-name).  The three helpers are generated from the rows,
// so the field list of each block is written once.
#define CHORUS_LIKE_FIELDS(FIELD) \
	FIELD(waveform_, waveform) FIELD(phase_, phase) FIELD(rate_, rate) FIELD(depth_, depth) \
	FIELD(feedback_, feedback) FIELD(delay_, delay)
#define COMPRESSOR_FIELDS(FIELD) FIELD(on_off_, on_off)
#define DEDICATED_FIELDS(FIELD) FIELD(gain_, gain)
#define DISTORTION_FIELDS(FIELD) \
	FIELD(edge_, edge) FIELD(gain_, gain) FIELD(low_pass_cutoff_, low_pass_cutoff) \
	FIELD(eq_center_, eq_center) FIELD(eq_bandwidth_, eq_bandwidth)
#define ECHO_FIELDS(FIELD) \
	FIELD(delay_, delay) FIELD(lr_delay_, lr_delay) FIELD(damping_, damping) FIELD(feedback_, feedback) \
	FIELD(spread_, spread)
#define EQUALIZER_FIELDS(FIELD) \
	FIELD(low_cutoff_, low_cutoff) FIELD(low_gain_, low_gain) FIELD(mid1_center_, mid1_center) \
	FIELD(mid1_gain_, mid1_gain) FIELD(mid1_width_, mid1_width) FIELD(mid2_center_, mid2_center) \
	FIELD(mid2_gain_, mid2_gain) FIELD(mid2_width_, mid2_width) FIELD(high_cutoff_, high_cutoff) \
	FIELD(high_gain_, high_gain)
#define RINGMOD_FIELDS(FIELD) \
	FIELD(frequency_, frequency) FIELD(high_pass_cutoff_, high_pass_cutoff) FIELD(waveform_, waveform)
#define SEND_FIELDS(FIELD) FIELD(gain_, gain) FIELD(gain_hf_, gain_hf) FIELD(gain_lf_, gain_lf)
// Reverb's scalar fields; the two pan vectors are handled by hand below.
#define REVERB_SCALAR_FIELDS(FIELD) \
	FIELD(density_, density) FIELD(diffusion_, diffusion) FIELD(gain_, gain) FIELD(gain_hf_, gain_hf) \
	FIELD(gain_lf_, gain_lf) FIELD(decay_time_, decay_time) FIELD(decay_hf_ratio_, decay_hf_ratio) \
	FIELD(decay_lf_ratio_, decay_lf_ratio) FIELD(reflections_gain_, reflections_gain) \
	FIELD(reflections_delay_, reflections_delay) FIELD(late_reverb_gain_, late_reverb_gain) \
	FIELD(late_reverb_delay_, late_reverb_delay) FIELD(echo_time_, echo_time) FIELD(echo_depth_, echo_depth) \
	FIELD(modulation_time_, modulation_time) FIELD(modulation_depth_, modulation_depth) \
	FIELD(air_absorption_gain_hf_, air_absorption_gain_hf) FIELD(hf_reference_, hf_reference) \
	FIELD(lf_reference_, lf_reference) FIELD(room_rolloff_factor_, room_rolloff_factor) \
	FIELD(decay_hf_limit_, decay_hf_limit)

#define F_DEFAULT(m, r) m = default_##r;
#define F_CLAMP(m, r) clamp_to(m, min_##r, max_##r);
#define F_EQUAL(m, r) && a.m == b.m
#define DEFINE_HELPERS(S, FIELDS) \
	void S::set_defaults() { FIELDS(F_DEFAULT) } \
	void S::normalize() { FIELDS(F_CLAMP) } \
	bool S::are_equal(const S& a, const S& b) { return true FIELDS(F_EQUAL); }

DEFINE_HELPERS(EffectProps::Chorus, CHORUS_LIKE_FIELDS)
DEFINE_HELPERS(EffectProps::Flanger, CHORUS_LIKE_FIELDS)
DEFINE_HELPERS(EffectProps::Dedicated, DEDICATED_FIELDS)
DEFINE_HELPERS(EffectProps::Distortion, DISTORTION_FIELDS)
DEFINE_HELPERS(EffectProps::Echo, ECHO_FIELDS)
DEFINE_HELPERS(EffectProps::RingModulator, RINGMOD_FIELDS)
DEFINE_HELPERS(SendProps, SEND_FIELDS)

// The compressor's normalize() is empty in the reference (oalsfxpp.cpp:1451-1453).
void EffectProps::Compressor::set_defaults() { COMPRESSOR_FIELDS(F_DEFAULT) }
void EffectProps::Compressor::normalize() {}
bool EffectProps::Compressor::are_equal(const Compressor& a, const Compressor& b) { return true COMPRESSOR_FIELDS(F_EQUAL); }

// Equalizer::set_defaults assigns default_high_gain to low_gain_ (oalsfxpp.cpp:1540); both are 1.0.
void EffectProps::Equalizer::set_defaults() { EQUALIZER_FIELDS(F_DEFAULT) low_gain_ = default_high_gain; }
void EffectProps::Equalizer::normalize() { EQUALIZER_FIELDS(F_CLAMP) }
bool EffectProps::Equalizer::are_equal(const Equalizer& a, const Equalizer& b) { return true EQUALIZER_FIELDS(F_EQUAL); }

void EffectProps::Reverb::set_defaults()
{
	REVERB_SCALAR_FIELDS(F_DEFAULT)
	reflections_pan_.fill(default_reflections_pan_xyz);
	late_reverb_pan_.fill(default_late_reverb_pan_xyz);
}

void EffectProps::Reverb::normalize()
{
	REVERB_SCALAR_FIELDS(F_CLAMP)
	for (float& v : reflections_pan_) {
		clamp_to(v, min_reflections_pan_xyz, max_reflections_pan_xyz);
	}
	for (float& v : late_reverb_pan_) {
		clamp_to(v, min_late_reverb_pan_xyz, max_late_reverb_pan_xyz);
	}
}

bool EffectProps::Reverb::are_equal(const Reverb& a, const Reverb& b)
{
	return true REVERB_SCALAR_FIELDS(F_EQUAL) && a.reflections_pan_ == b.reflections_pan_ &&
		a.late_reverb_pan_ == b.late_reverb_pan_;
}

// ---- Effect (oalsfxpp.cpp:1727-1893) ----------------------------------------------------------
// Calls `op` on the property block that belongs to the effect type.
#define DISPATCH_PROPS(type, props, CALL) \
	switch (type) { \
	case EffectType::chorus: CALL(props.chorus_, EffectProps::Chorus) break; \
	case EffectType::compressor: CALL(props.compressor_, EffectProps::Compressor) break; \
	case EffectType::dedicated_dialog: \
	case EffectType::dedicated_low_frequency: CALL(props.dedicated_, EffectProps::Dedicated) break; \
	case EffectType::distortion: CALL(props.distortion_, EffectProps::Distortion) break; \
	case EffectType::echo: CALL(props.echo_, EffectProps::Echo) break; \
	case EffectType::equalizer: CALL(props.equalizer_, EffectProps::Equalizer) break; \
	case EffectType::flanger: CALL(props.flanger_, EffectProps::Flanger) break; \
	case EffectType::eax_reverb: \
	case EffectType::reverb: CALL(props.reverb_, EffectProps::Reverb) break; \
	case EffectType::ring_modulator: CALL(props.ring_modulator_, EffectProps::RingModulator) break; \
	case EffectType::null: \
	default: break; \
	}

void Effect::set_defaults()
{
#define CALL(block, S) block.set_defaults();
	DISPATCH_PROPS(type_, props_, CALL)
#undef CALL
}

void Effect::set_type_and_defaults(const EffectType effect_type)
{
	type_ = effect_type;
	set_defaults();
}

void Effect::normalize()
{
#define CALL(block, S) block.normalize();
	DISPATCH_PROPS(type_, props_, CALL)
#undef CALL
}

bool Effect::are_equal(const Effect& a, const Effect& b)
{
	if (a.type_ != b.type_) {
		return false;
	}
	if (a.type_ == EffectType::null) {
		return true;
	}
	bool equal = false; // unknown type values compare unequal (oalsfxpp.cpp:1890-1891)
	const EffectProps& pb = b.props_;
	switch (a.type_) {
	case EffectType::chorus: equal = EffectProps::Chorus::are_equal(a.props_.chorus_, pb.chorus_); break;
	case EffectType::compressor: equal = EffectProps::Compressor::are_equal(a.props_.compressor_, pb.compressor_); break;
	case EffectType::dedicated_dialog:
	case EffectType::dedicated_low_frequency: equal = EffectProps::Dedicated::are_equal(a.props_.dedicated_, pb.dedicated_); break;
	case EffectType::distortion: equal = EffectProps::Distortion::are_equal(a.props_.distortion_, pb.distortion_); break;
	case EffectType::echo: equal = EffectProps::Echo::are_equal(a.props_.echo_, pb.echo_); break;
	case EffectType::equalizer: equal = EffectProps::Equalizer::are_equal(a.props_.equalizer_, pb.equalizer_); break;
	case EffectType::flanger: equal = EffectProps::Flanger::are_equal(a.props_.flanger_, pb.flanger_); break;
	case EffectType::eax_reverb:
	case EffectType::reverb: equal = EffectProps::Reverb::are_equal(a.props_.reverb_, pb.reverb_); break;
	case EffectType::ring_modulator: equal = EffectProps::RingModulator::are_equal(a.props_.ring_modulator_, pb.ring_modulator_); break;
	default: break;
	}
	return equal;
}

// ---- preset constants ---------------------------------------------------------------------------
#define OALSFX_PRESET_GROUP_BEGIN(G)
#define OALSFX_PRESET(G, N, ...) const EffectProps::Reverb ReverbPresets::G::N = {__VA_ARGS__};
#define OALSFX_PRESET_GROUP_END(G)
#include "oalsfxpp_presets.inc"
#undef OALSFX_PRESET_GROUP_BEGIN
#undef OALSFX_PRESET
#undef OALSFX_PRESET_GROUP_END

// ---- C++14 definitions of the static constexpr members (harmless redeclarations in C++17) ----------
#if __cplusplus < 201703L
#define DEF3(S, T, r) constexpr T S::min_##r; constexpr T S::max_##r; constexpr T S::default_##r;
#define DEF_WAVE2(S) constexpr int S::waveform_sinusoid; constexpr int S::waveform_triangle;
#define CHORUS_LIKE_DEFS(S) DEF_WAVE2(S) DEF3(S, int, waveform) DEF3(S, int, phase) DEF3(S, float, rate) \
	DEF3(S, float, depth) DEF3(S, float, feedback) DEF3(S, float, delay)
CHORUS_LIKE_DEFS(EffectProps::Chorus)
CHORUS_LIKE_DEFS(EffectProps::Flanger)
DEF3(EffectProps::Compressor, bool, on_off)
DEF3(EffectProps::Dedicated, float, gain)
#define FLOAT_DEF(m, r) DEF3(CUR, float, r)
#define CUR EffectProps::Distortion
DISTORTION_FIELDS(FLOAT_DEF)
#undef CUR
#define CUR EffectProps::Echo
ECHO_FIELDS(FLOAT_DEF)
#undef CUR
#define CUR EffectProps::Equalizer
EQUALIZER_FIELDS(FLOAT_DEF)
#undef CUR
#define CUR SendProps
SEND_FIELDS(FLOAT_DEF)
#undef CUR
constexpr float SendProps::lp_frequency_reference;
constexpr float SendProps::hp_frequency_reference;
DEF3(EffectProps::RingModulator, float, frequency)
DEF3(EffectProps::RingModulator, float, high_pass_cutoff)
DEF3(EffectProps::RingModulator, int, waveform)
constexpr int EffectProps::RingModulator::waveform_sinusoid;
constexpr int EffectProps::RingModulator::waveform_sawtooth;
constexpr int EffectProps::RingModulator::waveform_square;
#define R3(r) DEF3(EffectProps::Reverb, float, r)
R3(density) R3(diffusion) R3(gain) R3(gain_hf) R3(gain_lf) R3(decay_time) R3(decay_hf_ratio) R3(decay_lf_ratio)
R3(reflections_gain) R3(reflections_delay) R3(reflections_pan_xyz) R3(late_reverb_gain) R3(late_reverb_delay)
R3(late_reverb_pan_xyz) R3(echo_time) R3(echo_depth) R3(modulation_time) R3(modulation_depth)
R3(air_absorption_gain_hf) R3(hf_reference) R3(lf_reference) R3(room_rolloff_factor)
DEF3(EffectProps::Reverb, bool, decay_hf_limit)
#endif

} // namespace oalsfxpp
