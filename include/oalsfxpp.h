// oalsfxpp.h -- drop-in public C++ API of the B200-native effects engine.
//
// Source-compatible re-statement of the reference's public surface
// (reference: src/oalsfxpp.h:36-62 enums, :65-530 EffectProps, :532-581 Effect/SendProps,
// :583-757 ReverbPresets, :760-922 Api).  Same namespace, type names, member names, member order
// and constants, so code written against the reference header (e.g. src/oalsfxpp_test.cpp:774-891)
// compiles and links unchanged against liboalsfx_b200.so.  Layout facts kept:
// sizeof(EffectProps)=108, sizeof(Effect)=112, sizeof(SendProps)=12, all trivially copyable.
//
// Behind Api sits the batched CUDA engine (include/oalsfx_engine.h); one Api == one engine
// stream.  There is no CPU implementation: initialize() fails if no CUDA device is usable.
#ifndef OALSFXPP_INCLUDED
#define OALSFXPP_INCLUDED

#include <array>
#include <memory>

namespace oalsfxpp {

enum class ChannelFormat { none, mono, stereo, quad, five_point_one, five_point_one_rear, six_point_one, seven_point_one };

enum class EffectType {
	null, chorus, compressor, dedicated_dialog, dedicated_low_frequency, distortion,
	echo, equalizer, flanger, ring_modulator, reverb, eax_reverb
};

// min_/max_/default_ triple of one property (reference declares each constant separately).
#define OALSFX_RANGE(T, name, lo, hi, def) \
	static constexpr T min_##name = lo; static constexpr T max_##name = hi; static constexpr T default_##name = def;
// Every property block has the same three helpers (reference: src/oalsfxpp.cpp:1409-1725).
#define OALSFX_PROP_METHODS(S) \
	void set_defaults(); void normalize(); static bool are_equal(const S& a, const S& b);

union EffectProps {
	using Pan = std::array<float, 3>;

	struct Null {};

	struct Chorus {
		static constexpr int waveform_sinusoid = 0;
		static constexpr int waveform_triangle = 1;
		OALSFX_RANGE(int, waveform, waveform_sinusoid, waveform_triangle, waveform_triangle)
		OALSFX_RANGE(int, phase, -180, 180, 90)
		OALSFX_RANGE(float, rate, 0.0F, 10.0F, 1.1F)
		OALSFX_RANGE(float, depth, 0.0F, 1.0F, 0.1F)
		OALSFX_RANGE(float, feedback, -1.0F, 1.0F, 0.25F)
		OALSFX_RANGE(float, delay, 0.0F, 0.016F, 0.016F)
		int waveform_; int phase_; float rate_; float depth_; float feedback_; float delay_;
		OALSFX_PROP_METHODS(Chorus)
	};

	struct Compressor {
		OALSFX_RANGE(bool, on_off, false, true, true)
		bool on_off_;
		OALSFX_PROP_METHODS(Compressor)
	};

	struct Dedicated {
		OALSFX_RANGE(float, gain, 0.0F, 1.0F, 1.0F)
		float gain_;
		OALSFX_PROP_METHODS(Dedicated)
	};

	struct Distortion {
		OALSFX_RANGE(float, edge, 0.0F, 1.0F, 0.2F)
		OALSFX_RANGE(float, gain, 0.01F, 1.0F, 0.05F)
		OALSFX_RANGE(float, low_pass_cutoff, 80.0F, 24000.0F, 8000.0F)
		OALSFX_RANGE(float, eq_center, 80.0F, 24000.0F, 3600.0F)
		OALSFX_RANGE(float, eq_bandwidth, 80.0F, 24000.0F, 3600.0F)
		float edge_; float gain_; float low_pass_cutoff_; float eq_center_; float eq_bandwidth_;
		OALSFX_PROP_METHODS(Distortion)
	};

	struct Echo {
		OALSFX_RANGE(float, delay, 0.0F, 0.207F, 0.1F)
		OALSFX_RANGE(float, lr_delay, 0.0F, 0.404F, 0.1F)
		OALSFX_RANGE(float, damping, 0.0F, 0.99F, 0.5F)
		OALSFX_RANGE(float, feedback, 0.0F, 1.0F, 0.5F)
		OALSFX_RANGE(float, spread, -1.0F, 1.0F, -1.0F)
		float delay_; float lr_delay_; float damping_; float feedback_; float spread_;
		OALSFX_PROP_METHODS(Echo)
	};

	struct Equalizer {
		OALSFX_RANGE(float, low_gain, 0.126F, 7.943F, 1.0F)
		OALSFX_RANGE(float, low_cutoff, 50.0F, 800.0F, 200.0F)
		OALSFX_RANGE(float, mid1_gain, 0.126F, 7.943F, 1.0F)
		OALSFX_RANGE(float, mid1_center, 200.0F, 3000.0F, 500.0F)
		OALSFX_RANGE(float, mid1_width, 0.01F, 1.0F, 1.0F)
		OALSFX_RANGE(float, mid2_gain, 0.126F, 7.943F, 1.0F)
		OALSFX_RANGE(float, mid2_center, 1000.0F, 8000.0F, 3000.0F)
		OALSFX_RANGE(float, mid2_width, 0.01F, 1.0F, 1.0F)
		OALSFX_RANGE(float, high_gain, 0.126F, 7.943F, 1.0F)
		OALSFX_RANGE(float, high_cutoff, 4000.0F, 16000.0F, 6000.0F)
		float low_cutoff_; float low_gain_;
		float mid1_center_; float mid1_gain_; float mid1_width_;
		float mid2_center_; float mid2_gain_; float mid2_width_;
		float high_cutoff_; float high_gain_;
		OALSFX_PROP_METHODS(Equalizer)
	};

	struct Flanger {
		static constexpr int waveform_sinusoid = 0;
		static constexpr int waveform_triangle = 1;
		OALSFX_RANGE(int, waveform, waveform_sinusoid, waveform_triangle, waveform_triangle)
		OALSFX_RANGE(int, phase, -180, 180, 0)
		OALSFX_RANGE(float, rate, 0.0F, 10.0F, 0.27F)
		OALSFX_RANGE(float, depth, 0.0F, 1.0F, 1.0F)
		OALSFX_RANGE(float, feedback, -1.0F, 1.0F, -0.5F)
		OALSFX_RANGE(float, delay, 0.0F, 0.004F, 0.002F)
		int waveform_; int phase_; float rate_; float depth_; float feedback_; float delay_;
		OALSFX_PROP_METHODS(Flanger)
	};

	struct Reverb {
		OALSFX_RANGE(float, density, 0.0F, 1.0F, 1.0F)
		OALSFX_RANGE(float, diffusion, 0.0F, 1.0F, 1.0F)
		OALSFX_RANGE(float, gain, 0.0F, 1.0F, 0.32F)
		OALSFX_RANGE(float, gain_hf, 0.0F, 1.0F, 0.89F)
		OALSFX_RANGE(float, gain_lf, 0.0F, 1.0F, 1.0F)
		OALSFX_RANGE(float, decay_time, 0.1F, 20.0F, 1.49F)
		OALSFX_RANGE(float, decay_hf_ratio, 0.1F, 2.0F, 0.83F)
		OALSFX_RANGE(float, decay_lf_ratio, 0.1F, 2.0F, 1.0F)
		OALSFX_RANGE(float, reflections_gain, 0.0F, 3.16F, 0.05F)
		OALSFX_RANGE(float, reflections_delay, 0.0F, 0.3F, 0.007F)
		OALSFX_RANGE(float, reflections_pan_xyz, -1.0F, 1.0F, 0.0F)
		OALSFX_RANGE(float, late_reverb_gain, 0.0F, 10.0F, 1.26F)
		OALSFX_RANGE(float, late_reverb_delay, 0.0F, 0.1F, 0.011F)
		OALSFX_RANGE(float, late_reverb_pan_xyz, -1.0F, 1.0F, 0.0F)
		OALSFX_RANGE(float, echo_time, 0.075F, 0.25F, 0.25F)
		OALSFX_RANGE(float, echo_depth, 0.0F, 1.0F, 0.0F)
		OALSFX_RANGE(float, modulation_time, 0.04F, 4.0F, 0.25F)
		OALSFX_RANGE(float, modulation_depth, 0.0F, 1.0F, 0.0F)
		OALSFX_RANGE(float, air_absorption_gain_hf, 0.892F, 1.0F, 0.994F)
		OALSFX_RANGE(float, hf_reference, 1000.0F, 20000.0F, 5000.0F)
		OALSFX_RANGE(float, lf_reference, 20.0F, 1000.0F, 250.0F)
		OALSFX_RANGE(float, room_rolloff_factor, 0.0F, 10.0F, 0.0F)
		OALSFX_RANGE(bool, decay_hf_limit, false, true, true)
		float density_; float diffusion_; float gain_; float gain_hf_;
		float gain_lf_;                 // EAX
		float decay_time_; float decay_hf_ratio_;
		float decay_lf_ratio_;          // EAX
		float reflections_gain_; float reflections_delay_;
		Pan reflections_pan_;           // EAX
		float late_reverb_gain_; float late_reverb_delay_;
		Pan late_reverb_pan_;           // EAX
		float echo_time_; float echo_depth_; float modulation_time_; float modulation_depth_; // EAX
		float air_absorption_gain_hf_;
		float hf_reference_; float lf_reference_; // EAX
		float room_rolloff_factor_;
		bool decay_hf_limit_;
		OALSFX_PROP_METHODS(Reverb)
	};

	struct RingModulator {
		static constexpr int waveform_sinusoid = 0;
		static constexpr int waveform_sawtooth = 1;
		static constexpr int waveform_square = 2;
		OALSFX_RANGE(float, frequency, 0.0F, 8000.0F, 440.0F)
		OALSFX_RANGE(float, high_pass_cutoff, 0.0F, 24000.0F, 800.0F)
		OALSFX_RANGE(int, waveform, waveform_sinusoid, waveform_square, waveform_sinusoid)
		float frequency_; float high_pass_cutoff_; int waveform_;
		OALSFX_PROP_METHODS(RingModulator)
	};

	Chorus chorus_;
	Compressor compressor_;
	Dedicated dedicated_;
	Distortion distortion_;
	Echo echo_;
	Equalizer equalizer_;
	Flanger flanger_;
	Reverb reverb_;
	RingModulator ring_modulator_;
}; // EffectProps

struct Effect {
	EffectType type_;
	EffectProps props_;

	void set_defaults();
	void set_type_and_defaults(const EffectType effect_type);
	void normalize();
	static bool are_equal(const Effect& a, const Effect& b);
};

struct SendProps {
	static constexpr float lp_frequency_reference = 5000.0F;
	static constexpr float hp_frequency_reference = 250.0F;
	OALSFX_RANGE(float, gain, 0.0F, 1.0F, 1.0F)
	OALSFX_RANGE(float, gain_hf, 0.0F, 1.0F, 1.0F)
	OALSFX_RANGE(float, gain_lf, 0.0F, 1.0F, 1.0F)
	float gain_; float gain_hf_; float gain_lf_;
	OALSFX_PROP_METHODS(SendProps)
};

#undef OALSFX_RANGE
#undef OALSFX_PROP_METHODS

// The 113 EFX reverb presets (reference: src/oalsfxpp.h:583-757, values src/oalsfxpp.cpp:1938-2191).
// The member lists live in oalsfxpp_presets.inc, generated by oracle/gen_presets.py.
struct ReverbPresets {
#define OALSFX_PRESET_GROUP_BEGIN(G) struct G {
#define OALSFX_PRESET(G, N, ...) static const EffectProps::Reverb N;
#define OALSFX_PRESET_GROUP_END(G) };
#include "oalsfxpp_presets.inc"
#undef OALSFX_PRESET_GROUP_BEGIN
#undef OALSFX_PRESET
#undef OALSFX_PRESET_GROUP_END
};

class Api {
public:
	Api();
	Api(const Api&) = delete;
	Api& operator=(const Api&) = delete;
	~Api();

	// All bool-returning calls: true = success.  Getters return 0 / ChannelFormat::none on error.
	bool initialize(const ChannelFormat channel_format, const int sampling_rate, const int effect_count);
	bool is_initialized() const;
	int get_sampling_rate() const;
	ChannelFormat get_channel_format() const;
	int get_channel_count() const;
	int get_effect_count() const;

	bool get_effect(const int effect_index, Effect& effect) const;           // active
	bool get_deferred_effect(const int effect_index, Effect& effect) const;  // staged
	bool set_effect_type(const int effect_index, const EffectType effect_type);
	bool set_effect_props(const int effect_index, const EffectProps& effect_props);
	bool set_effect(const int effect_index, const Effect& effect);

	// effect_index < 0 addresses the direct (dry) send.
	bool get_send_props(const int effect_index, SendProps& send_props) const;
	bool get_deferred_send_props(const int effect_index, SendProps& send_props) const;
	bool set_send_props(const int effect_index, const SendProps& send_props);

	bool apply_changes();

	// Interleaved fp32 in -> interleaved fp32 out (overwritten, not clipped); host pointers.
	bool mix(const int sample_count, const float* src_samples, float* dst_samples);

	void uninitialize();
	const char* get_error_message() const;

	static int get_min_channels();
	static int get_max_channels();
	static int get_min_sampling_rate();
	static int get_max_sampling_rate();
	static int get_min_effects();
	static int get_max_effects();
	static ChannelFormat channel_count_to_channel_format(const int channel_count);
	static int channel_format_to_channel_count(const ChannelFormat channel_format);

private:
	class Impl;
	using ApiImplUPtr = std::unique_ptr<Impl>;
	ApiImplUPtr pimpl_;
	mutable const char* error_message_;
};

} // namespace oalsfxpp

#endif // OALSFXPP_INCLUDED
