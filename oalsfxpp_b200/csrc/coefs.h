// coefs.h -- flat POD "coefficient blocks" shipped from the host derivation (derive.cpp) to the
// CUDA kernels (kernels.cuh) as kernel arguments.  Everything transcendental is evaluated on the
// host with the same libm as the CPU reference, so coefficients are bit-identical; the device only
// does + - * / with FMA contraction disabled.
//
// One block per effect type; what each field means is documented against the reference's
// `do_update` that produces the same quantity (reference: src/oalsfxpp.cpp).
#ifndef OALSFX_COEFS_H
#define OALSFX_COEFS_H

#include <cstdint>

namespace oalsfx {

constexpr int kMaxChannels = 8;      // reference: oalsfxpp.cpp:44
constexpr int kMaxSlots = 4;         // reference: oalsfxpp.cpp:47
constexpr int kWetChannels = 4;      // reference: oalsfxpp.cpp:49 (first-order ambisonic wet bus)
constexpr int kMaxBlockFrames = 2048; // reference: oalsfxpp.cpp:68
constexpr float kSilenceGain = 0.00001F; // reference: oalsfxpp.cpp:56
constexpr int kLanes = 32;           // streams per tile (one warp lane each)

// Same numbering as oalsfxpp::EffectType (include/oalsfxpp.h).
enum FxType : int32_t {
	kFxNull = 0, kFxChorus, kFxCompressor, kFxDedicatedDialog, kFxDedicatedLfe, kFxDistortion,
	kFxEcho, kFxEqualizer, kFxFlanger, kFxRingModulator, kFxReverb, kFxEaxReverb, kFxTypeCount
};

// Direct-form-I biquad, a0 pre-divided (reference: FilterState, oalsfxpp.cpp:828-982).
struct Biquad { float b0, b1, b2, a1, a2; };

// One send (direct or aux) of the source (reference: Source::Send, oalsfxpp.cpp:1095-1125;
// derived in calc_panning_and_filters, oalsfxpp.cpp:3172-3346).
struct SendCoef {
	int32_t filter_type;               // ActiveFilters bit mask: 1 = high-shelf ("low_pass_"), 2 = low-shelf
	Biquad lp, hp;
	float gains[kMaxChannels][kMaxChannels]; // [input channel][output]; aux sends use outputs 0..3
};

// Chorus and flanger share one algorithm (reference: oalsfxpp.cpp:4042-4111 / 5314-5382).
struct ModDelayCoef {
	int32_t waveform;                  // 0 sinusoid, 1 triangle
	int32_t delay;                     // samples
	int32_t lfo_range, lfo_disp;
	int32_t mask;                      // ring length - 1
	float depth, feedback, lfo_scale;
	float gains[2][kMaxChannels];      // left / right side panning
	const int32_t* sin_delays;         // device table [lfo_range] of host-evaluated sinusoid delays, or null
};

struct CompressorCoef {               // reference: oalsfxpp.cpp:4319-4350
	int32_t enabled;
	float attack_rate, release_rate;
	float gains[kWetChannels][kMaxChannels];
};

struct DedicatedCoef { float gains[kMaxChannels]; }; // reference: oalsfxpp.cpp:4509-4554

struct DistortionCoef {               // reference: oalsfxpp.cpp:4627-4673
	Biquad low_pass, band_pass;
	float edge_coeff;
	float gains[kMaxChannels];         // ambient gain * attenuation (oalsfxpp.cpp:4736)
};

struct EchoCoef {                     // reference: oalsfxpp.cpp:4835-4885
	int32_t tap1, tap2, mask;
	Biquad filter;
	float feed_gain;
	float gains[2][kMaxChannels];
};

struct EqualizerCoef {                // reference: oalsfxpp.cpp:5076-5159
	Biquad band[4];
	float gains[kWetChannels][kMaxChannels];
};

struct RingModCoef {                  // reference: oalsfxpp.cpp:5598-5650
	int32_t waveform;                  // 0 sin, 1 saw, 2 square
	int32_t step;
	Biquad filter;                     // b0=a, b1=-a, b2=0, a1=-a, a2=0
	float gains[kWetChannels][kMaxChannels];
};

// Reverb / EAX reverb (reference: do_update_device oalsfxpp.cpp:5928-5950, do_update :5952-6076,
// update_* :7014-7187, update_3d_panning :7306-7350).  "tap1"-style arrays are the NEW ([..][1])
// tap sets; the OLD ([..][0]) sets are per-stream device state because the reference commits them
// only when a 128-sample cross-fade completes (oalsfxpp.cpp:6118-6138).
struct ReverbCoef {
	int32_t is_eax;
	Biquad lp, hp;                     // master input shelves (hp used by EAX only)
	int32_t early_tap[4];
	float early_tap_coeff[4];
	int32_t late_feed_tap;
	int32_t late_tap[4];
	float ap_feed_coeff, mix_x, mix_y;
	int32_t early_ap_off[4];
	int32_t early_off[4];
	float early_coeff[4];
	int32_t mod_range;
	float mod_depth, mod_coeff;
	float density_gain;
	int32_t late_off[4];
	int32_t late_ap_off[4];
	float t60_lf[4][3], t60_hf[4][3], t60_mid[4];
	float pan_early[4][kMaxChannels], pan_late[4][kMaxChannels];
	// Ring geometry (reference: alloc_lines, oalsfxpp.cpp:6556-6598): masks and word offsets of the
	// five 4-line rings inside the slot's per-lane ring region; line j of ring r starts at
	// ring_base[r] + j * (mask[r] + 1).
	int32_t mask[5];                   // main, early all-pass, early line, late all-pass, late line
	int32_t ring_base[5];
	const float* mod_sinus;            // device table [mod_range] of host-evaluated sin(tau*i/range), or null
};

// A shelf / peaking / pass filter whose reference frequency is >= rate/2 is unstable in the reference
// itself (outputs run to Inf/NaN, e.g. the echo's 5 kHz shelf at 8 kHz).  Such slots are kept off the
// "fast" kernels, whose unconditional `sample * 0` for skipped gains would spread a NaN to outputs
// the reference leaves untouched.
constexpr uint32_t kCoefUnstable = 1U;

struct SlotCoef {
	int32_t type;                      // FxType
	uint32_t flags;                    // kCoefUnstable: a biquad was designed at or above Nyquist
	union {
		ModDelayCoef mod_delay;        // chorus, flanger
		CompressorCoef compressor;
		DedicatedCoef dedicated;
		DistortionCoef distortion;
		EchoCoef echo;
		EqualizerCoef equalizer;
		RingModCoef ring_mod;
		ReverbCoef reverb;
	} u;
};

// Per-lane ring words a slot of the given type needs at `rate` Hz (0 for ring-less effects).
int ring_words_for(int fx_type, int sampling_rate);

} // namespace oalsfx

#endif
