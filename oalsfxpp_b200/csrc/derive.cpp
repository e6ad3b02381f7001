// derive.cpp -- host-side parameter -> coefficient derivation.  See derive.h.
//
// Every formula below restates what the reference evaluates in its `do_update*` routines, with the
// same fp32 expression order and the same libm entry points, so the resulting coefficient blocks
// are bit-identical to the reference's private state.  Citations are to
// /root/reference/src/oalsfxpp.cpp.  Must be compiled WITHOUT FMA contraction / fast-math.
#include "derive.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace oalsfx {
namespace {

constexpr float kPi = 3.14159265358979323846F;
constexpr float kPi2 = 1.57079632679489661923F;
constexpr float kTau = 6.28318530717958647692F;

inline float clampf(float v, float lo, float hi) { return std::min(hi, std::max(lo, v)); }
inline float lerpf(float a, float b, float mu) { return a + ((b - a) * mu); }

int next_pow2(int v)
{
	if (v > 0) {
		v -= 1;
		v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
	}
	return v + 1;
}

// ---- output decoders ------------------------------------------------------------------------
// Speaker ids in the order of the reference's ChannelId enum (oalsfxpp.cpp:71-83).
enum Spk { kNone, kFL, kFR, kFC, kLFE, kBL, kBR, kBC, kSL, kSR };

struct DecoderRow { Spk spk; float c[16]; };

// Ambisonic decoder rows (ACN/N3D), one table per layout (oalsfxpp.cpp:295-477).  Only the
// non-zero leading coefficients are listed; the rest are zero.
const DecoderRow kMonoDec[] = {{kFC, {1.0F}}};
const DecoderRow kStereoDec[] = {
	{kFL, {5.00000000E-1F, 2.88675135E-1F, 0.0F, 1.19573156E-1F}},
	{kFR, {5.00000000E-1F, -2.88675135E-1F, 0.0F, 1.19573156E-1F}},
};
const DecoderRow kQuadDec[] = {
	{kBL, {3.53553391E-1F, 2.04124145E-1F, 0.0F, -2.04124145E-1F}},
	{kFL, {3.53553391E-1F, 2.04124145E-1F, 0.0F, 2.04124145E-1F}},
	{kFR, {3.53553391E-1F, -2.04124145E-1F, 0.0F, 2.04124145E-1F}},
	{kBR, {3.53553391E-1F, -2.04124145E-1F, 0.0F, -2.04124145E-1F}},
};
#define X51_ROWS(L, R) \
	{L, {3.33001372E-1F, 1.89085671E-1F, 0.0F, -2.00041334E-1F, -2.12309737E-2F, 0.0F, 0.0F, 0.0F, -1.14573483E-2F}}, \
	{kFL, {1.47751298E-1F, 1.28994110E-1F, 0.0F, 1.15190495E-1F, 7.44949143E-2F, 0.0F, 0.0F, 0.0F, -6.47739980E-3F}}, \
	{kFC, {7.73595729E-2F, 0.0F, 0.0F, 9.71390298E-2F, 0.0F, 0.0F, 0.0F, 0.0F, 5.18625335E-2F}}, \
	{kFR, {1.47751298E-1F, -1.28994110E-1F, 0.0F, 1.15190495E-1F, -7.44949143E-2F, 0.0F, 0.0F, 0.0F, -6.47739980E-3F}}, \
	{R, {3.33001372E-1F, -1.89085671E-1F, 0.0F, -2.00041334E-1F, 2.12309737E-2F, 0.0F, 0.0F, 0.0F, -1.14573483E-2F}},
const DecoderRow kX51SideDec[] = {X51_ROWS(kSL, kSR)};
const DecoderRow kX51RearDec[] = {X51_ROWS(kBL, kBR)};
#undef X51_ROWS
const DecoderRow kX61Dec[] = {
	{kSL, {2.04462744E-1F, 2.17178497E-1F, 0.0F, -4.39990188E-2F, -2.60787329E-2F, 0.0F, 0.0F, 0.0F, -6.87238843E-2F}},
	{kFL, {1.18130342E-1F, 9.34633906E-2F, 0.0F, 1.08553749E-1F, 6.80658795E-2F, 0.0F, 0.0F, 0.0F, 1.08999485E-2F}},
	{kFC, {7.73595729E-2F, 0.0F, 0.0F, 9.71390298E-2F, 0.0F, 0.0F, 0.0F, 0.0F, 5.18625335E-2F}},
	{kFR, {1.18130342E-1F, -9.34633906E-2F, 0.0F, 1.08553749E-1F, -6.80658795E-2F, 0.0F, 0.0F, 0.0F, 1.08999485E-2F}},
	{kSR, {2.04462744E-1F, -2.17178497E-1F, 0.0F, -4.39990188E-2F, 2.60787329E-2F, 0.0F, 0.0F, 0.0F, -6.87238843E-2F}},
	{kBC, {2.50001688E-1F, 0.0F, 0.0F, -2.50000094E-1F, 0.0F, 0.0F, 0.0F, 0.0F, 6.05133395E-2F}},
};
// NB: the 7.1 table has no front-centre row, so FC stays silent (reference quirk, oalsfxpp.cpp:428-477).
const DecoderRow kX71Dec[] = {
	{kBL, {2.04124145E-1F, 1.08880247E-1F, 0.0F, -1.88586120E-1F, -1.29099444E-1F, 0.0F, 0.0F, 0.0F, 7.45355993E-2F, 3.73460789E-2F}},
	{kSL, {2.04124145E-1F, 2.17760495E-1F, 0.0F, 0.0F, 0.0F, 0.0F, 0.0F, 0.0F, -1.49071198E-1F, -3.73460789E-2F}},
	{kFL, {2.04124145E-1F, 1.08880247E-1F, 0.0F, 1.88586120E-1F, 1.29099444E-1F, 0.0F, 0.0F, 0.0F, 7.45355993E-2F, 3.73460789E-2F}},
	{kFR, {2.04124145E-1F, -1.08880247E-1F, 0.0F, 1.88586120E-1F, -1.29099444E-1F, 0.0F, 0.0F, 0.0F, 7.45355993E-2F, -3.73460789E-2F}},
	{kSR, {2.04124145E-1F, -2.17760495E-1F, 0.0F, 0.0F, 0.0F, 0.0F, 0.0F, 0.0F, -1.49071198E-1F, 3.73460789E-2F}},
	{kBR, {2.04124145E-1F, -1.08880247E-1F, 0.0F, -1.88586120E-1F, 1.29099444E-1F, 0.0F, 0.0F, 0.0F, 7.45355993E-2F, -3.73460789E-2F}},
};

struct LayoutDesc {
	int channels;
	Spk order[kMaxChannels];           // WFX channel order (oalsfxpp.cpp:2420-2486)
	const DecoderRow* dec; int dec_rows; int coeff_count; // oalsfxpp.cpp:2489-2553
	int map_count; float map_deg[kMaxChannels]; // source channel map (oalsfxpp.cpp:3048-3098); LFE keyed off `order`
};

constexpr float kNoAngle = 0.0F;
const LayoutDesc kLayouts[8] = {
	{0, {}, nullptr, 0, 0, 0, {}},
	{1, {kFC}, kMonoDec, 1, 1, 1, {0.0F}},
	{2, {kFL, kFR}, kStereoDec, 2, 4, 2, {-30.0F, 30.0F}},
	{4, {kFL, kFR, kBL, kBR}, kQuadDec, 4, 4, 4, {-45.0F, 45.0F, -135.0F, 135.0F}},
	{6, {kFL, kFR, kFC, kLFE, kSL, kSR}, kX51SideDec, 5, 9, 6, {-30.0F, 30.0F, 0.0F, kNoAngle, -110.0F, 110.0F}},
	// 5.1-rear has no case in the reference's calc_panning_and_filters (oalsfxpp.cpp:3190-3225):
	// no source gains are ever produced, the output is silent.  Reproduced via map_count = 0.
	{6, {kFL, kFR, kFC, kLFE, kBL, kBR}, kX51RearDec, 5, 9, 0, {}},
	{7, {kFL, kFR, kFC, kLFE, kBC, kSL, kSR}, kX61Dec, 6, 9, 7, {-30.0F, 30.0F, 0.0F, kNoAngle, 180.0F, -90.0F, 90.0F}},
	{8, {kFL, kFR, kFC, kLFE, kBL, kBR, kSL, kSR}, kX71Dec, 6, 16, 8, {-30.0F, 30.0F, 0.0F, kNoAngle, -150.0F, 150.0F, -90.0F, 90.0F}},
};

// ---- panning helpers (oalsfxpp.cpp:483-767) ---------------------------------------------------
void angle_coeffs(float azimuth, float elevation, float spread, float coeffs[16])
{
	// calc_angle_coeffs -> calc_direction_coeffs (oalsfxpp.cpp:583-597, 483-577)
	const float dir[3] = {
		std::sin(azimuth) * std::cos(elevation),
		std::sin(elevation),
		-std::cos(azimuth) * std::cos(elevation),
	};
	const float x = -dir[2];
	const float y = -dir[0];
	const float z = dir[1];

	coeffs[0] = 1.0F;
	coeffs[1] = 1.732050808F * y;
	coeffs[2] = 1.732050808F * z;
	coeffs[3] = 1.732050808F * x;
	coeffs[4] = 3.872983346F * x * y;
	coeffs[5] = 3.872983346F * y * z;
	coeffs[6] = 1.118033989F * ((3.0F * z * z) - 1.0F);
	coeffs[7] = 3.872983346F * x * z;
	coeffs[8] = 1.936491673F * ((x * x) - (y * y));
	coeffs[9] = 2.091650066F * y * ((3.0F * x * x) - (y * y));
	coeffs[10] = 10.246950766F * z * x * y;
	coeffs[11] = 1.620185175F * y * ((5.0F * z * z) - 1.0F);
	coeffs[12] = 1.322875656F * z * ((5.0F * z * z) - 3.0F);
	coeffs[13] = 1.620185175F * x * ((5.0F * z * z) - 1.0F);
	coeffs[14] = 5.123475383F * z * ((x * x) - (y * y));
	coeffs[15] = 2.091650066F * x * ((x * x) - (3.0F * y * y));

	if (spread > 0.0F) {
		const float ca = std::cos(spread * 0.5F);
		const float scale = std::sqrt(1.0F + (spread / kTau));
		const float zh[4] = {
			scale,
			0.5F * (ca + 1.0F) * scale,
			0.5F * (ca + 1.0F) * ca * scale,
			0.125F * (ca + 1.0F) * ((5.0F * ca * ca) - 1.0F) * scale,
		};
		static const int order_of[16] = {0, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3};
		for (int i = 0; i < 16; ++i) {
			coeffs[i] *= zh[order_of[i]];
		}
	}
}

// compute_panning_gains_mc (oalsfxpp.cpp:670-696)
void panning_gains(const DeviceLayout& dev, const float coeffs[16], float in_gain, float out[kMaxChannels])
{
	for (int i = 0; i < kMaxChannels; ++i) {
		if (i < dev.channels) {
			float gain = 0.0F;
			for (int j = 0; j < dev.dry_coeff_count; ++j) {
				gain += dev.dry[i][j] * coeffs[j];
			}
			out[i] = clampf(gain, 0.0F, 1.0F) * in_gain;
		} else {
			out[i] = 0.0F;
		}
	}
}

// compute_first_order_gains_mc on the FOA decoder (oalsfxpp.cpp:730-755)
void first_order_gains(const DeviceLayout& dev, const float matrix[4], float in_gain, float out[kMaxChannels])
{
	for (int i = 0; i < kMaxChannels; ++i) {
		if (i < dev.channels) {
			float gain = 0.0F;
			for (int j = 0; j < 4; ++j) {
				gain += dev.foa[i][j] * matrix[j];
			}
			out[i] = clampf(gain, 0.0F, 1.0F) * in_gain;
		} else {
			out[i] = 0.0F;
		}
	}
}

// compute_ambient_gains_mc (oalsfxpp.cpp:616-626)
void ambient_gains(const DeviceLayout& dev, float in_gain, float out[kMaxChannels])
{
	for (int i = 0; i < kMaxChannels; ++i) {
		out[i] = (i < dev.channels ? dev.dry[i][0] * 1.414213562F * in_gain : 0.0F);
	}
}

void identity_first_order_gains(const DeviceLayout& dev, float out[kWetChannels][kMaxChannels])
{
	for (int i = 0; i < kWetChannels; ++i) {
		float row[4] = {0.0F, 0.0F, 0.0F, 0.0F};
		row[i] = 1.0F;
		first_order_gains(dev, row, 1.0F, out[i]);
	}
}

// ---- biquad design (oalsfxpp.cpp:867-982, 1074-1090) ------------------------------------------
enum BiquadKind { kHighShelf, kLowShelf, kPeaking, kLowPass, kHighPass, kBandPass };

thread_local bool g_saw_unstable_design = false;

Biquad design_biquad(BiquadKind kind, float gain, float freq_mult, float rcp_q)
{
	if (!(freq_mult < 0.5F)) {
		g_saw_unstable_design = true;
	}
	const float w0 = kTau * freq_mult;
	const float sin_w0 = std::sin(w0);
	const float cos_w0 = std::cos(w0);
	const float alpha = sin_w0 / 2.0F * rcp_q;
	float a[3] = {1.0F, 0.0F, 0.0F};
	float b[3] = {1.0F, 0.0F, 0.0F};
	switch (kind) {
	case kHighShelf: {
		const float sa = 2.0F * std::sqrt(gain) * alpha;
		b[0] = gain * ((gain + 1.0F) + ((gain - 1.0F) * cos_w0) + sa);
		b[1] = -2.0F * gain * ((gain - 1.0F) + ((gain + 1.0F) * cos_w0));
		b[2] = gain * ((gain + 1.0F) + ((gain - 1.0F) * cos_w0) - sa);
		a[0] = (gain + 1.0F) - ((gain - 1.0F) * cos_w0) + sa;
		a[1] = 2.0F * ((gain - 1.0F) - ((gain + 1.0F) * cos_w0));
		a[2] = (gain + 1.0F) - ((gain - 1.0F) * cos_w0) - sa;
		break;
	}
	case kLowShelf: {
		const float sa = 2.0F * std::sqrt(gain) * alpha;
		b[0] = gain * ((gain + 1.0F) - ((gain - 1.0F) * cos_w0) + sa);
		b[1] = 2.0F * gain * ((gain - 1.0F) - ((gain + 1.0F) * cos_w0));
		b[2] = gain * ((gain + 1.0F) - ((gain - 1.0F) * cos_w0) - sa);
		a[0] = (gain + 1.0F) + ((gain - 1.0F) * cos_w0) + sa;
		a[1] = -2.0F * ((gain - 1.0F) + ((gain + 1.0F) * cos_w0));
		a[2] = (gain + 1.0F) + ((gain - 1.0F) * cos_w0) - sa;
		break;
	}
	case kPeaking: {
		const float sg = std::sqrt(gain);
		b[0] = 1.0F + (alpha * sg);
		b[1] = -2.0F * cos_w0;
		b[2] = 1.0F - (alpha * sg);
		a[0] = 1.0F + (alpha / sg);
		a[1] = -2.0F * cos_w0;
		a[2] = 1.0F - (alpha / sg);
		break;
	}
	case kLowPass:
		b[0] = (1.0F - cos_w0) / 2.0F;
		b[1] = 1.0F - cos_w0;
		b[2] = (1.0F - cos_w0) / 2.0F;
		a[0] = 1.0F + alpha;
		a[1] = -2.0F * cos_w0;
		a[2] = 1.0F - alpha;
		break;
	case kHighPass:
		b[0] = (1.0F + cos_w0) / 2.0F;
		b[1] = -(1.0F + cos_w0);
		b[2] = (1.0F + cos_w0) / 2.0F;
		a[0] = 1.0F + alpha;
		a[1] = -2.0F * cos_w0;
		a[2] = 1.0F - alpha;
		break;
	case kBandPass:
		b[0] = alpha;
		b[1] = 0;
		b[2] = -alpha;
		a[0] = 1.0F + alpha;
		a[1] = -2.0F * cos_w0;
		a[2] = 1.0F - alpha;
		break;
	}
	Biquad out;
	out.a1 = a[1] / a[0];
	out.a2 = a[2] / a[0];
	out.b0 = b[0] / a[0];
	out.b1 = b[1] / a[0];
	out.b2 = b[2] / a[0];
	return out;
}

float rcp_q_from_slope(float gain, float slope)
{
	return std::sqrt((gain + (1.0F / gain)) * ((1.0F / slope) - 1.0F) + 2.0F);
}

float rcp_q_from_bandwidth(float freq_mult, float bandwidth)
{
	const float w0 = kTau * freq_mult;
	return 2.0F * std::sinh(std::log(2.0F) / 2.0F * bandwidth * w0 / std::sin(w0));
}

// ---- per-effect derivation ---------------------------------------------------------------------
void side_gains(const DeviceLayout& dev, float out[2][kMaxChannels])
{
	// chorus/flanger left/right panning (oalsfxpp.cpp:4070-4076, 5344-5350)
	float coeffs[16];
	angle_coeffs(-kPi2, 0.0F, 0.0F, coeffs);
	panning_gains(dev, coeffs, 1.0F, out[0]);
	angle_coeffs(kPi2, 0.0F, 0.0F, coeffs);
	panning_gains(dev, coeffs, 1.0F, out[1]);
}

void derive_mod_delay(const DeviceLayout& dev, int rate_hz, float max_delay, int waveform, int phase, float rate,
	float depth, float feedback, float delay, ModDelayCoef& c, SlotTables& tables)
{
	// do_update_device (oalsfxpp.cpp:4020-4040 / 5291-5312): ring length
	int max_len = static_cast<int>(max_delay * 2.0F * rate_hz) + 1;
	max_len = next_pow2(max_len);
	c.mask = max_len - 1;

	// do_update (oalsfxpp.cpp:4042-4111 / 5314-5382)
	const float frequency = static_cast<float>(rate_hz);
	c.waveform = waveform;
	c.feedback = feedback;
	c.delay = static_cast<int>(delay * frequency);
	c.depth = depth * c.delay;
	side_gains(dev, c.gains);
	if (!(rate > 0.0F)) {
		c.lfo_scale = 0.0F;
		c.lfo_range = 1;
		c.lfo_disp = 0;
	} else {
		c.lfo_range = static_cast<int>(frequency / rate + 0.5F);
		c.lfo_scale = (waveform == 1 ? 4.0F / c.lfo_range : kTau / c.lfo_range);
		if (phase >= 0) {
			c.lfo_disp = static_cast<int>(c.lfo_range * (phase / 360.0F));
		} else {
			c.lfo_disp = static_cast<int>(c.lfo_range * ((360 + phase) / 360.0F));
		}
	}
	c.sin_delays = nullptr;
	if (waveform == 0) {
		// get_sinusoid_delays (oalsfxpp.cpp:4262-4276): the LFO delay is truncated to int, so one
		// ulp of sinf would flip a delay.  Evaluate with the host libm once per phase value; the
		// device indexes the table with its bit-exact integer phase.
		tables.sin_delays.resize(static_cast<size_t>(c.lfo_range));
		for (int p = 0; p < c.lfo_range; ++p) {
			tables.sin_delays[static_cast<size_t>(p)] = static_cast<int>(std::sin(c.lfo_scale * p) * c.depth) + c.delay;
		}
	}
}

void derive_compressor(const DeviceLayout& dev, int rate_hz, const oalsfxpp::EffectProps::Compressor& p, CompressorCoef& c)
{
	// oalsfxpp.cpp:4319-4350
	const float attack_time = rate_hz * 0.2F;
	const float release_time = rate_hz * 0.4F;
	c.attack_rate = 1.0F / attack_time;
	c.release_rate = 1.0F / release_time;
	c.enabled = p.on_off_ ? 1 : 0;
	identity_first_order_gains(dev, c.gains);
}

void derive_dedicated(const DeviceLayout& dev, int fx_type, const oalsfxpp::EffectProps::Dedicated& p, DedicatedCoef& c)
{
	// oalsfxpp.cpp:4509-4554.  Device::get_channel_index always answers -1 (its end iterator is
	// cbegin(), oalsfxpp.cpp:2577-2578), so the LFE variant is silent and dialog is panned to the
	// front-centre *direction*.  Reproduced, not fixed.
	for (float& g : c.gains) {
		g = 0.0F;
	}
	if (fx_type == kFxDedicatedDialog) {
		float coeffs[16];
		angle_coeffs(0.0F, 0.0F, 0.0F, coeffs);
		panning_gains(dev, coeffs, p.gain_, c.gains);
	}
}

void derive_distortion(const DeviceLayout& dev, int rate_hz, const oalsfxpp::EffectProps::Distortion& p, DistortionCoef& c)
{
	// oalsfxpp.cpp:4627-4673
	const float frequency = static_cast<float>(rate_hz);
	const float attenuation = p.gain_;
	float edge = std::sin(p.edge_ * kPi2);
	edge = std::min(edge, 0.99F);
	c.edge_coeff = 2.0F * edge / (1.0F - edge);
	float cutoff = p.low_pass_cutoff_;
	float bandwidth = (cutoff / 2.0F) / (cutoff * 0.67F);
	c.low_pass = design_biquad(kLowPass, 1.0F, cutoff / (frequency * 4.0F),
		rcp_q_from_bandwidth(cutoff / (frequency * 4.0F), bandwidth));
	cutoff = p.eq_center_;
	bandwidth = p.eq_bandwidth_ / (cutoff * 0.67F);
	c.band_pass = design_biquad(kBandPass, 1.0F, cutoff / (frequency * 4.0F),
		rcp_q_from_bandwidth(cutoff / (frequency * 4.0F), bandwidth));
	float amb[kMaxChannels];
	ambient_gains(dev, 1.0F, amb);
	for (int k = 0; k < kMaxChannels; ++k) {
		c.gains[k] = amb[k] * attenuation; // oalsfxpp.cpp:4736
	}
}

void derive_echo(const DeviceLayout& dev, int rate_hz, const oalsfxpp::EffectProps::Echo& p, EchoCoef& c)
{
	// ring length: oalsfxpp.cpp:4817-4833
	int maxlen = static_cast<int>(0.207F * rate_hz) + 1;
	maxlen += static_cast<int>(0.404F * rate_hz) + 1;
	maxlen = next_pow2(maxlen);
	c.mask = maxlen - 1;

	// oalsfxpp.cpp:4835-4885
	const int frequency = rate_hz;
	c.tap1 = static_cast<int>(p.delay_ * frequency) + 1;
	c.tap2 = static_cast<int>(p.lr_delay_ * frequency);
	c.tap2 += c.tap1;
	float spread = p.spread_;
	const float lrpan = (spread < 0.0F ? -1.0F : 1.0F);
	spread = std::asin(1.0F - std::abs(spread)) * 4.0F;
	c.feed_gain = p.feedback_;
	const float effect_gain = std::max(1.0F - p.damping_, 0.0625F);
	c.filter = design_biquad(kHighShelf, effect_gain, 5000.0F / frequency, rcp_q_from_slope(effect_gain, 1.0F));
	float coeffs[16];
	angle_coeffs(-kPi2 * lrpan, 0.0F, spread, coeffs);
	panning_gains(dev, coeffs, 1.0F, c.gains[0]);
	angle_coeffs(kPi2 * lrpan, 0.0F, spread, coeffs);
	panning_gains(dev, coeffs, 1.0F, c.gains[1]);
}

void derive_equalizer(const DeviceLayout& dev, int rate_hz, const oalsfxpp::EffectProps::Equalizer& p, EqualizerCoef& c)
{
	// oalsfxpp.cpp:5076-5159
	const float frequency = static_cast<float>(rate_hz);
	identity_first_order_gains(dev, c.gains);
	float gain = std::max(std::sqrt(p.low_gain_), 0.0625F);
	float freq_mult = p.low_cutoff_ / frequency;
	c.band[0] = design_biquad(kLowShelf, gain, freq_mult, rcp_q_from_slope(gain, 0.75F));
	gain = std::max(p.mid1_gain_, 0.0625F);
	freq_mult = p.mid1_center_ / frequency;
	c.band[1] = design_biquad(kPeaking, gain, freq_mult, rcp_q_from_bandwidth(freq_mult, p.mid1_width_));
	gain = std::max(p.mid2_gain_, 0.0625F);
	freq_mult = p.mid2_center_ / frequency;
	c.band[2] = design_biquad(kPeaking, gain, freq_mult, rcp_q_from_bandwidth(freq_mult, p.mid2_width_));
	gain = std::max(std::sqrt(p.high_gain_), 0.0625F);
	freq_mult = p.high_cutoff_ / frequency;
	c.band[3] = design_biquad(kHighShelf, gain, freq_mult, rcp_q_from_slope(gain, 0.75F));
}

void derive_ring_mod(const DeviceLayout& dev, int rate_hz, const oalsfxpp::EffectProps::RingModulator& p, RingModCoef& c)
{
	// oalsfxpp.cpp:5598-5650
	constexpr int frac_one = 1 << 24;
	c.waveform = p.waveform_;
	c.step = static_cast<int>(p.frequency_ * frac_one / rate_hz);
	if (c.step == 0) {
		c.step = 1;
	}
	const float cw = std::cos(kTau * p.high_pass_cutoff_ / rate_hz);
	const float a = (2.0F - cw) - std::sqrt(std::pow(2.0F - cw, 2.0F) - 1.0F);
	c.filter.b0 = a;
	c.filter.b1 = -a;
	c.filter.b2 = 0.0F;
	c.filter.a1 = -a;
	c.filter.a2 = 0.0F;
	identity_first_order_gains(dev, c.gains);
}

// ---- reverb (oalsfxpp.cpp:5799-7350) ----------------------------------------------------------
constexpr float kDecayGain = 0.001F;           // -60 dB, oalsfxpp.cpp:6174
constexpr float kLineMultiplier = 9.0F;        // oalsfxpp.cpp:6404
constexpr float kEarlyTapLen[4] = {0.000000E+0F, 1.010676E-3F, 2.126553E-3F, 3.358580E-3F};
constexpr float kEarlyApLen[4] = {4.854840E-4F, 5.360178E-4F, 5.918117E-4F, 6.534130E-4F};
constexpr float kEarlyLineLen[4] = {2.992520E-3F, 5.456575E-3F, 7.688329E-3F, 9.709681E-3F};
constexpr float kLateApLen[4] = {8.091400E-4F, 1.019453E-3F, 1.407968E-3F, 1.618280E-3F};
constexpr float kLateLineLen[4] = {9.709681E-3F, 1.223343E-2F, 1.689561E-2F, 1.941936E-2F};
constexpr float kModDepthCoeff = 1.0F / 4096.0F;
constexpr float kMaxReflectionsDelay = 0.3F;
constexpr float kMaxLateDelay = 0.1F;
constexpr float kMaxEchoTime = 0.25F;
constexpr float kMaxModTime = 4.0F;
constexpr float kSpeedOfSound = 343.3F;

int line_samples(float length, int frequency, int extra)
{
	// initialize_delay_line (oalsfxpp.cpp:6538-6552)
	const int n = static_cast<int>(std::ceil(length * frequency));
	return next_pow2(n + extra);
}

void reverb_ring_lengths(int frequency, int len[5])
{
	// alloc_lines (oalsfxpp.cpp:6556-6598)
	const float multiplier = 1.0F + kLineMultiplier;
	float length = kMaxReflectionsDelay + (kEarlyTapLen[3] * multiplier) + kMaxLateDelay +
		((kLateLineLen[3] - kLateLineLen[0]) * 0.25F * multiplier);
	len[0] = line_samples(length, frequency, 256);
	length = kEarlyApLen[3] * multiplier;
	len[1] = line_samples(length, frequency, 0);
	length = kEarlyLineLen[3] * multiplier;
	len[2] = line_samples(length, frequency, 0);
	length = kLateApLen[3] * multiplier;
	len[3] = line_samples(length, frequency, 0);
	length = std::max(kMaxEchoTime, kLateLineLen[3] * multiplier) + (kMaxModTime * kModDepthCoeff / 2.0F);
	len[4] = line_samples(length, frequency, 0);
}

float decay_coeff(float length, float decay_time) { return std::pow(kDecayGain, length / decay_time); }

float decay_length(float coeff, float decay_time)
{
	return std::log10(coeff) * decay_time / std::log10(kDecayGain);
}

void pass_through3(float c[3]) { c[0] = 1.0F; c[1] = 0.0F; c[2] = 0.0F; }

void highpass_coeffs(float gain, float w, float c[3])
{
	// oalsfxpp.cpp:6717-6738
	if (gain >= 1.0F) { pass_through3(c); return; }
	const float g = std::max(0.001F, gain);
	const float g2 = g * g;
	const float cw = std::cos(w);
	const float p = g / ((g * cw) + std::sqrt((cw - 1.0F) * ((g2 * cw) + g2 - 2.0F)));
	c[0] = p; c[1] = -p; c[2] = p;
}

void lowpass_coeffs(float gain, float w, float c[3])
{
	// oalsfxpp.cpp:6762-6786
	if (gain >= 1.0F) { pass_through3(c); return; }
	const float g = std::max(0.001F, gain);
	const float g2 = g * g;
	const float cw = std::cos(w);
	const float a = (1.0F - (g2 * cw) - std::sqrt((2.0F * g2 * (1.0F - cw)) - (g2 * g2 * (1.0F - (cw * cw))))) /
		(1.0F - g2);
	c[0] = 1.0F - a; c[1] = 0.0F; c[2] = a;
}

void shelf_core(float g, float p, float& alpha, float& beta0, float& beta1)
{
	const float n = (g + 1.0F) / (g - 1.0F);
	alpha = n + std::sqrt((n * n) - 1.0F);
	beta0 = (1.0F + g + (1.0F - g) * alpha) / 2.0F;
	beta1 = (1.0F - g + (1.0F + g) * alpha) / 2.0F;
	(void)p;
}

void low_shelf_coeffs(float gain, float w, float c[3])
{
	// oalsfxpp.cpp:6832-6857
	if (gain >= 1.0F) { pass_through3(c); return; }
	const float g = std::max(0.001F, gain);
	const float rw = kPi - w;
	const float p = std::sin((0.5F * rw) - (0.25F * kPi)) / std::sin((0.5F * rw) + (0.25F * kPi));
	float alpha, beta0, beta1;
	shelf_core(g, p, alpha, beta0, beta1);
	c[0] = (beta0 + (p * beta1)) / (1.0F + (p * alpha));
	c[1] = -(beta1 + (p * beta0)) / (1.0F + (p * alpha));
	c[2] = (p + alpha) / (1.0F + (p * alpha));
}

void high_shelf_coeffs(float gain, float w, float c[3])
{
	// oalsfxpp.cpp:6904-6927
	if (gain >= 1.0F) { pass_through3(c); return; }
	const float g = std::max(0.001F, gain);
	const float p = std::sin((0.5F * w) - (0.25F * kPi)) / std::sin((0.5F * w) + (0.25F * kPi));
	float alpha, beta0, beta1;
	shelf_core(g, p, alpha, beta0, beta1);
	c[0] = (beta0 + (p * beta1)) / (1.0F + (p * alpha));
	c[1] = (beta1 + (p * beta0)) / (1.0F + (p * alpha));
	c[2] = -(p + alpha) / (1.0F + (p * alpha));
}

void t60_coeffs(float length, float lf_t, float mf_t, float hf_t, float lf_w, float hf_w,
	float lf[3], float hf[3], float& mid)
{
	// calc_t60_damping_coeffs (oalsfxpp.cpp:6934-7010)
	const float lf_gain = decay_coeff(length, lf_t);
	const float mf_gain = decay_coeff(length, mf_t);
	const float hf_gain = decay_coeff(length, hf_t);
	if (lf_gain < mf_gain) {
		if (mf_gain < hf_gain) {
			low_shelf_coeffs(mf_gain / hf_gain, hf_w, lf);
			highpass_coeffs(lf_gain / mf_gain, lf_w, hf);
			mid = hf_gain;
		} else if (mf_gain > hf_gain) {
			highpass_coeffs(lf_gain / mf_gain, lf_w, lf);
			lowpass_coeffs(hf_gain / mf_gain, hf_w, hf);
			mid = mf_gain;
		} else {
			pass_through3(lf);
			highpass_coeffs(lf_gain / mf_gain, lf_w, hf);
			mid = mf_gain;
		}
	} else if (lf_gain > mf_gain) {
		if (mf_gain < hf_gain) {
			const float hg = mf_gain / lf_gain;
			const float lg = mf_gain / hf_gain;
			high_shelf_coeffs(hg, lf_w, lf);
			low_shelf_coeffs(lg, hf_w, hf);
			mid = std::max(lf_gain, hf_gain) / std::max(hg, lg);
		} else if (mf_gain > hf_gain) {
			high_shelf_coeffs(mf_gain / lf_gain, lf_w, lf);
			lowpass_coeffs(hf_gain / mf_gain, hf_w, hf);
			mid = lf_gain;
		} else {
			pass_through3(lf);
			high_shelf_coeffs(mf_gain / lf_gain, lf_w, hf);
			mid = lf_gain;
		}
	} else {
		pass_through3(lf);
		if (mf_gain < hf_gain) {
			low_shelf_coeffs(mf_gain / hf_gain, hf_w, hf);
			mid = hf_gain;
		} else if (mf_gain > hf_gain) {
			lowpass_coeffs(hf_gain / mf_gain, hf_w, hf);
			mid = mf_gain;
		} else {
			pass_through3(hf);
			mid = mf_gain;
		}
	}
}

struct Mat4 { float m[4][4]; };

Mat4 mat_mul(const Mat4& a, const Mat4& b, bool transpose_result)
{
	// matrix_mult / matrix_mult_t (oalsfxpp.cpp:7189-7210, 7283-7304)
	Mat4 r;
	for (int col = 0; col < 4; ++col) {
		for (int row = 0; row < 4; ++row) {
			const float v = (a.m[row][0] * b.m[0][col]) + (a.m[row][1] * b.m[1][col]) +
				(a.m[row][2] * b.m[2][col]) + (a.m[row][3] * b.m[3][col]);
			if (transpose_result) {
				r.m[col][row] = v;
			} else {
				r.m[row][col] = v;
			}
		}
	}
	return r;
}

Mat4 transform_from_vector(const float vec[3])
{
	// get_transform_from_vector (oalsfxpp.cpp:7229-7281)
	const float length = std::sqrt((vec[0] * vec[0]) + (vec[1] * vec[1]) + (vec[2] * vec[2]));
	const float sa = std::sin(std::min(length, 1.0F) * (kPi / 4.0F));
	const Mat4 zfocus = {{
		{1.0F / (1.0F + sa), 0.0F, 0.0F, (sa / (1.0F + sa)) / 1.732050808F},
		{0.0F, std::sqrt((1.0F - sa) / (1.0F + sa)), 0.0F, 0.0F},
		{0.0F, 0.0F, std::sqrt((1.0F - sa) / (1.0F + sa)), 0.0F},
		{(sa / (1.0F + sa)) * 1.732050808F, 0.0F, 0.0F, 1.0F / (1.0F + sa)},
	}};
	float a = std::atan2(vec[1], std::sqrt((vec[0] * vec[0]) + (vec[2] * vec[2])));
	const Mat4 xrot = {{
		{1.0F, 0.0F, 0.0F, 0.0F},
		{0.0F, 1.0F, 0.0F, 0.0F},
		{0.0F, 0.0F, std::cos(a), std::sin(a)},
		{0.0F, 0.0F, -std::sin(a), std::cos(a)},
	}};
	a = std::atan2(-vec[0], vec[2]);
	const Mat4 yrot = {{
		{1.0F, 0.0F, 0.0F, 0.0F},
		{0.0F, std::cos(a), 0.0F, std::sin(a)},
		{0.0F, 0.0F, 1.0F, 0.0F},
		{0.0F, -std::sin(a), 0.0F, std::cos(a)},
	}};
	return mat_mul(yrot, mat_mul(xrot, zfocus, false), false);
}

void derive_reverb(const DeviceLayout& dev, int frequency, bool is_eax, const oalsfxpp::EffectProps::Reverb& p,
	ReverbCoef& c, SlotTables& tables)
{
	c.is_eax = is_eax ? 1 : 0;

	// ---- do_update_device (oalsfxpp.cpp:5928-5950) ----
	int len[5];
	reverb_ring_lengths(frequency, len);
	int base = 0;
	for (int r = 0; r < 5; ++r) {
		c.mask[r] = len[r] - 1;
		c.ring_base[r] = base;
		base += 4 * len[r];
	}
	c.mod_coeff = std::pow(0.048F, 100000.0F / frequency);
	c.late_feed_tap = static_cast<int>((kMaxReflectionsDelay + (kEarlyTapLen[3] * (1.0F + kLineMultiplier))) * frequency);

	// ---- do_update (oalsfxpp.cpp:5952-6076) ----
	const float hf_scale = p.hf_reference_ / frequency;
	const float gain_hf = std::max(p.gain_hf_, 0.001F);
	c.lp = design_biquad(kHighShelf, gain_hf, hf_scale, rcp_q_from_slope(gain_hf, 1.0F));
	const float lf_scale = p.lf_reference_ / frequency;
	const float gain_lf = std::max(p.gain_lf_, 0.001F);
	c.hp = design_biquad(kLowShelf, gain_lf, lf_scale, rcp_q_from_slope(gain_lf, 1.0F));

	const float multiplier = 1.0F + (p.density_ * kLineMultiplier);

	// update_delay_line (oalsfxpp.cpp:7046-7076)
	for (int i = 0; i < 4; ++i) {
		float length = p.reflections_delay_ + (kEarlyTapLen[i] * multiplier);
		c.early_tap[i] = static_cast<int>(length * frequency);
		length = kEarlyTapLen[i] * multiplier;
		c.early_tap_coeff[i] = decay_coeff(length, p.decay_time_);
		length = p.late_reverb_delay_ + (kLateLineLen[i] - kLateLineLen[0]) * 0.25F * multiplier;
		c.late_tap[i] = c.late_feed_tap + static_cast<int>(length * frequency);
	}

	c.ap_feed_coeff = std::sqrt(0.5F) * std::pow(p.diffusion_, 2.0F);

	// update_early_lines (oalsfxpp.cpp:7078-7101)
	for (int i = 0; i < 4; ++i) {
		float length = kEarlyApLen[i] * multiplier;
		c.early_ap_off[i] = static_cast<int>(length * frequency);
		length = kEarlyLineLen[i] * multiplier;
		c.early_off[i] = static_cast<int>(length * frequency);
		c.early_coeff[i] = decay_coeff(length, p.decay_time_);
	}

	// calc_matrix_coeffs (oalsfxpp.cpp:6645-6659)
	{
		const float n = std::sqrt(3.0F);
		const float t = p.diffusion_ * std::atan(n);
		c.mix_x = std::cos(t);
		c.mix_y = std::sin(t) / n;
	}

	// HF ratio limit (oalsfxpp.cpp:6009-6017, 6663-6679)
	float hf_ratio = p.decay_hf_ratio_;
	if (p.decay_hf_limit_ && p.air_absorption_gain_hf_ < 1.0F) {
		const float limit_ratio = 1.0F / (decay_length(p.air_absorption_gain_hf_, p.decay_time_) * kSpeedOfSound);
		hf_ratio = clampf(limit_ratio, 0.1F, hf_ratio);
	}
	const float lf_decay_time = clampf(p.decay_time_ * p.decay_lf_ratio_, 0.1F, 20.0F);
	const float hf_decay_time = clampf(p.decay_time_ * hf_ratio, 0.1F, 20.0F);

	// update_modulator (oalsfxpp.cpp:7014-7043).  The index rescale needs the stream's running
	// index and therefore happens on the device when the block's update flag is set.
	c.mod_range = std::max(static_cast<int>(p.modulation_time_ * frequency), 1);
	c.mod_depth = p.modulation_depth_ * kModDepthCoeff * p.modulation_time_ / 2.0F * frequency;
	c.mod_sinus = nullptr;
	// calc_modulation_delays (oalsfxpp.cpp:7443-7470) rounds range*sinus to an integer delay, so the
	// sinus must come from the host libm: table it over the index range.
	tables.mod_sinus.resize(static_cast<size_t>(c.mod_range));
	for (int i = 0; i < c.mod_range; ++i) {
		tables.mod_sinus[static_cast<size_t>(i)] = std::sin(kTau * i / c.mod_range);
	}

	// update_late_lines (oalsfxpp.cpp:7103-7187)
	{
		const float lf_w = kTau * lf_scale;
		const float hf_w = kTau * hf_scale;
		float length = (kLateLineLen[0] + kLateLineLen[1] + kLateLineLen[2] + kLateLineLen[3]) / 4.0F * multiplier;
		length = lerpf(length, p.echo_time_, p.echo_depth_);
		length += (kLateApLen[0] + kLateApLen[1] + kLateApLen[2] + kLateApLen[3]) / 4.0F * multiplier;
		float band_weights[3];
		band_weights[0] = lf_w;
		band_weights[1] = hf_w - lf_w;
		band_weights[2] = kTau - hf_w;
		const float a = decay_coeff(length,
			((band_weights[0] * lf_decay_time) + (band_weights[1] * p.decay_time_) + (band_weights[2] * hf_decay_time)) / kTau);
		c.density_gain = std::sqrt(1.0F - (a * a));
		for (int i = 0; i < 4; ++i) {
			length = kLateApLen[i] * multiplier;
			c.late_ap_off[i] = static_cast<int>(length * frequency);
			length = lerpf(kLateLineLen[i] * multiplier, p.echo_time_, p.echo_depth_);
			c.late_off[i] = static_cast<int>(length * frequency);
			length += lerpf(kLateApLen[i],
				(kLateApLen[0] + kLateApLen[1] + kLateApLen[2] + kLateApLen[3]) / 4.0F, p.diffusion_) * multiplier;
			t60_coeffs(length, lf_decay_time, p.decay_time_, hf_decay_time, lf_w, hf_w,
				c.t60_lf[i], c.t60_hf[i], c.t60_mid[i]);
		}
	}

	// update_3d_panning (oalsfxpp.cpp:7306-7350)
	{
		static const Mat4 a2b = {{
			{0.866025403785F, 0.866025403785F, 0.866025403785F, 0.866025403785F},
			{0.866025403785F, -0.866025403785F, 0.866025403785F, -0.866025403785F},
			{0.866025403785F, -0.866025403785F, -0.866025403785F, 0.866025403785F},
			{0.866025403785F, 0.866025403785F, -0.866025403785F, -0.866025403785F},
		}};
		Mat4 rot = transform_from_vector(p.reflections_pan_.data());
		Mat4 transform = mat_mul(rot, a2b, true);
		for (int i = 0; i < 4; ++i) {
			first_order_gains(dev, transform.m[i], p.gain_ * p.reflections_gain_, c.pan_early[i]);
		}
		rot = transform_from_vector(p.late_reverb_pan_.data());
		transform = mat_mul(rot, a2b, true);
		for (int i = 0; i < 4; ++i) {
			first_order_gains(dev, transform.m[i], p.gain_ * p.late_reverb_gain_, c.pan_late[i]);
		}
	}
}

} // namespace

// ---- public -----------------------------------------------------------------------------------
bool make_device_layout(int channel_format, DeviceLayout& out)
{
	if (channel_format <= 0 || channel_format > 7) {
		return false;
	}
	const LayoutDesc& d = kLayouts[channel_format];
	out = DeviceLayout{};
	out.channel_format = channel_format;
	out.channels = d.channels;
	out.dry_coeff_count = d.coeff_count;
	// set_channel_map (oalsfxpp.cpp:769-807): LFE rows stay zero, others copy their decoder row.
	for (int i = 0; i < d.channels; ++i) {
		if (d.order[i] == kLFE) {
			continue;
		}
		for (int j = 0; j < d.dec_rows; ++j) {
			if (d.dec[j].spk == d.order[i]) {
				for (int k = 0; k < 16; ++k) {
					out.dry[i][k] = d.dec[j].c[k];
				}
				break;
			}
		}
		for (int k = 0; k < 4; ++k) {
			out.foa[i][k] = out.dry[i][k]; // oalsfxpp.cpp:2557-2568
		}
	}
	out.source_channels = d.map_count;
	for (int i = 0; i < d.map_count; ++i) {
		out.source_is_lfe[i] = (d.order[i] == kLFE);
		out.source_angle[i] = d.map_deg[i] * (kPi / 180.0F); // Math::deg_to_rad, oalsfxpp.cpp:156-160
	}
	return true;
}

void derive_sends(
	const DeviceLayout& dev, int sampling_rate, int effect_count,
	const SendSettings& direct, const SendSettings* aux,
	SendCoef& direct_out, SendCoef* aux_out)
{
	// calc_non_attn_source_params (oalsfxpp.cpp:3348-3395): gains capped at +24 dB
	const float dry_gain = std::min(direct.gain, 16.0F);
	std::memset(&direct_out, 0, sizeof(SendCoef));
	for (int i = 0; i < effect_count; ++i) {
		std::memset(&aux_out[i], 0, sizeof(SendCoef));
	}

	// calc_panning_and_filters (oalsfxpp.cpp:3172-3346)
	for (int c = 0; c < dev.source_channels; ++c) {
		if (dev.source_is_lfe[c]) {
			continue; // LFE input is dropped: get_channel_index answers -1 (oalsfxpp.cpp:3233-3247)
		}
		float coeffs[16];
		angle_coeffs(dev.source_angle[c], 0.0F, 0.0F, coeffs);
		panning_gains(dev, coeffs, dry_gain, direct_out.gains[c]);
		for (int i = 0; i < effect_count; ++i) {
			const float wet_gain = std::min(aux[i].gain, 16.0F);
			for (int k = 0; k < kMaxChannels; ++k) {
				aux_out[i].gains[c][k] = (k < kWetChannels ? coeffs[k] * wet_gain : 0.0F);
			}
		}
	}

	// NB: hf_scale uses the 250 Hz constant and lf_scale the 5 kHz one (oalsfxpp.cpp:3271-3272).
	const float hf_scale = 250.0F / sampling_rate;
	const float lf_scale = 5000.0F / sampling_rate;
	auto filters = [&](const SendSettings& s, SendCoef& out) {
		const float gain_hf = std::max(s.gain_hf, 0.001F);
		const float gain_lf = std::max(s.gain_lf, 0.001F);
		out.filter_type = 0;
		if (gain_hf != 1.0F) {
			out.filter_type |= 1;
		}
		if (gain_lf != 1.0F) {
			out.filter_type |= 2;
		}
		out.lp = design_biquad(kHighShelf, gain_hf, hf_scale, rcp_q_from_slope(gain_hf, 1.0F));
		out.hp = design_biquad(kLowShelf, gain_lf, lf_scale, rcp_q_from_slope(gain_lf, 1.0F));
	};
	filters(direct, direct_out);
	for (int i = 0; i < effect_count; ++i) {
		filters(aux[i], aux_out[i]);
	}
}

int ring_words_for(int fx_type, int sampling_rate)
{
	switch (fx_type) {
	case kFxChorus:
		return 2 * next_pow2(static_cast<int>(0.016F * 2.0F * sampling_rate) + 1);
	case kFxFlanger:
		return 2 * next_pow2(static_cast<int>(0.004F * 2.0F * sampling_rate) + 1);
	case kFxEcho: {
		int maxlen = static_cast<int>(0.207F * sampling_rate) + 1;
		maxlen += static_cast<int>(0.404F * sampling_rate) + 1;
		return next_pow2(maxlen);
	}
	case kFxReverb:
	case kFxEaxReverb: {
		int len[5];
		reverb_ring_lengths(sampling_rate, len);
		return 4 * (len[0] + len[1] + len[2] + len[3] + len[4]);
	}
	default:
		return 0;
	}
}

void derive_slot(
	const DeviceLayout& dev, int sampling_rate, int fx_type, const oalsfxpp::EffectProps& props,
	SlotCoef& out, SlotTables& tables)
{
	std::memset(&out, 0, sizeof(out));
	out.type = fx_type;
	tables.sin_delays.clear();
	tables.mod_sinus.clear();
	g_saw_unstable_design = false;
	struct FlagOnExit {
		SlotCoef& c;
		~FlagOnExit() { c.flags = g_saw_unstable_design ? kCoefUnstable : 0U; }
	} flag_on_exit{out};
	switch (fx_type) {
	case kFxChorus: {
		const auto& p = props.chorus_;
		derive_mod_delay(dev, sampling_rate, 0.016F, p.waveform_, p.phase_, p.rate_, p.depth_, p.feedback_, p.delay_,
			out.u.mod_delay, tables);
		break;
	}
	case kFxFlanger: {
		const auto& p = props.flanger_;
		derive_mod_delay(dev, sampling_rate, 0.004F, p.waveform_, p.phase_, p.rate_, p.depth_, p.feedback_, p.delay_,
			out.u.mod_delay, tables);
		break;
	}
	case kFxCompressor:
		derive_compressor(dev, sampling_rate, props.compressor_, out.u.compressor);
		break;
	case kFxDedicatedDialog:
	case kFxDedicatedLfe:
		derive_dedicated(dev, fx_type, props.dedicated_, out.u.dedicated);
		break;
	case kFxDistortion:
		derive_distortion(dev, sampling_rate, props.distortion_, out.u.distortion);
		break;
	case kFxEcho:
		derive_echo(dev, sampling_rate, props.echo_, out.u.echo);
		break;
	case kFxEqualizer:
		derive_equalizer(dev, sampling_rate, props.equalizer_, out.u.equalizer);
		break;
	case kFxRingModulator:
		derive_ring_mod(dev, sampling_rate, props.ring_modulator_, out.u.ring_mod);
		break;
	case kFxReverb:
	case kFxEaxReverb:
		derive_reverb(dev, sampling_rate, fx_type == kFxEaxReverb, props.reverb_, out.u.reverb, tables);
		break;
	default:
		break;
	}
}

} // namespace oalsfx
