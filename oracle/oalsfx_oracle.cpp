// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product (oalsfxpp_b200/).
//
// oalsfx_oracle.cpp: a self-contained, single-stream, scalar CPU restatement of the reference's
// per-buffer effect path (bibendovsky/oalsfxpp, src/oalsfxpp.cpp), written from the algorithm
// description in SURVEY.md section 3/8a.  Block-structured like the reference (de-interleave ->
// sends -> per-slot process over <=2048-frame chunks with scratch buffers), i.e. deliberately a
// DIFFERENT formulation from the product's sample-serial fused kernel, so that agreement between
// the two is meaningful.  Every routine cites the reference lines it follows.
//
// Pinning: tests/test_oracle.py checks this file (a) bit-for-bit against the compiled reference
// (oracle/_ref, whenever /root/reference or a prebuilt oracle/_ref is there) over the whole case
// matrix and (b) against the committed golden vectors tests/golden/*.npz that were generated from
// the compiled reference by tests/golden/make_golden.py.  It exports the same `orc_*` C ABI as
// oracle/ref_shim.cpp.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
// this library.  Build: g++ -O2 -ffp-contract=off (FMA contraction breaks parity).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace {

// ---- constants (oalsfxpp.cpp:44-68, 149-153) ----------------------------------------------------
const int MAX_CH = 8;
const int MAX_FX = 4;
const int WET_CH = 4;
const int CHUNK = 2048;
const float SILENCE = 0.00001F;
const float MAX_GAIN = 16.0F;
const float PI = 3.14159265358979323846F;
const float PI_2 = 1.57079632679489661923F;
const float TAU = 6.28318530717958647692F;

enum Type { NUL, CHORUS, COMPRESSOR, DIALOG, LFE_FX, DISTORTION, ECHO, EQUALIZER, FLANGER, RINGMOD, REVERB, EAXREVERB };

// ---- property PODs (layout of oalsfxpp.h:65-530; limits oalsfxpp.h:76-511) -------------------------
struct PChorus { int waveform, phase; float rate, depth, feedback, delay; };
struct PCompressor { bool on_off; };
struct PDedicated { float gain; };
struct PDistortion { float edge, gain, lp_cutoff, eq_center, eq_bandwidth; };
struct PEcho { float delay, lr_delay, damping, feedback, spread; };
struct PEqualizer { float low_cutoff, low_gain, mid1_center, mid1_gain, mid1_width, mid2_center, mid2_gain, mid2_width,
	high_cutoff, high_gain; };
struct PReverb {
	float density, diffusion, gain, gain_hf, gain_lf, decay_time, decay_hf_ratio, decay_lf_ratio, refl_gain, refl_delay;
	float refl_pan[3];
	float late_gain, late_delay;
	float late_pan[3];
	float echo_time, echo_depth, mod_time, mod_depth, air_gain_hf, hf_ref, lf_ref, rolloff;
	bool hf_limit;
};
struct PRingMod { float frequency, hp_cutoff; int waveform; };
union Props {
	PChorus chorus; PCompressor compressor; PDedicated dedicated; PDistortion distortion; PEcho echo;
	PEqualizer equalizer; PReverb reverb; PRingMod ringmod;
};
static_assert(sizeof(Props) == 108, "EffectProps layout");
struct Effect { int type; Props props; };
struct Send3 { float gain, gain_hf, gain_lf; };

template <class T> T clampv(T v, T lo, T hi) { return std::min(hi, std::max(lo, v)); }

// Effect::set_defaults (oalsfxpp.cpp:1409-1725)
void set_defaults(Effect& e)
{
	Props& p = e.props;
	switch (e.type) {
	case CHORUS: p.chorus = PChorus{1, 90, 1.1F, 0.1F, 0.25F, 0.016F}; break;
	case FLANGER: p.chorus = PChorus{1, 0, 0.27F, 1.0F, -0.5F, 0.002F}; break;
	case COMPRESSOR: p.compressor.on_off = true; break;
	case DIALOG: case LFE_FX: p.dedicated.gain = 1.0F; break;
	case DISTORTION: p.distortion = PDistortion{0.2F, 0.05F, 8000.0F, 3600.0F, 3600.0F}; break;
	case ECHO: p.echo = PEcho{0.1F, 0.1F, 0.5F, 0.5F, -1.0F}; break;
	case EQUALIZER: p.equalizer = PEqualizer{200.0F, 1.0F, 500.0F, 1.0F, 1.0F, 3000.0F, 1.0F, 1.0F, 6000.0F, 1.0F}; break;
	case RINGMOD: p.ringmod = PRingMod{440.0F, 800.0F, 0}; break;
	case REVERB: case EAXREVERB: {
		PReverb& r = p.reverb;
		r.density = 1.0F; r.diffusion = 1.0F; r.gain = 0.32F; r.gain_hf = 0.89F; r.gain_lf = 1.0F;
		r.decay_time = 1.49F; r.decay_hf_ratio = 0.83F; r.decay_lf_ratio = 1.0F; r.refl_gain = 0.05F;
		r.refl_delay = 0.007F; r.refl_pan[0] = r.refl_pan[1] = r.refl_pan[2] = 0.0F; r.late_gain = 1.26F;
		r.late_delay = 0.011F; r.late_pan[0] = r.late_pan[1] = r.late_pan[2] = 0.0F; r.echo_time = 0.25F;
		r.echo_depth = 0.0F; r.mod_time = 0.25F; r.mod_depth = 0.0F; r.air_gain_hf = 0.994F; r.hf_ref = 5000.0F;
		r.lf_ref = 250.0F; r.rolloff = 0.0F; r.hf_limit = true;
		break;
	}
	default: break;
	}
}

// Effect::normalize (oalsfxpp.cpp:1419-1710, 1790-1833)
void normalize(Effect& e)
{
	Props& p = e.props;
	switch (e.type) {
	case CHORUS: case FLANGER: {
		PChorus& c = p.chorus;
		c.waveform = clampv(c.waveform, 0, 1); c.phase = clampv(c.phase, -180, 180);
		c.rate = clampv(c.rate, 0.0F, 10.0F); c.depth = clampv(c.depth, 0.0F, 1.0F);
		c.feedback = clampv(c.feedback, -1.0F, 1.0F);
		c.delay = clampv(c.delay, 0.0F, e.type == CHORUS ? 0.016F : 0.004F);
		break;
	}
	case DIALOG: case LFE_FX: p.dedicated.gain = clampv(p.dedicated.gain, 0.0F, 1.0F); break;
	case DISTORTION: {
		PDistortion& d = p.distortion;
		d.edge = clampv(d.edge, 0.0F, 1.0F); d.gain = clampv(d.gain, 0.01F, 1.0F);
		d.lp_cutoff = clampv(d.lp_cutoff, 80.0F, 24000.0F); d.eq_center = clampv(d.eq_center, 80.0F, 24000.0F);
		d.eq_bandwidth = clampv(d.eq_bandwidth, 80.0F, 24000.0F);
		break;
	}
	case ECHO: {
		PEcho& c = p.echo;
		c.delay = clampv(c.delay, 0.0F, 0.207F); c.lr_delay = clampv(c.lr_delay, 0.0F, 0.404F);
		c.damping = clampv(c.damping, 0.0F, 0.99F); c.feedback = clampv(c.feedback, 0.0F, 1.0F);
		c.spread = clampv(c.spread, -1.0F, 1.0F);
		break;
	}
	case EQUALIZER: {
		PEqualizer& q = p.equalizer;
		q.low_cutoff = clampv(q.low_cutoff, 50.0F, 800.0F); q.low_gain = clampv(q.low_gain, 0.126F, 7.943F);
		q.mid1_center = clampv(q.mid1_center, 200.0F, 3000.0F); q.mid1_gain = clampv(q.mid1_gain, 0.126F, 7.943F);
		q.mid1_width = clampv(q.mid1_width, 0.01F, 1.0F); q.mid2_center = clampv(q.mid2_center, 1000.0F, 8000.0F);
		q.mid2_gain = clampv(q.mid2_gain, 0.126F, 7.943F); q.mid2_width = clampv(q.mid2_width, 0.01F, 1.0F);
		q.high_cutoff = clampv(q.high_cutoff, 4000.0F, 16000.0F); q.high_gain = clampv(q.high_gain, 0.126F, 7.943F);
		break;
	}
	case RINGMOD: {
		PRingMod& r = p.ringmod;
		r.frequency = clampv(r.frequency, 0.0F, 8000.0F); r.hp_cutoff = clampv(r.hp_cutoff, 0.0F, 24000.0F);
		r.waveform = clampv(r.waveform, 0, 2);
		break;
	}
	case REVERB: case EAXREVERB: {
		PReverb& r = p.reverb;
		r.density = clampv(r.density, 0.0F, 1.0F); r.diffusion = clampv(r.diffusion, 0.0F, 1.0F);
		r.gain = clampv(r.gain, 0.0F, 1.0F); r.gain_hf = clampv(r.gain_hf, 0.0F, 1.0F);
		r.gain_lf = clampv(r.gain_lf, 0.0F, 1.0F); r.decay_time = clampv(r.decay_time, 0.1F, 20.0F);
		r.decay_hf_ratio = clampv(r.decay_hf_ratio, 0.1F, 2.0F); r.decay_lf_ratio = clampv(r.decay_lf_ratio, 0.1F, 2.0F);
		r.refl_gain = clampv(r.refl_gain, 0.0F, 3.16F); r.refl_delay = clampv(r.refl_delay, 0.0F, 0.3F);
		r.late_gain = clampv(r.late_gain, 0.0F, 10.0F); r.late_delay = clampv(r.late_delay, 0.0F, 0.1F);
		for (int i = 0; i < 3; ++i) {
			r.refl_pan[i] = clampv(r.refl_pan[i], -1.0F, 1.0F);
			r.late_pan[i] = clampv(r.late_pan[i], -1.0F, 1.0F);
		}
		r.echo_time = clampv(r.echo_time, 0.075F, 0.25F); r.echo_depth = clampv(r.echo_depth, 0.0F, 1.0F);
		r.mod_time = clampv(r.mod_time, 0.04F, 4.0F); r.mod_depth = clampv(r.mod_depth, 0.0F, 1.0F);
		r.air_gain_hf = clampv(r.air_gain_hf, 0.892F, 1.0F); r.hf_ref = clampv(r.hf_ref, 1000.0F, 20000.0F);
		r.lf_ref = clampv(r.lf_ref, 20.0F, 1000.0F); r.rolloff = clampv(r.rolloff, 0.0F, 10.0F);
		break;
	}
	default: break; // compressor: nothing to clamp (oalsfxpp.cpp:1451)
	}
}

// Effect::are_equal (oalsfxpp.cpp:1835-1893): field-wise comparison of the active block
bool effects_equal(const Effect& a, const Effect& b)
{
	if (a.type != b.type) return false;
	switch (a.type) {
	case NUL: return true;
	case CHORUS: case FLANGER: return std::memcmp(&a.props.chorus, &b.props.chorus, sizeof(PChorus)) == 0 ||
		(a.props.chorus.waveform == b.props.chorus.waveform && a.props.chorus.phase == b.props.chorus.phase &&
		 a.props.chorus.rate == b.props.chorus.rate && a.props.chorus.depth == b.props.chorus.depth &&
		 a.props.chorus.feedback == b.props.chorus.feedback && a.props.chorus.delay == b.props.chorus.delay);
	case COMPRESSOR: return a.props.compressor.on_off == b.props.compressor.on_off;
	case DIALOG: case LFE_FX: return a.props.dedicated.gain == b.props.dedicated.gain;
	case DISTORTION: { const PDistortion &x = a.props.distortion, &y = b.props.distortion;
		return x.edge == y.edge && x.gain == y.gain && x.lp_cutoff == y.lp_cutoff && x.eq_center == y.eq_center &&
			x.eq_bandwidth == y.eq_bandwidth; }
	case ECHO: { const PEcho &x = a.props.echo, &y = b.props.echo;
		return x.delay == y.delay && x.lr_delay == y.lr_delay && x.damping == y.damping && x.feedback == y.feedback &&
			x.spread == y.spread; }
	case EQUALIZER: { const float* x = &a.props.equalizer.low_cutoff; const float* y = &b.props.equalizer.low_cutoff;
		for (int i = 0; i < 10; ++i) if (!(x[i] == y[i])) return false;
		return true; }
	case RINGMOD: { const PRingMod &x = a.props.ringmod, &y = b.props.ringmod;
		return x.frequency == y.frequency && x.hp_cutoff == y.hp_cutoff && x.waveform == y.waveform; }
	case REVERB: case EAXREVERB: { const float* x = &a.props.reverb.density; const float* y = &b.props.reverb.density;
		for (int i = 0; i < 26; ++i) if (!(x[i] == y[i])) return false;
		return a.props.reverb.hf_limit == b.props.reverb.hf_limit; }
	default: return false;
	}
}

void normalize(Send3& s)
{
	s.gain = clampv(s.gain, 0.0F, 1.0F); s.gain_hf = clampv(s.gain_hf, 0.0F, 1.0F); s.gain_lf = clampv(s.gain_lf, 0.0F, 1.0F);
}
bool sends_equal(const Send3& a, const Send3& b) { return a.gain == b.gain && a.gain_hf == b.gain_hf && a.gain_lf == b.gain_lf; }

int next_pow2(int v)
{
	if (v > 0) { v -= 1; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; }
	return v + 1;
}

// ---- biquad (oalsfxpp.cpp:828-1091) ---------------------------------------------------------------
enum FilterKind { HIGH_SHELF, LOW_SHELF, PEAKING, LOW_PASS, HIGH_PASS, BAND_PASS };

struct Filter {
	float x[2] = {0, 0}, y[2] = {0, 0};
	float b0 = 0, b1 = 0, b2 = 0, a1 = 0, a2 = 0;

	void set(FilterKind kind, float gain, float freq_mult, float rcp_q) // set_params, :867-982
	{
		const float w0 = TAU * freq_mult;
		const float sn = std::sin(w0), cs = std::cos(w0);
		const float alpha = sn / 2.0F * rcp_q;
		float a[3] = {1, 0, 0}, b[3] = {1, 0, 0};
		float sq;
		switch (kind) {
		case HIGH_SHELF:
			sq = 2.0F * std::sqrt(gain) * alpha;
			b[0] = gain * ((gain + 1.0F) + ((gain - 1.0F) * cs) + sq);
			b[1] = -2.0F * gain * ((gain - 1.0F) + ((gain + 1.0F) * cs));
			b[2] = gain * ((gain + 1.0F) + ((gain - 1.0F) * cs) - sq);
			a[0] = (gain + 1.0F) - ((gain - 1.0F) * cs) + sq;
			a[1] = 2.0F * ((gain - 1.0F) - ((gain + 1.0F) * cs));
			a[2] = (gain + 1.0F) - ((gain - 1.0F) * cs) - sq;
			break;
		case LOW_SHELF:
			sq = 2.0F * std::sqrt(gain) * alpha;
			b[0] = gain * ((gain + 1.0F) - ((gain - 1.0F) * cs) + sq);
			b[1] = 2.0F * gain * ((gain - 1.0F) - ((gain + 1.0F) * cs));
			b[2] = gain * ((gain + 1.0F) - ((gain - 1.0F) * cs) - sq);
			a[0] = (gain + 1.0F) + ((gain - 1.0F) * cs) + sq;
			a[1] = -2.0F * ((gain - 1.0F) + ((gain + 1.0F) * cs));
			a[2] = (gain + 1.0F) + ((gain - 1.0F) * cs) - sq;
			break;
		case PEAKING:
			sq = std::sqrt(gain);
			b[0] = 1.0F + (alpha * sq); b[1] = -2.0F * cs; b[2] = 1.0F - (alpha * sq);
			a[0] = 1.0F + (alpha / sq); a[1] = -2.0F * cs; a[2] = 1.0F - (alpha / sq);
			break;
		case LOW_PASS:
			b[0] = (1.0F - cs) / 2.0F; b[1] = 1.0F - cs; b[2] = (1.0F - cs) / 2.0F;
			a[0] = 1.0F + alpha; a[1] = -2.0F * cs; a[2] = 1.0F - alpha;
			break;
		case HIGH_PASS:
			b[0] = (1.0F + cs) / 2.0F; b[1] = -(1.0F + cs); b[2] = (1.0F + cs) / 2.0F;
			a[0] = 1.0F + alpha; a[1] = -2.0F * cs; a[2] = 1.0F - alpha;
			break;
		case BAND_PASS:
			b[0] = alpha; b[1] = 0; b[2] = -alpha;
			a[0] = 1.0F + alpha; a[1] = -2.0F * cs; a[2] = 1.0F - alpha;
			break;
		}
		a1 = a[1] / a[0]; a2 = a[2] / a[0]; b0 = b[0] / a[0]; b1 = b[1] / a[0]; b2 = b[2] / a[0];
	}

	void copy_coeffs(const Filter& f) { b0 = f.b0; b1 = f.b1; b2 = f.b2; a1 = f.a1; a2 = f.a2; }

	float tick(float in) // one step of process(), :984-1036
	{
		const float out = (b0 * in) + (b1 * x[0]) + (b2 * x[1]) - (a1 * y[0]) - (a2 * y[1]);
		x[1] = x[0]; x[0] = in; y[1] = y[0]; y[0] = out;
		return out;
	}
	void process(int n, const float* src, float* dst) { for (int i = 0; i < n; ++i) dst[i] = tick(src[i]); }
	void pass_through(int n, const float* src) // :1038-1056
	{
		for (int i = 0; i < n; ++i) { x[1] = x[0]; x[0] = src[i]; y[1] = y[0]; y[0] = src[i]; }
	}
};

float rcp_q_slope(float gain, float slope) { return std::sqrt((gain + (1.0F / gain)) * ((1.0F / slope) - 1.0F) + 2.0F); }
float rcp_q_bandwidth(float fm, float bw)
{
	const float w0 = TAU * fm;
	return 2.0F * std::sinh(std::log(2.0F) / 2.0F * bw * w0 / std::sin(w0));
}

// ---- output device and panning (oalsfxpp.cpp:293-808, 2378-2625) -------------------------------------
enum Spk { S_NONE, S_FL, S_FR, S_FC, S_LFE, S_BL, S_BR, S_BC, S_SL, S_SR };
struct Row { Spk spk; float c[16]; };

const Row DEC_MONO[] = {{S_FC, {1.0F}}};
const Row DEC_STEREO[] = {{S_FL, {5.00000000E-1F, 2.88675135E-1F, 0.0F, 1.19573156E-1F}},
	{S_FR, {5.00000000E-1F, -2.88675135E-1F, 0.0F, 1.19573156E-1F}}};
const Row DEC_QUAD[] = {{S_BL, {3.53553391E-1F, 2.04124145E-1F, 0.0F, -2.04124145E-1F}},
	{S_FL, {3.53553391E-1F, 2.04124145E-1F, 0.0F, 2.04124145E-1F}},
	{S_FR, {3.53553391E-1F, -2.04124145E-1F, 0.0F, 2.04124145E-1F}},
	{S_BR, {3.53553391E-1F, -2.04124145E-1F, 0.0F, -2.04124145E-1F}}};
const Row DEC_51[] = { // side and rear variants share the numbers; the surround pair is renamed below
	{S_SL, {3.33001372E-1F, 1.89085671E-1F, 0.0F, -2.00041334E-1F, -2.12309737E-2F, 0.0F, 0.0F, 0.0F, -1.14573483E-2F}},
	{S_FL, {1.47751298E-1F, 1.28994110E-1F, 0.0F, 1.15190495E-1F, 7.44949143E-2F, 0.0F, 0.0F, 0.0F, -6.47739980E-3F}},
	{S_FC, {7.73595729E-2F, 0.0F, 0.0F, 9.71390298E-2F, 0.0F, 0.0F, 0.0F, 0.0F, 5.18625335E-2F}},
	{S_FR, {1.47751298E-1F, -1.28994110E-1F, 0.0F, 1.15190495E-1F, -7.44949143E-2F, 0.0F, 0.0F, 0.0F, -6.47739980E-3F}},
	{S_SR, {3.33001372E-1F, -1.89085671E-1F, 0.0F, -2.00041334E-1F, 2.12309737E-2F, 0.0F, 0.0F, 0.0F, -1.14573483E-2F}}};
const Row DEC_61[] = {
	{S_SL, {2.04462744E-1F, 2.17178497E-1F, 0.0F, -4.39990188E-2F, -2.60787329E-2F, 0.0F, 0.0F, 0.0F, -6.87238843E-2F}},
	{S_FL, {1.18130342E-1F, 9.34633906E-2F, 0.0F, 1.08553749E-1F, 6.80658795E-2F, 0.0F, 0.0F, 0.0F, 1.08999485E-2F}},
	{S_FC, {7.73595729E-2F, 0.0F, 0.0F, 9.71390298E-2F, 0.0F, 0.0F, 0.0F, 0.0F, 5.18625335E-2F}},
	{S_FR, {1.18130342E-1F, -9.34633906E-2F, 0.0F, 1.08553749E-1F, -6.80658795E-2F, 0.0F, 0.0F, 0.0F, 1.08999485E-2F}},
	{S_SR, {2.04462744E-1F, -2.17178497E-1F, 0.0F, -4.39990188E-2F, 2.60787329E-2F, 0.0F, 0.0F, 0.0F, -6.87238843E-2F}},
	{S_BC, {2.50001688E-1F, 0.0F, 0.0F, -2.50000094E-1F, 0.0F, 0.0F, 0.0F, 0.0F, 6.05133395E-2F}}};
const Row DEC_71[] = {
	{S_BL, {2.04124145E-1F, 1.08880247E-1F, 0.0F, -1.88586120E-1F, -1.29099444E-1F, 0.0F, 0.0F, 0.0F, 7.45355993E-2F, 3.73460789E-2F}},
	{S_SL, {2.04124145E-1F, 2.17760495E-1F, 0.0F, 0.0F, 0.0F, 0.0F, 0.0F, 0.0F, -1.49071198E-1F, -3.73460789E-2F}},
	{S_FL, {2.04124145E-1F, 1.08880247E-1F, 0.0F, 1.88586120E-1F, 1.29099444E-1F, 0.0F, 0.0F, 0.0F, 7.45355993E-2F, 3.73460789E-2F}},
	{S_FR, {2.04124145E-1F, -1.08880247E-1F, 0.0F, 1.88586120E-1F, -1.29099444E-1F, 0.0F, 0.0F, 0.0F, 7.45355993E-2F, -3.73460789E-2F}},
	{S_SR, {2.04124145E-1F, -2.17760495E-1F, 0.0F, 0.0F, 0.0F, 0.0F, 0.0F, 0.0F, -1.49071198E-1F, 3.73460789E-2F}},
	{S_BR, {2.04124145E-1F, -1.08880247E-1F, 0.0F, -1.88586120E-1F, 1.29099444E-1F, 0.0F, 0.0F, 0.0F, 7.45355993E-2F, -3.73460789E-2F}}};

struct Device {
	int format = 0, rate = 0, channels = 0, ncoef = 0;
	Spk order[MAX_CH] = {};
	float dry[MAX_CH][16] = {};
	float foa[MAX_CH][4] = {};
	int map_n = 0;                 // source channel map entries (0 for 5.1 rear: no case, oalsfxpp.cpp:3190-3225)
	float map_angle[MAX_CH] = {};
	bool map_lfe[MAX_CH] = {};

	bool init(int fmt, int r)
	{
		static const Spk O1[] = {S_FC}, O2[] = {S_FL, S_FR}, O4[] = {S_FL, S_FR, S_BL, S_BR},
			O51[] = {S_FL, S_FR, S_FC, S_LFE, S_SL, S_SR}, O51R[] = {S_FL, S_FR, S_FC, S_LFE, S_BL, S_BR},
			O61[] = {S_FL, S_FR, S_FC, S_LFE, S_BC, S_SL, S_SR}, O71[] = {S_FL, S_FR, S_FC, S_LFE, S_BL, S_BR, S_SL, S_SR};
		static const float A1[] = {0}, A2[] = {-30, 30}, A4[] = {-45, 45, -135, 135}, A51[] = {-30, 30, 0, 0, -110, 110},
			A61[] = {-30, 30, 0, 0, 180, -90, 90}, A71[] = {-30, 30, 0, 0, -150, 150, -90, 90};
		const Spk* ord = nullptr; const Row* dec = nullptr; int rows = 0; const float* ang = nullptr;
		bool rear = false;
		switch (fmt) { // set_default_wfx_channel_order + alu_init_renderer, :2420-2570
		case 1: channels = 1; ord = O1; dec = DEC_MONO; rows = 1; ncoef = 1; ang = A1; map_n = 1; break;
		case 2: channels = 2; ord = O2; dec = DEC_STEREO; rows = 2; ncoef = 4; ang = A2; map_n = 2; break;
		case 3: channels = 4; ord = O4; dec = DEC_QUAD; rows = 4; ncoef = 4; ang = A4; map_n = 4; break;
		case 4: channels = 6; ord = O51; dec = DEC_51; rows = 5; ncoef = 9; ang = A51; map_n = 6; break;
		case 5: channels = 6; ord = O51R; dec = DEC_51; rows = 5; ncoef = 9; ang = nullptr; map_n = 0; rear = true; break;
		case 6: channels = 7; ord = O61; dec = DEC_61; rows = 6; ncoef = 9; ang = A61; map_n = 7; break;
		case 7: channels = 8; ord = O71; dec = DEC_71; rows = 6; ncoef = 16; ang = A71; map_n = 8; break;
		default: return false;
		}
		format = fmt; rate = r;
		for (int i = 0; i < channels; ++i) { // set_channel_map, :769-807
			order[i] = ord[i];
			if (ord[i] == S_LFE) continue;
			Spk want = ord[i];
			if (rear && want == S_BL) want = S_SL; // x5_1_rear_panning = side numbers under back names
			if (rear && want == S_BR) want = S_SR;
			for (int j = 0; j < rows; ++j) {
				if (dec[j].spk == want) { for (int k = 0; k < 16; ++k) dry[i][k] = dec[j].c[k]; break; }
			}
			for (int k = 0; k < 4; ++k) foa[i][k] = dry[i][k];
		}
		for (int i = 0; i < map_n; ++i) {
			map_lfe[i] = ord[i] == S_LFE;
			map_angle[i] = ang[i] * (PI / 180.0F);
		}
		return true;
	}
};

void angle_coeffs(float az, float el, float spread, float c[16]) // :483-597
{
	const float dir[3] = {std::sin(az) * std::cos(el), std::sin(el), -std::cos(az) * std::cos(el)};
	const float x = -dir[2], y = -dir[0], z = dir[1];
	c[0] = 1.0F;
	c[1] = 1.732050808F * y; c[2] = 1.732050808F * z; c[3] = 1.732050808F * x;
	c[4] = 3.872983346F * x * y; c[5] = 3.872983346F * y * z; c[6] = 1.118033989F * ((3.0F * z * z) - 1.0F);
	c[7] = 3.872983346F * x * z; c[8] = 1.936491673F * ((x * x) - (y * y));
	c[9] = 2.091650066F * y * ((3.0F * x * x) - (y * y)); c[10] = 10.246950766F * z * x * y;
	c[11] = 1.620185175F * y * ((5.0F * z * z) - 1.0F); c[12] = 1.322875656F * z * ((5.0F * z * z) - 3.0F);
	c[13] = 1.620185175F * x * ((5.0F * z * z) - 1.0F); c[14] = 5.123475383F * z * ((x * x) - (y * y));
	c[15] = 2.091650066F * x * ((x * x) - (3.0F * y * y));
	if (spread > 0.0F) {
		const float ca = std::cos(spread * 0.5F);
		const float scale = std::sqrt(1.0F + (spread / TAU));
		const float z0 = scale, z1 = 0.5F * (ca + 1.0F) * scale, z2 = 0.5F * (ca + 1.0F) * ca * scale,
			z3 = 0.125F * (ca + 1.0F) * ((5.0F * ca * ca) - 1.0F) * scale;
		c[0] *= z0;
		for (int i = 1; i < 4; ++i) c[i] *= z1;
		for (int i = 4; i < 9; ++i) c[i] *= z2;
		for (int i = 9; i < 16; ++i) c[i] *= z3;
	}
}

void pan_gains(const Device& d, const float c[16], float g, float out[MAX_CH]) // compute_panning_gains_mc, :670-696
{
	for (int i = 0; i < MAX_CH; ++i) {
		if (i >= d.channels) { out[i] = 0.0F; continue; }
		float s = 0.0F;
		for (int j = 0; j < d.ncoef; ++j) s += d.dry[i][j] * c[j];
		out[i] = clampv(s, 0.0F, 1.0F) * g;
	}
}

void foa_gains(const Device& d, const float m[4], float g, float out[MAX_CH]) // compute_first_order_gains_mc, :730-755
{
	for (int i = 0; i < MAX_CH; ++i) {
		if (i >= d.channels) { out[i] = 0.0F; continue; }
		float s = 0.0F;
		for (int j = 0; j < 4; ++j) s += d.foa[i][j] * m[j];
		out[i] = clampv(s, 0.0F, 1.0F) * g;
	}
}

void identity_gains(const Device& d, float out[WET_CH][MAX_CH])
{
	for (int i = 0; i < WET_CH; ++i) {
		float m[4] = {0, 0, 0, 0};
		m[i] = 1.0F;
		foa_gains(d, m, 1.0F, out[i]);
	}
}

typedef float Bus[MAX_CH][CHUNK];
typedef float Wet[WET_CH][CHUNK];

// MixHelpers::mix (oalsfxpp.cpp:2752-2798)
void mix_ramped(const float* data, int channels, Bus bus, float* cur, const float* target, int counter, int pos0, int n)
{
	const float delta = (counter > 0 ? 1.0F / static_cast<float>(counter) : 0.0F);
	for (int c = 0; c < channels; ++c) {
		int pos = 0;
		float gain = cur[c];
		const float step = (target[c] - gain) * delta;
		if (std::abs(step) > 1.1920928955078125e-7F) {
			const int size = std::min(n, counter);
			for (; pos < size; ++pos) { bus[c][pos0 + pos] += data[pos] * gain; gain += step; }
			if (pos == counter) gain = target[c];
			cur[c] = gain;
		}
		if (!(std::abs(gain) > SILENCE)) continue;
		for (; pos < n; ++pos) bus[c][pos0 + pos] += data[pos] * gain;
	}
}

// ---- effects ------------------------------------------------------------------------------------
struct Fx {
	virtual ~Fx() {}
	virtual void update_device(const Device&) {}
	virtual void update(const Device&, int type, const Props&) = 0;
	virtual void process(int n, Wet wet, Bus bus, int channels) = 0;
};

void add_scaled(Bus bus, int k, int base, int n, const float* v, float g)
{
	if (!(std::abs(g) > SILENCE)) return;
	for (int i = 0; i < n; ++i) bus[k][base + i] += v[i] * g;
}

struct NullFx : Fx { // :3912-3963
	void update(const Device&, int, const Props&) override {}
	void process(int, Wet, Bus, int) override {}
};

struct ModDelayFx : Fx { // chorus :3972-4277, flanger :5243-5547
	float max_delay;
	std::vector<float> buf[2];
	int len = 0, offset = 0, range = 1, disp = 0, delay = 0, waveform = 1;
	float scale = 0, depth = 0, feedback = 0;
	float gains[2][MAX_CH] = {};
	explicit ModDelayFx(float md) : max_delay(md) {}
	void update_device(const Device& d) override
	{
		len = next_pow2(static_cast<int>(max_delay * 2.0F * d.rate) + 1);
		buf[0].assign(len, 0.0F); buf[1].assign(len, 0.0F);
	}
	void update(const Device& d, int, const Props& p) override
	{
		const PChorus& c = p.chorus;
		const float frequency = static_cast<float>(d.rate);
		waveform = c.waveform; feedback = c.feedback;
		delay = static_cast<int>(c.delay * frequency);
		depth = c.depth * delay;
		float co[16];
		angle_coeffs(-PI_2, 0.0F, 0.0F, co); pan_gains(d, co, 1.0F, gains[0]);
		angle_coeffs(PI_2, 0.0F, 0.0F, co); pan_gains(d, co, 1.0F, gains[1]);
		if (!(c.rate > 0.0F)) { scale = 0.0F; range = 1; disp = 0; return; }
		range = static_cast<int>(frequency / c.rate + 0.5F);
		scale = (waveform == 1 ? 4.0F / range : TAU / range);
		disp = (c.phase >= 0 ? static_cast<int>(range * (c.phase / 360.0F)) : static_cast<int>(range * ((360 + c.phase) / 360.0F)));
	}
	int lfo(int ph) const
	{
		if (waveform == 1) return static_cast<int>((1.0F - std::abs(2.0F - (scale * ph))) * depth) + delay;
		return static_cast<int>(std::sin(scale * ph) * depth) + delay;
	}
	void process(int n, Wet wet, Bus bus, int channels) override
	{
		const int mask = len - 1;
		for (int base = 0; base < n;) {
			float t[2][128];
			const int todo = std::min(128, n - base);
			int ph[2] = {offset % range, (offset + disp) % range};
			for (int i = 0; i < todo; ++i) {
				for (int s = 0; s < 2; ++s) {
					const int d = lfo(ph[s]);
					ph[s] = (ph[s] + 1) % range;
					buf[s][offset & mask] = wet[0][base + i];
					t[s][i] = buf[s][(offset - d) & mask] * feedback;
					buf[s][offset & mask] += t[s][i];
				}
				++offset;
			}
			for (int k = 0; k < channels; ++k) {
				add_scaled(bus, k, base, todo, t[0], gains[0][k]);
				add_scaled(bus, k, base, todo, t[1], gains[1][k]);
			}
			base += todo;
		}
	}
};

struct CompressorFx : Fx { // :4286-4468
	float gains[WET_CH][MAX_CH] = {};
	bool enabled = true;
	float attack = 0, release = 0, control = 1.0F;
	void update_device(const Device& d) override { attack = 1.0F / (d.rate * 0.2F); release = 1.0F / (d.rate * 0.4F); }
	void update(const Device& d, int, const Props& p) override { enabled = p.compressor.on_off; identity_gains(d, gains); }
	void process(int n, Wet wet, Bus bus, int channels) override
	{
		static thread_local float tmp[WET_CH][CHUNK];
		for (int i = 0; i < n; ++i) {
			float amp = 1.0F;
			if (enabled) {
				amp = std::abs(wet[0][i]);
				amp = std::max(amp + std::abs(wet[1][i]), std::max(amp + std::abs(wet[2][i]), amp + std::abs(wet[3][i])));
			}
			if (amp > control) control = std::min(control + attack, amp);
			else if (amp < control) control = std::max(control - release, amp);
			const float out = 1.0F / clampv(control, 0.5F, 2.0F);
			for (int j = 0; j < WET_CH; ++j) tmp[j][i] = wet[j][i] * out;
		}
		for (int j = 0; j < WET_CH; ++j)
			for (int k = 0; k < channels; ++k) add_scaled(bus, k, 0, n, tmp[j], gains[j][k]);
	}
};

struct DedicatedFx : Fx { // :4477-4581; get_channel_index is always -1 (:2577-2578)
	float gains[MAX_CH] = {};
	void update(const Device& d, int type, const Props& p) override
	{
		for (float& g : gains) g = 0.0F;
		if (type == DIALOG) {
			float co[16];
			angle_coeffs(0.0F, 0.0F, 0.0F, co);
			pan_gains(d, co, p.dedicated.gain, gains);
		}
	}
	void process(int n, Wet wet, Bus bus, int channels) override
	{
		for (int k = 0; k < channels; ++k) add_scaled(bus, k, 0, n, wet[0], gains[k]);
	}
};

struct DistortionFx : Fx { // :4590-4762
	float gains[MAX_CH] = {};
	Filter lp, bp;
	float atten = 0, edge = 0;
	void update(const Device& d, int, const Props& p) override
	{
		const PDistortion& q = p.distortion;
		const float frequency = static_cast<float>(d.rate);
		atten = q.gain;
		float e = std::sin(q.edge * PI_2);
		e = std::min(e, 0.99F);
		edge = 2.0F * e / (1.0F - e);
		float cutoff = q.lp_cutoff;
		float bw = (cutoff / 2.0F) / (cutoff * 0.67F);
		lp.set(LOW_PASS, 1.0F, cutoff / (frequency * 4.0F), rcp_q_bandwidth(cutoff / (frequency * 4.0F), bw));
		cutoff = q.eq_center;
		bw = q.eq_bandwidth / (cutoff * 0.67F);
		bp.set(BAND_PASS, 1.0F, cutoff / (frequency * 4.0F), rcp_q_bandwidth(cutoff / (frequency * 4.0F), bw));
		for (int i = 0; i < MAX_CH; ++i) gains[i] = (i < d.channels ? d.dry[i][0] * 1.414213562F * 1.0F : 0.0F); // :616-626
	}
	void process(int n, Wet wet, Bus bus, int channels) override
	{
		static thread_local float up[2][CHUNK * 4];
		for (int i = 0; i < n; ++i) {
			up[0][i * 4] = wet[0][i] * 4.0F;
			up[0][i * 4 + 1] = up[0][i * 4 + 2] = up[0][i * 4 + 3] = 0.0F;
		}
		lp.process(n * 4, up[0], up[1]);
		for (int i = 0; i < n * 4; ++i) {
			float s = up[1][i];
			s = (1.0F + edge) * s / (1.0F + (edge * std::abs(s)));
			s = (1.0F + edge) * s / (1.0F + (edge * std::abs(s))) * -1.0F;
			s = (1.0F + edge) * s / (1.0F + (edge * std::abs(s)));
			up[0][i] = s;
		}
		bp.process(n * 4, up[0], up[1]);
		for (int k = 0; k < channels; ++k) {
			const float g = gains[k] * atten;
			if (!(std::abs(g) > SILENCE)) continue;
			for (int i = 0; i < n; ++i) bus[k][i] += g * up[1][i * 4];
		}
	}
};

struct EchoFx : Fx { // :4771-4985
	std::vector<float> buf;
	int len = 0, tap1 = 0, tap2 = 0, offset = 0;
	float gains[2][MAX_CH] = {};
	float feed = 0;
	Filter filt;
	void update_device(const Device& d) override
	{
		int m = static_cast<int>(0.207F * d.rate) + 1;
		m += static_cast<int>(0.404F * d.rate) + 1;
		len = next_pow2(m);
		buf.assign(len, 0.0F);
	}
	void update(const Device& d, int, const Props& p) override
	{
		const PEcho& e = p.echo;
		tap1 = static_cast<int>(e.delay * d.rate) + 1;
		tap2 = static_cast<int>(e.lr_delay * d.rate);
		tap2 += tap1;
		float spread = e.spread;
		const float lrpan = (spread < 0.0F ? -1.0F : 1.0F);
		spread = std::asin(1.0F - std::abs(spread)) * 4.0F;
		feed = e.feedback;
		const float g = std::max(1.0F - e.damping, 0.0625F);
		filt.set(HIGH_SHELF, g, 5000.0F / d.rate, rcp_q_slope(g, 1.0F));
		float co[16];
		angle_coeffs(-PI_2 * lrpan, 0.0F, spread, co); pan_gains(d, co, 1.0F, gains[0]);
		angle_coeffs(PI_2 * lrpan, 0.0F, spread, co); pan_gains(d, co, 1.0F, gains[1]);
	}
	void process(int n, Wet wet, Bus bus, int channels) override
	{
		static thread_local float t[2][CHUNK];
		const int mask = len - 1;
		for (int i = 0; i < n; ++i) {
			t[0][i] = buf[(offset - tap1) & mask];
			t[1][i] = buf[(offset - tap2) & mask];
			const float out = filt.tick(t[1][i] + wet[0][i]);
			buf[offset & mask] = out * feed;
			++offset;
		}
		for (int k = 0; k < channels; ++k) {
			// per sample the reference adds tap 1 then tap 2 (128-sample sub-blocks, :4934-4954)
			const bool a0 = std::abs(gains[0][k]) > SILENCE, a1 = std::abs(gains[1][k]) > SILENCE;
			for (int i = 0; i < n; ++i) {
				if (a0) bus[k][i] += t[0][i] * gains[0][k];
				if (a1) bus[k][i] += t[1][i] * gains[1][k];
			}
		}
	}
};

struct EqualizerFx : Fx { // :5034-5232
	float gains[WET_CH][MAX_CH] = {};
	Filter f[4][WET_CH];
	void update(const Device& d, int, const Props& p) override
	{
		const PEqualizer& q = p.equalizer;
		const float frequency = static_cast<float>(d.rate);
		identity_gains(d, gains);
		float g = std::max(std::sqrt(q.low_gain), 0.0625F);
		float fm = q.low_cutoff / frequency;
		f[0][0].set(LOW_SHELF, g, fm, rcp_q_slope(g, 0.75F));
		g = std::max(q.mid1_gain, 0.0625F); fm = q.mid1_center / frequency;
		f[1][0].set(PEAKING, g, fm, rcp_q_bandwidth(fm, q.mid1_width));
		g = std::max(q.mid2_gain, 0.0625F); fm = q.mid2_center / frequency;
		f[2][0].set(PEAKING, g, fm, rcp_q_bandwidth(fm, q.mid2_width));
		g = std::max(std::sqrt(q.high_gain), 0.0625F); fm = q.high_cutoff / frequency;
		f[3][0].set(HIGH_SHELF, g, fm, rcp_q_slope(g, 0.75F));
		for (int b = 0; b < 4; ++b) for (int c = 1; c < WET_CH; ++c) f[b][c].copy_coeffs(f[b][0]);
	}
	void process(int n, Wet wet, Bus bus, int channels) override
	{
		static thread_local float a[CHUNK], b[CHUNK];
		for (int c = 0; c < WET_CH; ++c) {
			f[0][c].process(n, wet[c], a);
			f[1][c].process(n, a, b);
			f[2][c].process(n, b, a);
			f[3][c].process(n, a, b);
			for (int k = 0; k < channels; ++k) add_scaled(bus, k, 0, n, b, gains[c][k]);
		}
	}
};

struct RingModFx : Fx { // :5556-5785
	int index = 0, step = 1, waveform = 0;
	float gains[WET_CH][MAX_CH] = {};
	Filter f[WET_CH];
	void update(const Device& d, int, const Props& p) override
	{
		const PRingMod& r = p.ringmod;
		waveform = r.waveform;
		step = static_cast<int>(r.frequency * (1 << 24) / d.rate);
		if (step == 0) step = 1;
		const float cw = std::cos(TAU * r.hp_cutoff / d.rate);
		const float a = (2.0F - cw) - std::sqrt(std::pow(2.0F - cw, 2.0F) - 1.0F);
		for (int i = 0; i < WET_CH; ++i) { f[i].b0 = a; f[i].b1 = -a; f[i].b2 = 0.0F; f[i].a1 = -a; f[i].a2 = 0.0F; }
		identity_gains(d, gains);
	}
	float osc(int idx) const
	{
		if (waveform == 0) return std::sin(idx * (TAU / (1 << 24)) - PI) * 0.5F + 0.5F;
		if (waveform == 1) return static_cast<float>(idx) / (1 << 24);
		return static_cast<float>((idx >> 23) & 1);
	}
	void process(int n, Wet wet, Bus bus, int channels) override
	{
		static thread_local float t[CHUNK];
		for (int j = 0; j < WET_CH; ++j) {
			int idx = index;
			for (int i = 0; i < n; ++i) {
				idx = (idx + step) & 0xFFFFFF;
				t[i] = f[j].tick(wet[j][i]) * osc(idx);
			}
			for (int k = 0; k < channels; ++k) add_scaled(bus, k, 0, n, t, gains[j][k]);
		}
		for (int i = 0; i < n; ++i) index = (index + step) & 0xFFFFFF;
	}
};

// Reverb / EAX reverb (oalsfxpp.cpp:5799-7904).
struct Line4 { // DelayLineI: interleaved 4-float frames, power-of-two length
	int mask = 0;
	std::vector<float> d;
	void init(float seconds, int rate, int extra)
	{
		const int n = next_pow2(static_cast<int>(std::ceil(seconds * rate)) + extra); // :6538-6552
		mask = n - 1;
		d.assign(static_cast<size_t>(n) * 4, 0.0F);
	}
	float get(int off, int c) const { return d[static_cast<size_t>(off & mask) * 4 + c]; }
	void put(int off, int c, float v) { d[static_cast<size_t>(off & mask) * 4 + c] = v; }
};

const float ETAP[4] = {0.000000E+0F, 1.010676E-3F, 2.126553E-3F, 3.358580E-3F};
const float EAP[4] = {4.854840E-4F, 5.360178E-4F, 5.918117E-4F, 6.534130E-4F};
const float ELINE[4] = {2.992520E-3F, 5.456575E-3F, 7.688329E-3F, 9.709681E-3F};
const float LAP[4] = {8.091400E-4F, 1.019453E-3F, 1.407968E-3F, 1.618280E-3F};
const float LLINE[4] = {9.709681E-3F, 1.223343E-2F, 1.689561E-2F, 1.941936E-2F};

struct ReverbFx : Fx {
	bool eax = false;
	Filter lp[4], hp[4];
	Line4 main, eap, eline, lap, lline;
	int etap[4][2] = {}, ltap[4][2] = {}, eapo[4][2] = {}, eoff[4][2] = {}, lapo[4][2] = {}, loff[4][2] = {};
	float etapc[4] = {}, ecoef[4] = {};
	int feed_tap = 0;
	float apc = 0, mx = 0, my = 0;
	int mod_index = 0, mod_range = 1;
	float mod_depth = 0, mod_coeff = 0, mod_filter = 0;
	float density_gain = 0;
	float t60lf[4][3] = {}, t60hf[4][3] = {}, t60mid[4] = {}, t60s[4][2][2] = {};
	float ecur[4][MAX_CH] = {}, epan[4][MAX_CH] = {}, lcur[4][MAX_CH] = {}, lpan[4][MAX_CH] = {};
	int fade_count = 0, offset = 0;

	void update_device(const Device& d) override // :5928-5950, alloc_lines :6556-6598
	{
		const int f = d.rate;
		const float mult = 1.0F + 9.0F;
		float length = 0.3F + (ETAP[3] * mult) + 0.1F + ((LLINE[3] - LLINE[0]) * 0.25F * mult);
		main.init(length, f, 256);
		eap.init(EAP[3] * mult, f, 0);
		eline.init(ELINE[3] * mult, f, 0);
		lap.init(LAP[3] * mult, f, 0);
		length = std::max(0.25F, LLINE[3] * mult) + (4.0F * (1.0F / 4096.0F) / 2.0F);
		lline.init(length, f, 0);
		mod_coeff = std::pow(0.048F, 100000.0F / f);
		feed_tap = static_cast<int>((0.3F + (ETAP[3] * mult)) * f);
	}

	static float decay_coeff(float length, float t) { return std::pow(0.001F, length / t); }
	static void unity(float c[3]) { c[0] = 1.0F; c[1] = 0.0F; c[2] = 0.0F; }
	static void hpass(float gain, float w, float c[3]) // :6717-6738
	{
		if (gain >= 1.0F) { unity(c); return; }
		const float g = std::max(0.001F, gain), g2 = g * g, cw = std::cos(w);
		const float p = g / ((g * cw) + std::sqrt((cw - 1.0F) * ((g2 * cw) + g2 - 2.0F)));
		c[0] = p; c[1] = -p; c[2] = p;
	}
	static void lpass(float gain, float w, float c[3]) // :6762-6786
	{
		if (gain >= 1.0F) { unity(c); return; }
		const float g = std::max(0.001F, gain), g2 = g * g, cw = std::cos(w);
		const float a = (1.0F - (g2 * cw) - std::sqrt((2.0F * g2 * (1.0F - cw)) - (g2 * g2 * (1.0F - (cw * cw))))) / (1.0F - g2);
		c[0] = 1.0F - a; c[1] = 0.0F; c[2] = a;
	}
	static void shelf(bool high, float gain, float w, float c[3]) // :6832-6857 (low), :6904-6927 (high)
	{
		if (gain >= 1.0F) { unity(c); return; }
		const float g = std::max(0.001F, gain);
		float p;
		if (high) {
			p = std::sin((0.5F * w) - (0.25F * PI)) / std::sin((0.5F * w) + (0.25F * PI));
		} else {
			const float rw = PI - w;
			p = std::sin((0.5F * rw) - (0.25F * PI)) / std::sin((0.5F * rw) + (0.25F * PI));
		}
		const float n = (g + 1.0F) / (g - 1.0F);
		const float alpha = n + std::sqrt((n * n) - 1.0F);
		const float beta0 = (1.0F + g + (1.0F - g) * alpha) / 2.0F;
		const float beta1 = (1.0F - g + (1.0F + g) * alpha) / 2.0F;
		c[0] = (beta0 + (p * beta1)) / (1.0F + (p * alpha));
		if (high) {
			c[1] = (beta1 + (p * beta0)) / (1.0F + (p * alpha));
			c[2] = -(p + alpha) / (1.0F + (p * alpha));
		} else {
			c[1] = -(beta1 + (p * beta0)) / (1.0F + (p * alpha));
			c[2] = (p + alpha) / (1.0F + (p * alpha));
		}
	}
	static void t60(float length, float lft, float mft, float hft, float lfw, float hfw, float lf[3], float hf[3], float& mid)
	{ // calc_t60_damping_coeffs, :6934-7010
		const float lg = decay_coeff(length, lft), mg = decay_coeff(length, mft), hg = decay_coeff(length, hft);
		if (lg < mg) {
			if (mg < hg) { shelf(false, mg / hg, hfw, lf); hpass(lg / mg, lfw, hf); mid = hg; }
			else if (mg > hg) { hpass(lg / mg, lfw, lf); lpass(hg / mg, hfw, hf); mid = mg; }
			else { unity(lf); hpass(lg / mg, lfw, hf); mid = mg; }
		} else if (lg > mg) {
			if (mg < hg) {
				const float h = mg / lg, l = mg / hg;
				shelf(true, h, lfw, lf); shelf(false, l, hfw, hf);
				mid = std::max(lg, hg) / std::max(h, l);
			} else if (mg > hg) { shelf(true, mg / lg, lfw, lf); lpass(hg / mg, hfw, hf); mid = lg; }
			else { unity(lf); shelf(true, mg / lg, lfw, hf); mid = lg; }
		} else {
			unity(lf);
			if (mg < hg) { shelf(false, mg / hg, hfw, hf); mid = hg; }
			else if (mg > hg) { lpass(hg / mg, hfw, hf); mid = mg; }
			else { unity(hf); mid = mg; }
		}
	}

	struct M4 { float m[4][4]; };
	static M4 mul(const M4& a, const M4& b, bool transposed)
	{
		M4 r;
		for (int col = 0; col < 4; ++col)
			for (int row = 0; row < 4; ++row) {
				const float v = (a.m[row][0] * b.m[0][col]) + (a.m[row][1] * b.m[1][col]) + (a.m[row][2] * b.m[2][col]) +
					(a.m[row][3] * b.m[3][col]);
				if (transposed) r.m[col][row] = v; else r.m[row][col] = v;
			}
		return r;
	}
	static M4 pan_transform(const float v[3]) // get_transform_from_vector, :7229-7281
	{
		const float length = std::sqrt((v[0] * v[0]) + (v[1] * v[1]) + (v[2] * v[2]));
		const float sa = std::sin(std::min(length, 1.0F) * (PI / 4.0F));
		const M4 zf = {{{1.0F / (1.0F + sa), 0.0F, 0.0F, (sa / (1.0F + sa)) / 1.732050808F},
			{0.0F, std::sqrt((1.0F - sa) / (1.0F + sa)), 0.0F, 0.0F},
			{0.0F, 0.0F, std::sqrt((1.0F - sa) / (1.0F + sa)), 0.0F},
			{(sa / (1.0F + sa)) * 1.732050808F, 0.0F, 0.0F, 1.0F / (1.0F + sa)}}};
		float a = std::atan2(v[1], std::sqrt((v[0] * v[0]) + (v[2] * v[2])));
		const M4 xr = {{{1.0F, 0.0F, 0.0F, 0.0F}, {0.0F, 1.0F, 0.0F, 0.0F}, {0.0F, 0.0F, std::cos(a), std::sin(a)},
			{0.0F, 0.0F, -std::sin(a), std::cos(a)}}};
		a = std::atan2(-v[0], v[2]);
		const M4 yr = {{{1.0F, 0.0F, 0.0F, 0.0F}, {0.0F, std::cos(a), 0.0F, std::sin(a)}, {0.0F, 0.0F, 1.0F, 0.0F},
			{0.0F, -std::sin(a), 0.0F, std::cos(a)}}};
		return mul(yr, mul(xr, zf, false), false);
	}

	void update(const Device& d, int type, const Props& props) override // :5952-6076
	{
		const PReverb& p = props.reverb;
		eax = type == EAXREVERB;
		const int f = d.rate;
		const float hf_scale = p.hf_ref / f;
		const float ghf = std::max(p.gain_hf, 0.001F);
		lp[0].set(HIGH_SHELF, ghf, hf_scale, rcp_q_slope(ghf, 1.0F));
		const float lf_scale = p.lf_ref / f;
		const float glf = std::max(p.gain_lf, 0.001F);
		hp[0].set(LOW_SHELF, glf, lf_scale, rcp_q_slope(glf, 1.0F));
		for (int i = 1; i < 4; ++i) { lp[i].copy_coeffs(lp[0]); hp[i].copy_coeffs(hp[0]); }

		const float mult = 1.0F + (p.density * 9.0F);
		for (int i = 0; i < 4; ++i) { // update_delay_line, :7046-7076
			float length = p.refl_delay + (ETAP[i] * mult);
			etap[i][1] = static_cast<int>(length * f);
			length = ETAP[i] * mult;
			etapc[i] = decay_coeff(length, p.decay_time);
			length = p.late_delay + (LLINE[i] - LLINE[0]) * 0.25F * mult;
			ltap[i][1] = feed_tap + static_cast<int>(length * f);
		}
		apc = std::sqrt(0.5F) * std::pow(p.diffusion, 2.0F);
		for (int i = 0; i < 4; ++i) { // update_early_lines, :7078-7101
			float length = EAP[i] * mult;
			eapo[i][1] = static_cast<int>(length * f);
			length = ELINE[i] * mult;
			eoff[i][1] = static_cast<int>(length * f);
			ecoef[i] = decay_coeff(length, p.decay_time);
		}
		{ // calc_matrix_coeffs, :6645-6659
			const float n = std::sqrt(3.0F);
			const float t = p.diffusion * std::atan(n);
			mx = std::cos(t);
			my = std::sin(t) / n;
		}
		float hf_ratio = p.decay_hf_ratio;
		if (p.hf_limit && p.air_gain_hf < 1.0F) { // calc_limited_hf_ratio, :6663-6679
			const float limit = 1.0F / ((std::log10(p.air_gain_hf) * p.decay_time / std::log10(0.001F)) * 343.3F);
			hf_ratio = clampv(limit, 0.1F, hf_ratio);
		}
		const float lft = clampv(p.decay_time * p.decay_lf_ratio, 0.1F, 20.0F);
		const float hft = clampv(p.decay_time * hf_ratio, 0.1F, 20.0F);
		{ // update_modulator, :7014-7043
			const int range = std::max(static_cast<int>(p.mod_time * f), 1);
			mod_index = static_cast<int>(mod_index * static_cast<int64_t>(range) / mod_range);
			mod_range = range;
			mod_depth = p.mod_depth * (1.0F / 4096.0F) * p.mod_time / 2.0F * f;
		}
		{ // update_late_lines, :7103-7187
			const float lfw = TAU * lf_scale, hfw = TAU * hf_scale;
			float length = (LLINE[0] + LLINE[1] + LLINE[2] + LLINE[3]) / 4.0F * mult;
			length = length + ((p.echo_time - length) * p.echo_depth);
			length += (LAP[0] + LAP[1] + LAP[2] + LAP[3]) / 4.0F * mult;
			const float bw[3] = {lfw, hfw - lfw, TAU - hfw};
			const float a = decay_coeff(length, ((bw[0] * lft) + (bw[1] * p.decay_time) + (bw[2] * hft)) / TAU);
			density_gain = std::sqrt(1.0F - (a * a));
			for (int i = 0; i < 4; ++i) {
				length = LAP[i] * mult;
				lapo[i][1] = static_cast<int>(length * f);
				const float ll = LLINE[i] * mult;
				length = ll + ((p.echo_time - ll) * p.echo_depth);
				loff[i][1] = static_cast<int>(length * f);
				const float avg = (LAP[0] + LAP[1] + LAP[2] + LAP[3]) / 4.0F;
				length += (LAP[i] + ((avg - LAP[i]) * p.diffusion)) * mult;
				t60(length, lft, p.decay_time, hft, lfw, hfw, t60lf[i], t60hf[i], t60mid[i]);
			}
		}
		{ // update_3d_panning, :7306-7350
			static const M4 a2b = {{{0.866025403785F, 0.866025403785F, 0.866025403785F, 0.866025403785F},
				{0.866025403785F, -0.866025403785F, 0.866025403785F, -0.866025403785F},
				{0.866025403785F, -0.866025403785F, -0.866025403785F, 0.866025403785F},
				{0.866025403785F, 0.866025403785F, -0.866025403785F, -0.866025403785F}}};
			M4 tr = mul(pan_transform(p.refl_pan), a2b, true);
			for (int i = 0; i < 4; ++i) foa_gains(d, tr.m[i], p.gain * p.refl_gain, epan[i]);
			tr = mul(pan_transform(p.late_pan), a2b, true);
			for (int i = 0; i < 4; ++i) foa_gains(d, tr.m[i], p.gain * p.late_gain, lpan[i]);
		}
		for (int i = 0; i < 4; ++i) { // :6061-6075
			if (etap[i][1] != etap[i][0] || eapo[i][1] != eapo[i][0] || eoff[i][1] != eoff[i][0] || ltap[i][1] != ltap[i][0] ||
				lapo[i][1] != lapo[i][0] || loff[i][1] != loff[i][0]) {
				fade_count = 0;
				break;
			}
		}
	}

	static float rd(const Line4& l, int o0, int o1, int c, float mu, bool faded) // :7358-7406
	{
		if (!faded) return l.get(o0, c);
		const float a = l.get(o0, c), b = l.get(o1, c);
		return a + ((b - a) * mu);
	}
	static void scatter(float v[4], float x, float y) // vector_partial_scatter, :7510-7521
	{
		const float f[4] = {v[0], v[1], v[2], v[3]};
		v[0] = (x * f[0]) + (y * (f[1] + -f[2] + f[3]));
		v[1] = (x * f[1]) + (y * (-f[0] + f[2] + f[3]));
		v[2] = (x * f[2]) + (y * (f[0] + -f[1] + f[3]));
		v[3] = (x * f[3]) + (y * (-f[0] + -f[1] + -f[2]));
	}
	void allpass(Line4& l, int offs[4][2], float v[4], int off, float mu, bool faded) const // vector_allpass_x, :7533-7562
	{
		float f[4];
		for (int i = 0; i < 4; ++i) {
			const float in = v[i];
			v[i] = rd(l, off - offs[i][0], off - offs[i][1], i, mu, faded) - (apc * in);
			f[i] = in + (apc * v[i]);
		}
		scatter(f, mx, my);
		for (int i = 0; i < 4; ++i) l.put(off, i, f[i]);
	}

	void process(int n, Wet wet, Bus bus, int channels) override // do_process, :6078-6170
	{
		static thread_local float afmt[4][256], early[4][256], late[4][256];
		float fade = static_cast<float>(fade_count) / 128;
		for (int base = 0; base < n;) {
			int todo = std::min(n - base, 256);
			if (128 - fade_count > 0) todo = std::min(todo, 128 - fade_count);
			static const float q = 0.288675134595F;
			static const float b2a[4][4] = {{q, q, q, q}, {q, -q, -q, q}, {q, q, -q, -q}, {q, -q, q, -q}};
			for (int c = 0; c < 4; ++c) {
				for (int i = 0; i < todo; ++i) afmt[c][i] = 0.0F;
				for (int k = 0; k < 4; ++k) // mix_row, :2728-2750
					for (int i = 0; i < todo; ++i) afmt[c][i] += wet[k][base + i] * b2a[c][k];
			}
			// (eax_)verb_pass, :7814-7903
			for (int c = 0; c < 4; ++c) {
				for (int i = 0; i < todo; ++i) {
					float v = lp[c].tick(afmt[c][i]);
					if (eax) v = hp[c].tick(v);
					main.put(offset + i, c, v);
				}
			}
			const bool faded = fade < 1.0F;
			{ // early_reflection_x, :7625-7672
				float mu = fade;
				int off = offset;
				for (int i = 0; i < todo; ++i) {
					float f[4];
					for (int j = 0; j < 4; ++j) f[j] = rd(main, off - etap[j][0], off - etap[j][1], j, mu, faded) * etapc[j];
					allpass(eap, eapo, f, off, mu, faded);
					for (int j = 0; j < 4; ++j) eline.put(off, j, f[3 - j]);
					for (int j = 0; j < 4; ++j) f[j] += rd(eline, off - eoff[j][0], off - eoff[j][1], j, mu, faded) * ecoef[j];
					for (int j = 0; j < 4; ++j) early[j][i] = f[j];
					float r[4] = {f[3], f[2], f[1], f[0]};
					scatter(r, mx, my);
					for (int j = 0; j < 4; ++j) main.put(off - feed_tap, j, r[j]);
					++off;
					mu += 1.0F / 128;
				}
			}
			{ // late_reverb_x, :7735-7794 with calc_modulation_delays, :7443-7470
				int md[256];
				for (int i = 0; i < todo; ++i) {
					const float sinus = std::sin(TAU * mod_index / mod_range);
					mod_index = (mod_index + 1) % mod_range;
					mod_filter = mod_filter + ((mod_depth - mod_filter) * mod_coeff);
					md[i] = static_cast<int>(std::lround(mod_filter * sinus));
				}
				float mu = fade;
				int off = offset;
				for (int i = 0; i < todo; ++i) {
					float f[4];
					for (int j = 0; j < 4; ++j) f[j] = rd(main, off - ltap[j][0], off - ltap[j][1], j, mu, faded) * density_gain;
					const int dl = off - md[i];
					for (int j = 0; j < 4; ++j) f[j] += rd(lline, dl - loff[j][0], dl - loff[j][1], j, mu, faded);
					for (int j = 0; j < 4; ++j) { // late_t60_filter, :7691-7719
						const float o1 = (t60lf[j][0] * f[j]) + (t60lf[j][1] * t60s[j][0][0]) + (t60lf[j][2] * t60s[j][0][1]);
						t60s[j][0][0] = f[j]; t60s[j][0][1] = o1;
						const float o2 = (t60hf[j][0] * o1) + (t60hf[j][1] * t60s[j][1][0]) + (t60hf[j][2] * t60s[j][1][1]);
						t60s[j][1][0] = o1; t60s[j][1][1] = o2;
						f[j] = t60mid[j] * o2;
					}
					allpass(lap, lapo, f, off, mu, faded);
					for (int j = 0; j < 4; ++j) late[j][i] = f[j];
					float r[4] = {f[3], f[2], f[1], f[0]};
					scatter(r, mx, my);
					for (int j = 0; j < 4; ++j) lline.put(off, j, r[j]);
					++off;
					mu += 1.0F / 128;
				}
			}
			if (faded) fade = std::min(1.0F, fade + (todo * (1.0F / 128)));
			offset += todo;
			if (fade_count < 128) { // :6118-6138
				fade_count += todo;
				if (fade_count >= 128) {
					fade_count = 128;
					fade = 1.0F;
					for (int c = 0; c < 4; ++c) {
						etap[c][0] = etap[c][1]; eapo[c][0] = eapo[c][1]; eoff[c][0] = eoff[c][1];
						ltap[c][0] = ltap[c][1]; lapo[c][0] = lapo[c][1]; loff[c][0] = loff[c][1];
					}
				}
			}
			for (int c = 0; c < 4; ++c) mix_ramped(early[c], channels, bus, ecur[c], epan[c], n - base, base, todo);
			for (int c = 0; c < 4; ++c) mix_ramped(late[c], channels, bus, lcur[c], lpan[c], n - base, base, todo);
			base += todo;
		}
	}
};

Fx* make_fx(int type)
{
	switch (type) {
	case CHORUS: return new ModDelayFx(0.016F);
	case FLANGER: return new ModDelayFx(0.004F);
	case COMPRESSOR: return new CompressorFx;
	case DIALOG: case LFE_FX: return new DedicatedFx;
	case DISTORTION: return new DistortionFx;
	case ECHO: return new EchoFx;
	case EQUALIZER: return new EqualizerFx;
	case RINGMOD: return new RingModFx;
	case REVERB: case EAXREVERB: return new ReverbFx;
	default: return new NullFx;
	}
}

// ---- one instance = one reference Api (oalsfxpp.cpp:2820-3432, 3468-3903) -----------------------------
struct SendState {
	Send3 props, deferred;
	int filter_type = 0;
	Filter lpf[MAX_CH], hpf[MAX_CH];
	float gains[MAX_CH][MAX_CH] = {};
	bool enabled = false;
};

struct Slot {
	Effect deferred, active;
	std::unique_ptr<Fx> fx;
	bool changed = false;
	Wet wet;
};

struct Instance {
	Device dev;
	int nfx = 0;
	Slot slots[MAX_FX];
	SendState direct, aux[MAX_FX];
	bool source_changed = true;
	Bus bus;

	bool init(int fmt, int rate, int n)
	{
		if (!dev.init(fmt, rate)) return false;
		if (rate < 8000) return false;
		if (n <= 0 || n > MAX_FX) return false;
		nfx = n;
		for (int i = 0; i < nfx; ++i) {
			std::memset(&slots[i].deferred, 0, sizeof(Effect));
			std::memset(&slots[i].active, 0, sizeof(Effect));
			slots[i].fx.reset(make_fx(NUL));
			slots[i].changed = true;
			aux[i].props = aux[i].deferred = Send3{1.0F, 1.0F, 1.0F};
		}
		direct.props = direct.deferred = Send3{1.0F, 1.0F, 1.0F};
		return true;
	}

	void apply() // Api::apply_changes, :3738-3783
	{
		for (int i = 0; i < nfx; ++i) {
			Slot& s = slots[i];
			normalize(s.deferred);
			if (!effects_equal(s.deferred, s.active)) {
				if (s.active.type != s.deferred.type) { // EffectSlot::set_effect, :2688-2709
					s.fx.reset(make_fx(s.deferred.type));
					s.fx->update_device(dev);
				}
				s.active = s.deferred;
				s.changed = true;
			}
		}
		normalize(direct.deferred);
		if (!sends_equal(direct.deferred, direct.props)) { source_changed = true; direct.props = direct.deferred; }
		for (int i = 0; i < nfx; ++i) {
			normalize(aux[i].deferred);
			if (!sends_equal(aux[i].props, aux[i].deferred)) source_changed = true;
		}
	}

	void send_filters(SendState& s, float ghf, float glf) const // :3271-3343
	{
		const float hf_scale = 250.0F / dev.rate, lf_scale = 5000.0F / dev.rate; // swapped constants, :3271-3272
		ghf = std::max(ghf, 0.001F); glf = std::max(glf, 0.001F);
		s.filter_type = (ghf != 1.0F ? 1 : 0) | (glf != 1.0F ? 2 : 0);
		s.lpf[0].set(HIGH_SHELF, ghf, hf_scale, rcp_q_slope(ghf, 1.0F));
		s.hpf[0].set(LOW_SHELF, glf, lf_scale, rcp_q_slope(glf, 1.0F));
		for (int c = 1; c < dev.map_n; ++c) { s.lpf[c].copy_coeffs(s.lpf[0]); s.hpf[c].copy_coeffs(s.hpf[0]); }
	}

	void refresh() // update_context_sources, :3397-3412
	{
		bool updated = false;
		for (int i = 0; i < nfx; ++i) {
			if (slots[i].changed) {
				slots[i].changed = false;
				slots[i].fx->update(dev, slots[i].active.type, slots[i].active.props);
				updated = true;
			}
		}
		if (source_changed) { source_changed = false; updated = true; }
		if (!updated) return;
		// calc_non_attn_source_params + calc_panning_and_filters, :3348-3395, :3172-3346
		for (int i = 0; i < nfx; ++i) aux[i].enabled = slots[i].active.type != NUL;
		const float dry = std::min(direct.props.gain, MAX_GAIN);
		for (int c = 0; c < dev.map_n; ++c) {
			for (int k = 0; k < MAX_CH; ++k) {
				direct.gains[c][k] = 0.0F;
				for (int i = 0; i < nfx; ++i) aux[i].gains[c][k] = 0.0F;
			}
			if (dev.map_lfe[c]) continue;
			float co[16];
			angle_coeffs(dev.map_angle[c], 0.0F, 0.0F, co);
			pan_gains(dev, co, dry, direct.gains[c]);
			for (int i = 0; i < nfx; ++i) {
				const float wg = std::min(aux[i].props.gain, MAX_GAIN);
				for (int k = 0; k < WET_CH; ++k) aux[i].gains[c][k] = co[k] * wg;
			}
		}
		send_filters(direct, direct.props.gain_hf, direct.props.gain_lf);
		for (int i = 0; i < nfx; ++i) send_filters(aux[i], aux[i].props.gain_hf, aux[i].props.gain_lf);
	}

	static const float* filtered(SendState& s, int c, const float* src, float* tmp, int n) // apply_filters, :3101-3143
	{
		switch (s.filter_type) {
		case 1: s.lpf[c].process(n, src, tmp); s.hpf[c].pass_through(n, tmp); return tmp;
		case 2: s.lpf[c].pass_through(n, src); s.hpf[c].process(n, src, tmp); return tmp;
		case 3: for (int i = 0; i < n; ++i) tmp[i] = s.hpf[c].tick(s.lpf[c].tick(src[i])); return tmp;
		default: s.lpf[c].pass_through(n, src); s.hpf[c].pass_through(n, src); return src;
		}
	}

	void chunk(int n, const float* src, float* dst) // mix_data + mix_source + write_f32, :2917-3037, :3414-3431
	{
		const int ch = dev.channels;
		for (int c = 0; c < ch; ++c) std::fill_n(bus[c], n, 0.0F);
		refresh();
		for (int i = 0; i < nfx; ++i)
			for (int k = 0; k < WET_CH; ++k) std::fill_n(slots[i].wet[k], n, 0.0F);
		static thread_local float in[CHUNK], tmp[CHUNK];
		for (int c = 0; c < ch; ++c) {
			for (int i = 0; i < n; ++i) in[i] = src[i * ch + c];
			const float* s = filtered(direct, c, in, tmp, n);
			for (int k = 0; k < ch; ++k) add_scaled(bus, k, 0, n, s, direct.gains[c][k]);
			for (int a = 0; a < nfx; ++a) {
				if (!aux[a].enabled) continue;
				s = filtered(aux[a], c, in, tmp, n);
				for (int k = 0; k < WET_CH; ++k) {
					const float g = aux[a].gains[c][k];
					if (!(std::abs(g) > SILENCE)) continue;
					for (int i = 0; i < n; ++i) slots[a].wet[k][i] += s[i] * g;
				}
			}
		}
		for (int a = 0; a < nfx; ++a) slots[a].fx->process(n, slots[a].wet, bus, ch);
		for (int c = 0; c < ch; ++c)
			for (int i = 0; i < n; ++i) dst[i * ch + c] = bus[c][i];
	}

	void mix(int n, const float* src, float* dst) // Api::mix, :3785-3829
	{
		const int ch = dev.channels;
		for (int done = 0; done < n;) {
			const int todo = std::min(n - done, CHUNK);
			chunk(todo, src + static_cast<size_t>(done) * ch, dst + static_cast<size_t>(done) * ch);
			done += todo;
		}
	}
};

inline uint32_t fmix32(uint32_t h)
{
	h ^= h >> 16; h *= 0x85EBCA6BU; h ^= h >> 13; h *= 0xC2B2AE35U; h ^= h >> 16;
	return h;
}

} // namespace

extern "C" {

const char* orc_kind() { return "port"; }

void* orc_create(int channel_format, int sampling_rate, int effect_count)
{
	Instance* p = new Instance;
	if (!p->init(channel_format, sampling_rate, effect_count)) { delete p; return nullptr; }
	return p;
}
void orc_destroy(void* h) { delete static_cast<Instance*>(h); }
int orc_channel_count(void* h) { return static_cast<Instance*>(h)->dev.channels; }

int orc_set_effect_type(void* h, int slot, int type)
{
	Instance* p = static_cast<Instance*>(h);
	if (slot < 0 || slot >= p->nfx) return 0;
	p->slots[slot].deferred.type = type;
	set_defaults(p->slots[slot].deferred);
	return 1;
}
int orc_set_effect_props(void* h, int slot, const void* props)
{
	Instance* p = static_cast<Instance*>(h);
	if (slot < 0 || slot >= p->nfx) return 0;
	std::memcpy(&p->slots[slot].deferred.props, props, sizeof(Props));
	return 1;
}
int orc_get_effect(void* h, int slot, int* type, void* props)
{
	Instance* p = static_cast<Instance*>(h);
	if (slot < 0 || slot >= p->nfx) return 0;
	*type = p->slots[slot].active.type;
	std::memcpy(props, &p->slots[slot].active.props, sizeof(Props));
	return 1;
}
int orc_get_deferred_effect(void* h, int slot, int* type, void* props)
{
	Instance* p = static_cast<Instance*>(h);
	if (slot < 0 || slot >= p->nfx) return 0;
	*type = p->slots[slot].deferred.type;
	std::memcpy(props, &p->slots[slot].deferred.props, sizeof(Props));
	return 1;
}
int orc_set_send_props(void* h, int send_index, const float* g)
{
	Instance* p = static_cast<Instance*>(h);
	if (send_index >= p->nfx) return 0;
	const Send3 s = {g[0], g[1], g[2]};
	if (send_index < 0) p->direct.deferred = s; else p->aux[send_index].props = s; // aux bypasses deferral, :3728-3731
	return 1;
}
int orc_apply(void* h) { static_cast<Instance*>(h)->apply(); return 1; }
int orc_mix(void* h, int frames, const float* src, float* dst)
{
	if (frames == 0) return 1;
	if (!src || !dst) return 0;
	static_cast<Instance*>(h)->mix(frames, src, dst);
	return 1;
}
int orc_sizeof_effect_props() { return static_cast<int>(sizeof(Props)); }
int orc_sizeof_effect() { return static_cast<int>(sizeof(Effect)); }

void orc_noise(uint32_t seed, uint32_t stream, int channels, int first_frame, int frames, float* dst)
{
	for (int n = 0; n < frames; ++n)
		for (int c = 0; c < channels; ++c) {
			const uint32_t hsh = fmix32(seed ^ (stream * 0x9E3779B9U) ^ (static_cast<uint32_t>(c) * 0x85EBCA6BU) ^
				(static_cast<uint32_t>(first_frame + n) * 0xC2B2AE35U));
			dst[n * channels + c] = (static_cast<float>(hsh >> 8) * (1.0F / 8388608.0F) - 1.0F) * 0.5F;
		}
}

// Same contract as orc_bench in oracle/ref_shim.cpp.
double orc_bench(int n_threads, int n_streams, int channel_format, int sampling_rate, const int* slot_types, int n_slots,
	int block_frames, int n_blocks, uint32_t seed, double* checksum_out)
{
	std::atomic<int> next{0};
	std::vector<double> sums(static_cast<size_t>(n_threads), 0.0);
	const auto t0 = std::chrono::steady_clock::now();
	auto worker = [&](int tid) {
		const int unique = 8;
		std::vector<float> src(static_cast<size_t>(unique) * block_frames * MAX_CH), dst(static_cast<size_t>(block_frames) * MAX_CH);
		int have = 0;
		for (;;) {
			const int s = next.fetch_add(1);
			if (s >= n_streams) break;
			std::unique_ptr<Instance> inst(new Instance);
			if (!inst->init(channel_format, sampling_rate, n_slots)) break;
			for (int i = 0; i < n_slots; ++i) { inst->slots[i].deferred.type = slot_types[i]; set_defaults(inst->slots[i].deferred); }
			inst->apply();
			const int ch = inst->dev.channels;
			if (have != ch) { orc_noise(seed, static_cast<uint32_t>(tid), ch, 0, unique * block_frames, src.data()); have = ch; }
			double acc = 0.0;
			for (int b = 0; b < n_blocks; ++b) {
				inst->mix(block_frames, src.data() + static_cast<size_t>(b % unique) * block_frames * ch, dst.data());
				acc += dst[static_cast<size_t>(block_frames) * ch - 1];
			}
			sums[static_cast<size_t>(tid)] += acc;
		}
	};
	std::vector<std::thread> threads;
	for (int t = 0; t < n_threads; ++t) threads.emplace_back(worker, t);
	for (auto& t : threads) t.join();
	const auto t1 = std::chrono::steady_clock::now();
	double total = 0.0;
	for (double v : sums) total += v;
	if (checksum_out) *checksum_out = total;
	return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
