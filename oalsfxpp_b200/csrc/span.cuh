// span.cuh -- kernels that are BLOCK-PARALLEL IN TIME: the single-reverb-slot signature (cfg1) and the 4-slot
// equalizer + chorus/flanger + echo + (EAX) reverb chain (cfg2, shards of cfg4), for launches of FEW tiles.
//
// With a thread per stream a block of F frames is F dependent iterations of a ~900-instruction sample body,
// however few streams there are.  But every feedback path of these effects runs through a delay line: the only
// sample-to-sample recurrences that do NOT are short IIR filters -- the reverb's input shelves (oalsfxpp.cpp:
// 7821-7832) and T60 filters (:7691-7719), the equalizer's cascade (:5161-5213) and the echo's damping filter
// (:4887-4962).  So a SPAN of T consecutive frames is processed in three phases, lanes = streams, warps = time:
//
//   A  (parallel over frames)  wet encodes; B-format -> A-format; late taps + late line reads; echo tap + input
//                              -> shared memory (the inputs of the recurrences)
//   B  (serial over frames)    one warp per recurrence: shelf pairs -> main delay line, T60 pairs, equalizer
//                              pair / equalizer channel 3 + echo filter -> echo ring; in place in shared memory
//   C  (parallel over frames)  dry mix, equalizer pan, chorus, echo taps, early reflections, late all-pass +
//                              scatter + line feed, pan of the 8 reverb lines, output
//
// and the phases of NEIGHBOURING spans overlap as a software pipeline -- in iteration i the parallel warps run
// C(i-1) and A(i+1) while the serial warps run B(i), one CTA barrier per iteration -- so the recurrences cost
// issue slots, not time (round 1 ran A, B, C one after the other: 43 % of the kernel was phase B with 4 of 16
// warps busy).  The staging buffers rotate over three spans.
//
// Every value is computed by the same expression as in fx.cuh / fx_reverb.cuh (same helpers, same order of
// additions), only the schedule differs -- the result is bit-identical.  What makes the schedule legal is
// checked on the host per coefficient block (plan_frames): a ring position written in phase Pw of span s and
// read in phase Pr of span s' must be complete before it is read and must not be overwritten before it is
// read, given that phase A of a span runs one iteration before its phase B and phase C one iteration after.
//
// Steady state only: no parameter update pending, tap cross-fade finished, modulator quiet, no pan-gain ramp
// (the host does not launch these kernels with an update pending; the device verifies the rest per tile and
// otherwise runs the exact thread-per-stream body).
//
// Memory: ring rows of the coming iteration are pulled into L2 by bulk prefetches (cp.async.bulk.prefetch.L2:
// one instruction per tap and span -- a span's T positions of one line are one contiguous run of T x 128
// bytes), the reads themselves bypass L1 (ld.global.cg).
#ifndef OALSFX_SPAN_CUH
#define OALSFX_SPAN_CUH

#include <vector>

#include "mix.cuh"

namespace oalsfx {
namespace span {

#if defined(__CUDACC__)
#define OALSFX_CX __host__ __device__ constexpr
#else
#define OALSFX_CX constexpr
#endif

constexpr int kParallelWarps = 8;
constexpr int kBuffers = 3;               // spans in flight: A(i+1), B(i), C(i-1)
OALSFX_CX int serial_warps(bool chain) { return chain ? 6 : 4; }
OALSFX_CX int threads(bool chain) { return (kParallelWarps + serial_warps(chain)) * kLanes; }
OALSFX_CX int staged_words(bool chain) { return chain ? 12 : 8; }
// frames a staging buffer holds (= the longest span) for SL streams per CTA
OALSFX_CX int capacity(int stream_lanes) { return stream_lanes == kLanes ? 32 : 64; }
OALSFX_CX int shared_floats(bool chain, int stream_lanes)
{
	return kBuffers * staged_words(chain) * capacity(stream_lanes) * stream_lanes;
}

// staged words of a frame
constexpr int kWA = 0;   // 4: A-format lines before the shelves                (A -> B)
constexpr int kWL = 4;   // 4: late lines before / after the T60 filters        (A -> B -> C)
constexpr int kWQ = 8;   // 3: equalizer wet channels 0, 1, 3 before / after     (A -> B -> C)
constexpr int kWX = 11;  // 1: echo filter input                                 (A -> B)

// ---- legality ------------------------------------------------------------------------------------------------
// Phases as iteration offsets: phase P of span s runs in iteration s + P.
constexpr int kPhA = -1, kPhB = 0, kPhC = 1;

struct Legality {
	int t;
	bool ok = true;
	// A read `delay` frames behind the frame position in phase pr against a write `wofs` frames behind it in
	// phase pw, on a ring of `len` positions.  zero_is_new: a read of the position written in the same frame
	// sees the new value (the reference writes first), else the one written `len` frames earlier.
	void pair(int pr, int delay, int pw, int wofs, int len, bool zero_is_new)
	{
		int k = ((delay - wofs) % len + len) % len;  // frames between the write of a position and this read of it
		if (k == 0 && !zero_is_new) {
			k = len;
		}
		// read after write: writer in iteration span(n - k) + pw, reader in span(n) + pr; min over n of the span
		// difference is floor(k / t)
		ok = ok && k / t > pw - pr;
		// write after read: the position is next written len - k frames after the read
		ok = ok && (len - k) / t > pr - pw;
	}
	// two writers of one ring: the writes to a position keep their order
	void writers(int p1, int o1, int p2, int o2, int len)
	{
		const int d = ((o2 - o1) % len + len) % len;   // writer 2 reaches a position d frames after writer 1
		ok = ok && d / t > p1 - p2;
		ok = ok && (len - d) / t > p2 - p1;
	}
};

inline bool reverb_legal(const ReverbCoef& c, int t)
{
	if (c.mod_depth != 0.0F) {
		return false;
	}
	Legality g;
	g.t = t;
	const int len0 = c.mask[0] + 1;
	for (int l = 0; l < 4; ++l) {
		// main line: shelves write at the position (B), the early scatter feeds it late_feed_tap behind (C)
		g.pair(kPhC, c.early_tap[l], kPhB, 0, len0, true);
		g.pair(kPhC, c.early_tap[l], kPhC, c.late_feed_tap, len0, false);
		g.pair(kPhA, c.late_tap[l], kPhB, 0, len0, true);
		g.pair(kPhA, c.late_tap[l], kPhC, c.late_feed_tap, len0, true);
		g.pair(kPhC, c.early_ap_off[l], kPhC, 0, c.mask[1] + 1, false);
		g.pair(kPhC, c.early_off[l], kPhC, 0, c.mask[2] + 1, true);
		g.pair(kPhC, c.late_ap_off[l], kPhC, 0, c.mask[3] + 1, false);
		g.pair(kPhA, c.late_off[l], kPhC, 0, c.mask[4] + 1, true);
	}
	g.writers(kPhB, 0, kPhC, c.late_feed_tap, len0);
	return g.ok;
}

inline bool mod_delay_legal(const ModDelayCoef& c, int t)
{
	Legality g;
	g.t = t;
	// the LFO keeps the delay within delay +- depth (oalsfxpp.cpp:4246-4276); two frames of slack for the rounding
	const int dmin = c.delay - static_cast<int>(c.depth) - 2, dmax = c.delay + static_cast<int>(c.depth) + 2;
	if (dmin < 1 || c.lfo_range <= kMaxBlockFrames) {
		return false;
	}
	g.pair(kPhC, dmin, kPhC, 0, c.mask + 1, false);
	g.pair(kPhC, dmax, kPhC, 0, c.mask + 1, false);
	return g.ok && dmax < c.mask + 1;
}

inline bool echo_legal(const EchoCoef& c, int t)
{
	Legality g;
	g.t = t;
	g.pair(kPhC, c.tap1, kPhB, 0, c.mask + 1, false);
	g.pair(kPhC, c.tap2, kPhB, 0, c.mask + 1, false);
	g.pair(kPhA, c.tap2, kPhB, 0, c.mask + 1, false);
	return g.ok;
}

// Longest legal span (32 / 64 by capacity, then halves down to 16) for the coefficient blocks of this launch, or 0.
// chain: slots 0..3 = equalizer, chorus / flanger, echo, reverb; else slot 0 = reverb.
inline int plan_frames(const MixArgs& a, bool chain, int stream_lanes)
{
	for (int t = capacity(stream_lanes); t >= 16; t /= 2) {
		bool ok = reverb_legal(a.slot[chain ? 3 : 0].u.reverb, t);
		if (chain) {
			ok = ok && mod_delay_legal(a.slot[1].u.mod_delay, t) && echo_legal(a.slot[2].u.echo, t);
		}
		if (ok) {
			return t;
		}
	}
	return 0;
}

// ---- staging buffers -------------------------------------------------------------------------------------------
// [buffer][word][frame][stream of this CTA]; `base` includes this thread's stream.
template <int SL, int W, int CAP>
struct Stage {
	float* base;
	OALSFX_HD float& at(int buf, int word, int t) const { return base[((buf * W + word) * CAP + t) * SL]; }
};

// ring read that bypasses L1 (every ring word is read once per tap)
OALSFX_HD float ring_ld(const LaneMem& m, int word)
{
#if defined(__CUDA_ARCH__)
	return __ldcg(m.p + static_cast<unsigned>(word) * kLanes);
#else
	return m.ld(word);
#endif
}

// wet bus of one aux send without shelf filters: the arithmetic of SlotRunner::step (mix.cuh)
template <int CT>
OALSFX_HD void encode_wet(const SendCoef& sc, const float* x, float* wet)
{
	if (CT == 2) {
		const F2 zero = f2(0.0F, 0.0F), x0 = f2_bcast(x[0]), x1 = f2_bcast(x[1]);
		const F2 wa = (zero + (x0 * f2(sc.gains[0][0], sc.gains[0][1]))) + (x1 * f2(sc.gains[1][0], sc.gains[1][1]));
		const F2 wb = (zero + (x0 * f2(sc.gains[0][2], sc.gains[0][3]))) + (x1 * f2(sc.gains[1][2], sc.gains[1][3]));
		wet[0] = f2_lo(wa);
		wet[1] = f2_hi(wa);
		wet[2] = f2_lo(wb);
		wet[3] = f2_hi(wb);
	} else {
		OALSFX_UNROLL
		for (int k = 0; k < kWetChannels; ++k) {
			wet[k] = 0.0F;
		}
		OALSFX_UNROLL
		for (int c = 0; c < CT; ++c) {
			OALSFX_UNROLL
			for (int k = 0; k < kWetChannels; ++k) {
				wet[k] += x[c] * sc.gains[c][k];
			}
		}
	}
}

// ---- per-stream context ------------------------------------------------------------------------------------------
template <int CT, bool CHAIN>
struct Context {
	using R = FxReverbTail;
	static constexpr int RP = CHAIN ? 3 : 0;  // slot position of the reverb
	const float* src;
	float* dst;
	bool io_ok;
	LaneMem ring_rev, ring_mod, ring_echo;
	uint32_t *st_rev, *st_eq, *st_mod, *st_echo, *ss;
	int32_t rev_off, mod_off, echo_off;       // ring offsets at the start of the block
	int32_t mod_ph[2];                        // chorus LFO phases at the start of the block
	float gain[8][CT];                        // the reverb's pan gains, inaudible ones as exact zeros

	// Loads the per-stream state the phases need; returns whether the stream is in the steady state.
	OALSFX_HD bool setup(const MixArgs& a, int tile, int lane)
	{
		io_ok = tile * kLanes + lane < a.num_streams;
		src = a.src + tile * a.io_ts + lane * a.io_ls;
		dst = a.dst + tile * a.io_ts + lane * a.io_ls;
		auto state = [&](int p) { return a.slot_state[p] + (static_cast<long long>(tile) * kSlotStateWords) * kLanes + lane; };
		auto ring = [&](int p) { return a.ring[p] + static_cast<long long>(tile) * a.ring_tile_stride[p] + lane; };
		ss = a.send_state + (static_cast<long long>(tile) * kSendStateWords) * kLanes + lane;
		st_rev = state(RP);
		ring_rev.p = ring(RP);
		const ReverbCoef& c = a.slot[RP].u.reverb;
		rev_off = static_cast<int32_t>(st_rev[(R::kWScalars + 0) * kLanes]);
		const int32_t fade_count = static_cast<int32_t>(st_rev[(R::kWScalars + 1) * kLanes]);
		const float mod_filter = word_as_float(st_rev[(R::kWScalars + 4) * kLanes]);
		bool ok = a.span_frames >= 1 && a.update_mask == 0 && fade_count >= R::kFadeSamples && c.mod_depth == 0.0F && mod_filter == 0.0F;
		float cur[8][CT];
		OALSFX_UNROLL
		for (int l = 0; l < 8; ++l) {
			OALSFX_UNROLL
			for (int k = 0; k < CT; ++k) {
				cur[l][k] = word_as_float(st_rev[(R::kWGain + l * kMaxChannels + k) * kLanes]);
				gain[l][k] = audible(cur[l][k]) ? cur[l][k] : 0.0F;
			}
		}
		// no pan-gain ramp in any sub-chunk (FxReverbT::begin_sub: sub-chunks of <= 256 frames, step = (target - gain) / frames left)
		for (int base = 0; base < a.frames; base += R::kMaxUpdate) {
			const float delta = 1.0F / static_cast<float>(a.frames - base);
			OALSFX_UNROLL
			for (int l = 0; l < 8; ++l) {
				const float* target = (l < 4 ? c.pan_early[l] : c.pan_late[l - 4]);
				OALSFX_UNROLL
				for (int k = 0; k < CT; ++k) {
					ok = ok && !(fabsf((target[k] - cur[l][k]) * delta) > FLT_EPSILON);
				}
			}
		}
		if (CHAIN) {
			st_eq = state(0);
			st_mod = state(1);
			st_echo = state(2);
			ring_mod.p = ring(1);
			ring_echo.p = ring(2);
			const ModDelayCoef& m = a.slot[1].u.mod_delay;
			mod_off = static_cast<int32_t>(st_mod[0]);
			mod_ph[0] = mod_off % m.lfo_range;                 // oalsfxpp.cpp:4137, 4146
			mod_ph[1] = (mod_off + m.lfo_disp) % m.lfo_range;
			echo_off = static_cast<int32_t>(st_echo[4 * kLanes]);
		}
		return ok;
	}

	OALSFX_HD void load_input(const MixArgs& a, int n, float* x) const
	{
		OALSFX_UNROLL
		for (int ch = 0; ch < CT; ++ch) {
			x[ch] = io_ok ? src[n * a.io_fs + ch * a.io_cs] : 0.0F;
		}
	}

	// ---- phase A of block frame n (frame t of its span, staging buffer buf) ----
	struct TapsA { float late[4], lline[4], echo2; };

	OALSFX_HD void load_a(const MixArgs& a, int n, TapsA& k) const
	{
		const ReverbCoef& c = a.slot[RP].u.reverb;
		const int pos = rev_off + n;
		const int main_len = c.mask[0] + 1, lline_len = c.mask[4] + 1;
		OALSFX_UNROLL
		for (int l = 0; l < 4; ++l) {
			k.late[l] = ring_ld(ring_rev, c.ring_base[0] + l * main_len + ((pos - c.late_tap[l]) & c.mask[0]));
			k.lline[l] = ring_ld(ring_rev, c.ring_base[4] + l * lline_len + ((pos - c.late_off[l]) & c.mask[4]));
		}
		if (CHAIN) {
			const EchoCoef& e = a.slot[2].u.echo;
			k.echo2 = ring_ld(ring_echo, (echo_off + n - e.tap2) & e.mask);
		}
	}

	template <class S>
	OALSFX_HD void phase_a(const MixArgs& a, const S& sg, int buf, int t, const float* x, const TapsA& k) const
	{
		const ReverbCoef& c = a.slot[RP].u.reverb;
		float wet[kWetChannels];
		if (CHAIN) {
			encode_wet<CT>(a.aux[0], x, wet);          // equalizer: wet channel 2 is dead (FxEqualizer::kDeadWet)
			sg.at(buf, kWQ + 0, t) = wet[0];
			sg.at(buf, kWQ + 1, t) = wet[1];
			sg.at(buf, kWQ + 2, t) = wet[3];
			encode_wet<CT>(a.aux[2], x, wet);          // echo: in = tap2 + wet[0] (oalsfxpp.cpp:4921-4925)
			sg.at(buf, kWX, t) = k.echo2 + wet[0];
		}
		encode_wet<CT>(a.aux[RP], x, wet);
		{
			// B-format -> A-format, as reverb_input_stage (fx_reverb.cuh)
			constexpr float q = 0.288675134595F;
			const F2 zero = f2(0.0F, 0.0F);
			const F2 p0 = f2_bcast(wet[0] * q), p3 = f2_bcast(wet[3] * q);
			const F2 p1 = f2_bcast(wet[1]) * f2(q, -q);
			const F2 p2 = f2_bcast(wet[2]) * f2(q, -q);
			const F2 a01 = (((zero + p0) + p1) + p2) + p3;
			const F2 a23 = (((zero + p0) + p1) - p2) - p3;
			sg.at(buf, kWA + 0, t) = f2_lo(a01);
			sg.at(buf, kWA + 1, t) = f2_hi(a01);
			sg.at(buf, kWA + 2, t) = f2_lo(a23);
			sg.at(buf, kWA + 3, t) = f2_hi(a23);
		}
		// late reverb up to the T60 filters (modulation delay 0: steady state), FxReverbT::body
		F2 fa = f2(k.late[0], k.late[1]) * c.density_gain;
		F2 fb = f2(k.late[2], k.late[3]) * c.density_gain;
		fa = fa + f2(k.lline[0], k.lline[1]);
		fb = fb + f2(k.lline[2], k.lline[3]);
		sg.at(buf, kWL + 0, t) = f2_lo(fa);
		sg.at(buf, kWL + 1, t) = f2_hi(fa);
		sg.at(buf, kWL + 2, t) = f2_lo(fb);
		sg.at(buf, kWL + 3, t) = f2_hi(fb);
	}

	// ---- phase C ----
	struct TapsC { float early[4], eap[4], eline[4], lap[4], mod[2], echo[2]; int32_t mod_pos; };

	OALSFX_HD void load_c(const MixArgs& a, int n, TapsC& k) const
	{
		const ReverbCoef& c = a.slot[RP].u.reverb;
		const int pos = rev_off + n;
		const int main_len = c.mask[0] + 1, eap_len = c.mask[1] + 1, eline_len = c.mask[2] + 1, lap_len = c.mask[3] + 1;
		OALSFX_UNROLL
		for (int l = 0; l < 4; ++l) {
			k.early[l] = ring_ld(ring_rev, c.ring_base[0] + l * main_len + ((pos - c.early_tap[l]) & c.mask[0]));
			k.eap[l] = ring_ld(ring_rev, c.ring_base[1] + l * eap_len + ((pos - c.early_ap_off[l]) & c.mask[1]));
			k.eline[l] = ring_ld(ring_rev, c.ring_base[2] + l * eline_len + ((pos - c.early_off[l]) & c.mask[2]));
			k.lap[l] = ring_ld(ring_rev, c.ring_base[3] + l * lap_len + ((pos - c.late_ap_off[l]) & c.mask[3]));
		}
		if (CHAIN) {
			const ModDelayCoef& m = a.slot[1].u.mod_delay;
			const int32_t mlen = m.mask + 1;
			k.mod_pos = mod_off + n;
			OALSFX_UNROLL
			for (int side = 0; side < 2; ++side) {
				int32_t ph = mod_ph[side] + n;         // n <= 2048 < lfo_range (host-checked): at most one wrap
				ph = (ph >= m.lfo_range ? ph - m.lfo_range : ph);
				const int32_t d = FxModDelay::lfo_delay(m, ph);
				k.mod[side] = ring_ld(ring_mod, side * mlen + ((k.mod_pos - d) & m.mask));
			}
			const EchoCoef& e = a.slot[2].u.echo;
			k.echo[0] = ring_ld(ring_echo, (echo_off + n - e.tap1) & e.mask);
			k.echo[1] = ring_ld(ring_echo, (echo_off + n - e.tap2) & e.mask);
		}
	}

	// vector_allpass_x with the taps already read (FxReverbT::vector_allpass2, fx_reverb.cuh)
	OALSFX_HD void allpass(const ReverbCoef& c, F2& va, F2& vb, const float* tp, int ring_idx, int pos) const
	{
		const int len = c.mask[ring_idx] + 1, word0 = c.ring_base[ring_idx], mask = c.mask[ring_idx];
		const F2 ta = f2(tp[0], tp[1]), tb = f2(tp[2], tp[3]);
		const F2 ina = va, inb = vb;
		va = ta - (ina * c.ap_feed_coeff);
		vb = tb - (inb * c.ap_feed_coeff);
		F2 fa = ina + (va * c.ap_feed_coeff);
		F2 fb = inb + (vb * c.ap_feed_coeff);
		R::scatter2(fa, fb, c.mix_x, c.mix_y);
		ring_rev.st(word0 + 0 * len + (pos & mask), f2_lo(fa));
		ring_rev.st(word0 + 1 * len + (pos & mask), f2_hi(fa));
		ring_rev.st(word0 + 2 * len + (pos & mask), f2_lo(fb));
		ring_rev.st(word0 + 3 * len + (pos & mask), f2_hi(fb));
	}

	template <class S>
	OALSFX_HD void phase_c(const MixArgs& a, const S& sg, int buf, int t, int n, const float* x, const TapsC& k) const
	{
		const ReverbCoef& c = a.slot[RP].u.reverb;
		float acc[CT];
		OALSFX_UNROLL
		for (int ch = 0; ch < CT; ++ch) {
			acc[ch] = 0.0F;
		}
		OALSFX_UNROLL
		for (int ch = 0; ch < CT; ++ch) {
			pan_add<CT, true>(acc, CT, a.direct.gains[ch], x[ch]); // direct send (oalsfxpp.cpp:2924-2950)
		}
		if (CHAIN) {
			{
				// equalizer outputs -> bus (FxEqualizer::step)
				const EqualizerCoef& q = a.slot[0].u.equalizer;
				const float q0 = sg.at(buf, kWQ + 0, t), q1 = sg.at(buf, kWQ + 1, t), q3 = sg.at(buf, kWQ + 2, t);
				pan_add<CT, true>(acc, CT, q.gains[0], q0);
				pan_add<CT, true>(acc, CT, q.gains[1], q1);
				pan_add<CT, true>(acc, CT, q.gains[3], q3);
			}
			{
				// chorus / flanger (FxModDelay::step; every delay >= the span, so never the sample just written)
				const ModDelayCoef& m = a.slot[1].u.mod_delay;
				float wet[kWetChannels];
				encode_wet<CT>(a.aux[1], x, wet);
				const int32_t mlen = m.mask + 1, mpos = k.mod_pos & m.mask;
				float tt[2];
				OALSFX_UNROLL
				for (int side = 0; side < 2; ++side) {
					tt[side] = k.mod[side] * m.feedback;
					ring_mod.st(side * mlen + mpos, wet[0] + tt[side]);
				}
				if (CT == 2) {
					pan_add<CT, true>(acc, CT, m.gains[0], tt[0]);
					pan_add<CT, true>(acc, CT, m.gains[1], tt[1]);
				} else {
					OALSFX_UNROLL
					for (int ch = 0; ch < CT; ++ch) {
						acc[ch] += tt[0] * m.gains[0][ch];
						acc[ch] += tt[1] * m.gains[1][ch];
					}
				}
			}
			{
				// echo taps -> bus (FxEcho::step)
				const EchoCoef& e = a.slot[2].u.echo;
				if (CT == 2) {
					pan_add<CT, true>(acc, CT, e.gains[0], k.echo[0]);
					pan_add<CT, true>(acc, CT, e.gains[1], k.echo[1]);
				} else {
					OALSFX_UNROLL
					for (int ch = 0; ch < CT; ++ch) {
						acc[ch] += k.echo[0] * e.gains[0][ch];
						acc[ch] += k.echo[1] * e.gains[1][ch];
					}
				}
			}
		}
		const int pos = rev_off + n;
		const int main_len = c.mask[0] + 1, main0 = c.ring_base[0], main_mask = c.mask[0];
		const int eline_len = c.mask[2] + 1, eline0 = c.ring_base[2], eline_mask = c.mask[2];
		const int lline_len = c.mask[4] + 1, lline0 = c.ring_base[4], lline_mask = c.mask[4];
		float out8[8];
		// early reflections (the EARLY half of FxReverbT::body)
		F2 fa = f2(k.early[0], k.early[1]) * f2(c.early_tap_coeff[0], c.early_tap_coeff[1]);
		F2 fb = f2(k.early[2], k.early[3]) * f2(c.early_tap_coeff[2], c.early_tap_coeff[3]);
		allpass(c, fa, fb, k.eap, 1, pos);
		ring_rev.st(eline0 + 0 * eline_len + (pos & eline_mask), f2_hi(fb));
		ring_rev.st(eline0 + 1 * eline_len + (pos & eline_mask), f2_lo(fb));
		ring_rev.st(eline0 + 2 * eline_len + (pos & eline_mask), f2_hi(fa));
		ring_rev.st(eline0 + 3 * eline_len + (pos & eline_mask), f2_lo(fa));
		fa = fa + (f2(k.eline[0], k.eline[1]) * f2(c.early_coeff[0], c.early_coeff[1]));
		fb = fb + (f2(k.eline[2], k.eline[3]) * f2(c.early_coeff[2], c.early_coeff[3]));
		out8[0] = f2_lo(fa);
		out8[1] = f2_hi(fa);
		out8[2] = f2_lo(fb);
		out8[3] = f2_hi(fb);
		{
			F2 ra = fa, rb = fb;
			R::scatter2_reversed(ra, rb, c.mix_x, c.mix_y);
			const int feed = (pos - c.late_feed_tap) & main_mask;
			ring_rev.st(main0 + 0 * main_len + feed, f2_hi(rb));
			ring_rev.st(main0 + 1 * main_len + feed, f2_lo(rb));
			ring_rev.st(main0 + 2 * main_len + feed, f2_hi(ra));
			ring_rev.st(main0 + 3 * main_len + feed, f2_lo(ra));
		}
		// late reverb after the T60 filters
		fa = f2(sg.at(buf, kWL + 0, t), sg.at(buf, kWL + 1, t));
		fb = f2(sg.at(buf, kWL + 2, t), sg.at(buf, kWL + 3, t));
		allpass(c, fa, fb, k.lap, 3, pos);
		out8[4] = f2_lo(fa);
		out8[5] = f2_hi(fa);
		out8[6] = f2_lo(fb);
		out8[7] = f2_hi(fb);
		{
			F2 ra = fa, rb = fb;
			R::scatter2_reversed(ra, rb, c.mix_x, c.mix_y);
			ring_rev.st(lline0 + 0 * lline_len + (pos & lline_mask), f2_hi(rb));
			ring_rev.st(lline0 + 1 * lline_len + (pos & lline_mask), f2_lo(rb));
			ring_rev.st(lline0 + 2 * lline_len + (pos & lline_mask), f2_hi(ra));
			ring_rev.st(lline0 + 3 * lline_len + (pos & lline_mask), f2_lo(ra));
		}
		// pan with static gains (oalsfxpp.cpp:6142-6166, 2752-2798): inaudible gains are exact zeros here
		OALSFX_UNROLL
		for (int l = 0; l < 8; ++l) {
			pan_add<CT, true>(acc, CT, gain[l], out8[l]);
		}
		if (io_ok) {
			OALSFX_UNROLL
			for (int ch = 0; ch < CT; ++ch) {
				dst[n * a.io_fs + ch * a.io_cs] = acc[ch];
			}
		}
	}

	// Send filter histories of the processed sends: with no shelf filter active they are the last two input
	// samples (oalsfxpp.cpp:1038-1056).
	OALSFX_HD void store_send_history(const MixArgs& a, int send) const
	{
		OALSFX_UNROLL
		for (int ch = 0; ch < CT; ++ch) {
			const float last1 = io_ok ? src[(a.frames - 1) * a.io_fs + ch * a.io_cs] : 0.0F;
			const float last2 = io_ok ? src[(a.frames - 2) * a.io_fs + ch * a.io_cs] : 0.0F;
			SendHist h;
			h.lp.x0 = h.lp.y0 = h.hp.x0 = h.hp.y0 = last1;
			h.lp.x1 = h.lp.y1 = h.hp.x1 = h.hp.y1 = last2;
			store_words(h, ss + (send * kMaxChannels + ch) * 8 * kLanes);
		}
	}

	// Scalars that only advance in the steady state.
	OALSFX_HD void store_scalars(const MixArgs& a) const
	{
		st_rev[(R::kWScalars + 0) * kLanes] = static_cast<uint32_t>(rev_off + a.frames);
		// the quiet modulator only advances its index (FxReverbT::body): +1 per frame, wrapping at the range
		int32_t mod_index = static_cast<int32_t>(st_rev[(R::kWScalars + 2) * kLanes]);
		int32_t mod_range = static_cast<int32_t>(st_rev[(R::kWScalars + 3) * kLanes]);
		if (mod_range == 0) {
			mod_range = 1;
		}
		mod_index = static_cast<int32_t>((static_cast<long long>(mod_index) + a.frames) % mod_range);
		st_rev[(R::kWScalars + 2) * kLanes] = static_cast<uint32_t>(mod_index);
		st_rev[(R::kWScalars + 3) * kLanes] = static_cast<uint32_t>(mod_range);
		if (CHAIN) {
			st_mod[0] = static_cast<uint32_t>(mod_off + a.frames);
			st_echo[4 * kLanes] = static_cast<uint32_t>(echo_off + a.frames);
		}
	}
};

// ---- the recurrences (phase B), one object per serial warp ----------------------------------------------------------
// Two biquads with common coefficients as one packed filter (FilterState::process, oalsfxpp.cpp:984-1036).
struct Biquad2 {
	F2 x0, x1, y0, y1;
	OALSFX_HD void load(const uint32_t* a, const uint32_t* b)
	{
		BiquadHist ha, hb;
		load_words(ha, a);
		load_words(hb, b);
		x0 = f2(ha.x0, hb.x0);
		x1 = f2(ha.x1, hb.x1);
		y0 = f2(ha.y0, hb.y0);
		y1 = f2(ha.y1, hb.y1);
	}
	OALSFX_HD void store(uint32_t* a, uint32_t* b) const
	{
		const BiquadHist ha = {f2_lo(x0), f2_lo(x1), f2_lo(y0), f2_lo(y1)}, hb = {f2_hi(x0), f2_hi(x1), f2_hi(y0), f2_hi(y1)};
		store_words(ha, a);
		store_words(hb, b);
	}
	OALSFX_HD F2 step(const Biquad& c, F2 x)
	{
		const F2 y = (x * c.b0) + (x0 * c.b1) + (x1 * c.b2) - (y0 * c.a1) - (y1 * c.a2);
		x1 = x0;
		x0 = x;
		y1 = y0;
		y0 = y;
		return y;
	}
};

// Master shelves of reverb lines (2p, 2p+1) -> main delay line (reverb_input_stage, fx_reverb.cuh).
struct ShelfPair {
	using R = FxReverbTail;
	Biquad2 lp, hp;
	OALSFX_HD void load(const uint32_t* st, int p)
	{
		lp.load(st + (R::kWLp + (2 * p) * 4) * kLanes, st + (R::kWLp + (2 * p + 1) * 4) * kLanes);
		hp.load(st + (R::kWHp + (2 * p) * 4) * kLanes, st + (R::kWHp + (2 * p + 1) * 4) * kLanes);
	}
	OALSFX_HD void store(uint32_t* st, int p) const
	{
		lp.store(st + (R::kWLp + (2 * p) * 4) * kLanes, st + (R::kWLp + (2 * p + 1) * 4) * kLanes);
		hp.store(st + (R::kWHp + (2 * p) * 4) * kLanes, st + (R::kWHp + (2 * p + 1) * 4) * kLanes);
	}
	template <class S>
	OALSFX_HD void run(const ReverbCoef& c, const S& sg, int buf, int count, const LaneMem& ring, int pos0, int p)
	{
		const int main_len = c.mask[0] + 1, main0 = c.ring_base[0], main_mask = c.mask[0];
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
		for (int t = 0; t < count; ++t) {
			F2 v = f2(sg.at(buf, kWA + 2 * p, t), sg.at(buf, kWA + 2 * p + 1, t));
			v = lp.step(c.lp, v);
			if (c.is_eax) {
				v = hp.step(c.hp, v);
			}
			const int at = (pos0 + t) & main_mask;
			ring.st(main0 + (2 * p) * main_len + at, f2_lo(v));
			ring.st(main0 + (2 * p + 1) * main_len + at, f2_hi(v));
		}
	}
};

// late_t60_filter of lines (2h, 2h+1), in place (oalsfxpp.cpp:7691-7719; FxReverbT::body).
struct T60Pair {
	using R = FxReverbTail;
	F2 p[2][2];
	OALSFX_HD void load(const uint32_t* st, int h)
	{
		OALSFX_UNROLL
		for (int q = 0; q < 4; ++q) {
			p[q >> 1][q & 1] = f2(word_as_float(st[(R::kWT60 + (2 * h) * 4 + q) * kLanes]), word_as_float(st[(R::kWT60 + (2 * h + 1) * 4 + q) * kLanes]));
		}
	}
	OALSFX_HD void store(uint32_t* st, int h) const
	{
		OALSFX_UNROLL
		for (int q = 0; q < 4; ++q) {
			st[(R::kWT60 + (2 * h) * 4 + q) * kLanes] = float_as_word(f2_lo(p[q >> 1][q & 1]));
			st[(R::kWT60 + (2 * h + 1) * 4 + q) * kLanes] = float_as_word(f2_hi(p[q >> 1][q & 1]));
		}
	}
	template <class S>
	OALSFX_HD void run(const ReverbCoef& c, const S& sg, int buf, int count, int h)
	{
		const int j = 2 * h;
		const F2 lf0 = f2(c.t60_lf[j][0], c.t60_lf[j + 1][0]), lf1 = f2(c.t60_lf[j][1], c.t60_lf[j + 1][1]), lf2 = f2(c.t60_lf[j][2], c.t60_lf[j + 1][2]);
		const F2 hf0 = f2(c.t60_hf[j][0], c.t60_hf[j + 1][0]), hf1 = f2(c.t60_hf[j][1], c.t60_hf[j + 1][1]), hf2 = f2(c.t60_hf[j][2], c.t60_hf[j + 1][2]);
		const F2 mid = f2(c.t60_mid[j], c.t60_mid[j + 1]);
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
		for (int t = 0; t < count; ++t) {
			const F2 in = f2(sg.at(buf, kWL + j, t), sg.at(buf, kWL + j + 1, t));
			const F2 o1 = (lf0 * in) + (lf1 * p[0][0]) + (lf2 * p[0][1]);
			p[0][0] = in;
			p[0][1] = o1;
			const F2 o2 = (hf0 * o1) + (hf1 * p[1][0]) + (hf2 * p[1][1]);
			p[1][0] = o1;
			p[1][1] = o2;
			const F2 out = mid * o2;
			sg.at(buf, kWL + j, t) = f2_lo(out);
			sg.at(buf, kWL + j + 1, t) = f2_hi(out);
		}
	}
};

// The equalizer's four cascaded bands on wet channels (0, 1) as a packed pair, in place (FxEqualizer, fx.cuh: the
// input history of band b + 1 is the output history of band b).
struct EqPair {
	F2 x0, x1, y0[4], y1[4];
	OALSFX_HD void load(const uint32_t* st)
	{
		OALSFX_UNROLL
		for (int b = 0; b < 4; ++b) {
			BiquadHist h0, h1;
			load_words(h0, st + ((b * 4 + 0) * 4) * kLanes);
			load_words(h1, st + ((b * 4 + 1) * 4) * kLanes);
			if (b == 0) {
				x0 = f2(h0.x0, h1.x0);
				x1 = f2(h0.x1, h1.x1);
			}
			y0[b] = f2(h0.y0, h1.y0);
			y1[b] = f2(h0.y1, h1.y1);
		}
	}
	OALSFX_HD void store(uint32_t* st) const
	{
		OALSFX_UNROLL
		for (int b = 0; b < 4; ++b) {
			const F2 ix0 = (b == 0 ? x0 : y0[b == 0 ? 0 : b - 1]), ix1 = (b == 0 ? x1 : y1[b == 0 ? 0 : b - 1]);
			const BiquadHist h0 = {f2_lo(ix0), f2_lo(ix1), f2_lo(y0[b]), f2_lo(y1[b])}, h1 = {f2_hi(ix0), f2_hi(ix1), f2_hi(y0[b]), f2_hi(y1[b])};
			store_words(h0, st + ((b * 4 + 0) * 4) * kLanes);
			store_words(h1, st + ((b * 4 + 1) * 4) * kLanes);
		}
	}
	template <class S>
	OALSFX_HD void run(const EqualizerCoef& c, const S& sg, int buf, int count)
	{
#if defined(__CUDA_ARCH__)
#pragma unroll 2
#endif
		for (int t = 0; t < count; ++t) {
			F2 v = f2(sg.at(buf, kWQ + 0, t), sg.at(buf, kWQ + 1, t));
			F2 in0 = x0, in1 = x1;
			x1 = x0;
			x0 = v;
			OALSFX_UNROLL
			for (int b = 0; b < 4; ++b) {
				const Biquad& q = c.band[b];
				const F2 y = (v * q.b0) + (in0 * q.b1) + (in1 * q.b2) - (y0[b] * q.a1) - (y1[b] * q.a2);
				in0 = y0[b];
				in1 = y1[b];
				y1[b] = y0[b];
				y0[b] = y;
				v = y;
			}
			sg.at(buf, kWQ + 0, t) = f2_lo(v);
			sg.at(buf, kWQ + 1, t) = f2_hi(v);
		}
	}
};

// The equalizer's wet channel 3 (scalar cascade, in place) and the echo's damping filter -> echo ring (FxEcho::step).
struct EqSingleEcho {
	float x0, x1, y0[4], y1[4];
	BiquadHist ef;
	OALSFX_HD void load(const uint32_t* st_eq, const uint32_t* st_echo)
	{
		OALSFX_UNROLL
		for (int b = 0; b < 4; ++b) {
			BiquadHist h;
			load_words(h, st_eq + ((b * 4 + 3) * 4) * kLanes);
			if (b == 0) {
				x0 = h.x0;
				x1 = h.x1;
			}
			y0[b] = h.y0;
			y1[b] = h.y1;
		}
		load_words(ef, st_echo);
	}
	OALSFX_HD void store(uint32_t* st_eq, uint32_t* st_echo) const
	{
		OALSFX_UNROLL
		for (int b = 0; b < 4; ++b) {
			const BiquadHist h = {b == 0 ? x0 : y0[b == 0 ? 0 : b - 1], b == 0 ? x1 : y1[b == 0 ? 0 : b - 1], y0[b], y1[b]};
			store_words(h, st_eq + ((b * 4 + 3) * 4) * kLanes);
		}
		store_words(ef, st_echo);
	}
	template <class S>
	OALSFX_HD void run(const EqualizerCoef& c, const EchoCoef& e, const S& sg, int buf, int count, const LaneMem& ring, int pos0)
	{
#if defined(__CUDA_ARCH__)
#pragma unroll 2
#endif
		for (int t = 0; t < count; ++t) {
			float v = sg.at(buf, kWQ + 2, t);
			float in0 = x0, in1 = x1;
			x1 = x0;
			x0 = v;
			OALSFX_UNROLL
			for (int b = 0; b < 4; ++b) {
				const Biquad& q = c.band[b];
				const float y = (q.b0 * v) + (q.b1 * in0) + (q.b2 * in1) - (q.a1 * y0[b]) - (q.a2 * y1[b]);
				in0 = y0[b];
				in1 = y1[b];
				y1[b] = y0[b];
				y0[b] = y;
				v = y;
			}
			sg.at(buf, kWQ + 2, t) = v;
			const float out = biquad_step(e.filter, ef, sg.at(buf, kWX, t));
			ring.st((pos0 + t) & e.mask, out * e.feed_gain);
		}
	}
};

// ---- one stream, the pipeline's schedule executed serially (CPU test build; the device kernel below runs the same
// phase bodies).  Within an iteration the three phases run in an order that rotates with the iteration and the parallel
// phases walk their frames backwards: if the schedule is legal any order gives the reference's values, if it is not
// the parity tests see it.
template <int CT, bool CHAIN>
inline bool emulate_stream(const MixArgs& a, int tile, int lane)
{
	using Cx = Context<CT, CHAIN>;
	constexpr int W = staged_words(CHAIN), CAP = 64;
	Cx cx;
	if (!cx.setup(a, tile, lane)) {
		return false;
	}
	std::vector<float> mem(static_cast<size_t>(kBuffers) * W * CAP);
	const Stage<1, W, CAP> sg = {mem.data()};
	const int T = a.span_frames, nspans = (a.frames + T - 1) / T;
	const ReverbCoef& c = a.slot[Cx::RP].u.reverb;
	ShelfPair shelf[2];
	T60Pair t60[2];
	EqPair eqp;
	EqSingleEcho eqs;
	for (int p = 0; p < 2; ++p) {
		shelf[p].load(cx.st_rev, p);
		t60[p].load(cx.st_rev, p);
	}
	if (CHAIN) {
		eqp.load(cx.st_eq);
		eqs.load(cx.st_eq, cx.st_echo);
	}
	auto count_of = [&](int s) { return a.frames - s * T < T ? a.frames - s * T : T; };
	auto run_a = [&](int s) {
		for (int t = count_of(s) - 1; t >= 0; --t) {
			float x[CT];
			typename Cx::TapsA k;
			cx.load_input(a, s * T + t, x);
			cx.load_a(a, s * T + t, k);
			cx.phase_a(a, sg, s % kBuffers, t, x, k);
		}
	};
	auto run_b = [&](int s) {
		const int buf = s % kBuffers, count = count_of(s);
		for (int p = 1; p >= 0; --p) {
			t60[p].run(c, sg, buf, count, p);
			shelf[p].run(c, sg, buf, count, cx.ring_rev, cx.rev_off + s * T, p);
		}
		if (CHAIN) {
			eqs.run(a.slot[0].u.equalizer, a.slot[2].u.echo, sg, buf, count, cx.ring_echo, cx.echo_off + s * T);
			eqp.run(a.slot[0].u.equalizer, sg, buf, count);
		}
	};
	auto run_c = [&](int s) {
		for (int t = count_of(s) - 1; t >= 0; --t) {
			float x[CT];
			typename Cx::TapsC k;
			cx.load_input(a, s * T + t, x);
			cx.load_c(a, s * T + t, k);
			cx.phase_c(a, sg, s % kBuffers, t, s * T + t, x, k);
		}
	};
	for (int it = -1; it <= nspans; ++it) {
		for (int r = 0; r < 3; ++r) {
			const int which = (r + it + 3) % 3;
			if (which == 0 && it + 1 < nspans) {
				run_a(it + 1);
			} else if (which == 1 && it >= 0 && it < nspans) {
				run_b(it);
			} else if (which == 2 && it >= 1) {
				run_c(it - 1);
			}
		}
	}
	for (int p = 0; p < 2; ++p) {
		shelf[p].store(cx.st_rev, p);
		t60[p].store(cx.st_rev, p);
	}
	if (CHAIN) {
		eqp.store(cx.st_eq);
		eqs.store(cx.st_eq, cx.st_echo);
	}
	cx.store_scalars(a);
	cx.store_send_history(a, 0);
	for (int p = 0; p < (CHAIN ? 4 : 1); ++p) {
		cx.store_send_history(a, 1 + a.aux_index[p]);
	}
	return true;
}

#if defined(__CUDACC__)

__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 0;" ::: "memory"); }

// `bytes` (a multiple of 16) from a 16-byte aligned global address -> L2, one instruction.
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, unsigned bytes)
{
	asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Rows [p0, p0 + count) (positions modulo mask + 1) of the ring line starting at word `word0` of a tile's ring region.
__device__ __forceinline__ void prefetch_rows(const float* tile_ring, int word0, int p0, int count, int mask)
{
	const int start = p0 & mask;
	const int n1 = min(count, mask + 1 - start);
	bulk_prefetch_l2(tile_ring + static_cast<size_t>(word0 + start) * kLanes, static_cast<unsigned>(n1) * kLanes * 4U);
	if (n1 < count) {
		bulk_prefetch_l2(tile_ring + static_cast<size_t>(word0) * kLanes, static_cast<unsigned>(count - n1) * kLanes * 4U);
	}
}

template <int CT, bool CHAIN>
__device__ __noinline__ void exact_stream(const MixArgs& a, int tile, int lane)
{
	if (CHAIN) {
		mix_stream<CT, false, FxEqualizer, FxModDelay, FxEcho, FxReverb>(a, tile, lane, nullptr);
	} else {
		mix_stream<CT, false, FxReverb, FxNull, FxNull, FxNull>(a, tile, lane, nullptr);
	}
}

// SL: streams of the tile one CTA handles (32, 16 or 8).  With SL < 32 a tile is shared by 32 / SL CTAs -- on as many
// SMs -- and a warp covers FR = 32 / SL consecutive frames of those streams: lane = frame * SL + stream.  Streams are
// independent, so the split needs no communication.
template <int CT, int SL, bool CHAIN>
__global__ void __launch_bounds__(threads(CHAIN), 1) span_kernel(const __grid_constant__ MixArgs a)
{
	using Cx = Context<CT, CHAIN>;
	constexpr int NS = serial_warps(CHAIN), NP = kParallelWarps, FR = kLanes / SL, SPLIT = kLanes / SL;
	constexpr int W = staged_words(CHAIN), CAP = capacity(SL);
	extern __shared__ __align__(16) float dyn[];

	const int tile_slot = static_cast<int>(blockIdx.x) / SPLIT;
	const int tile = a.tiles ? static_cast<int>(a.tiles[tile_slot].tile) : a.tile_first + tile_slot;
	const int sl = static_cast<int>(threadIdx.x % kLanes) % SL;          // stream within this CTA's share
	const int f = static_cast<int>(threadIdx.x % kLanes) / SL;           // frame within the warp's FR frames
	const int lane = (static_cast<int>(blockIdx.x) % SPLIT) * SL + sl;   // stream within the tile
	const int w = threadIdx.x / kLanes;

	Cx cx;
	const bool ok = cx.setup(a, tile, lane);
	if (!__syncthreads_and(ok || !cx.io_ok)) {
		if (w == 0 && f == 0 && cx.io_ok) {
			exact_stream<CT, CHAIN>(a, tile, lane);
		}
		return;
	}
	const Stage<SL, W, CAP> sg = {dyn + sl};
	const int T = a.span_frames, nspans = (a.frames + T - 1) / T;
	const ReverbCoef& c = a.slot[Cx::RP].u.reverb;

	if (w >= NS) {
		// ---- parallel warps: C(it - 1), then A(it + 1) ----
		const int pw = w - NS;
		const float* tile_rev = cx.ring_rev.p - lane;
		for (int it = -1; it <= nspans; ++it) {
			// ring rows the NEXT iteration reads -> L2: phase C of span `it`, phase A of span `it + 2`
			if (pw == ((it + 1) & (NP - 1))) {
				const int l = threadIdx.x & 3, grp = (threadIdx.x % kLanes) >> 2;
				if (it >= 0 && it < nspans) {
					const int p0 = cx.rev_off + it * T, count = min(T, a.frames - it * T);
					if (grp == 0) {
						prefetch_rows(tile_rev, c.ring_base[0] + l * (c.mask[0] + 1), p0 - c.early_tap[l], count, c.mask[0]);
					} else if (grp == 1) {
						prefetch_rows(tile_rev, c.ring_base[1] + l * (c.mask[1] + 1), p0 - c.early_ap_off[l], count, c.mask[1]);
					} else if (grp == 2) {
						prefetch_rows(tile_rev, c.ring_base[2] + l * (c.mask[2] + 1), p0 - c.early_off[l], count, c.mask[2]);
					} else if (grp == 3) {
						prefetch_rows(tile_rev, c.ring_base[3] + l * (c.mask[3] + 1), p0 - c.late_ap_off[l], count, c.mask[3]);
					} else if (CHAIN && grp == 4 && l < 2) {
						const EchoCoef& e = a.slot[2].u.echo;
						prefetch_rows(cx.ring_echo.p - lane, 0, cx.echo_off + it * T - (l == 0 ? e.tap1 : e.tap2), count, e.mask);
					} else if (CHAIN && grp == 5 && l < 2) {
						const ModDelayCoef& m = a.slot[1].u.mod_delay;
						const int dmax = m.delay + static_cast<int>(m.depth) + 2, dmin = m.delay - static_cast<int>(m.depth) - 2;
						prefetch_rows(cx.ring_mod.p - lane, l * (m.mask + 1), cx.mod_off + it * T - dmax, count + dmax - dmin, m.mask);
					}
				}
				if (it + 2 < nspans) {
					const int p0 = cx.rev_off + (it + 2) * T, count = min(T, a.frames - (it + 2) * T);
					if (grp == 6) {
						prefetch_rows(tile_rev, c.ring_base[0] + l * (c.mask[0] + 1), p0 - c.late_tap[l], count, c.mask[0]);
					} else if (grp == 7) {
						prefetch_rows(tile_rev, c.ring_base[4] + l * (c.mask[4] + 1), p0 - c.late_off[l], count, c.mask[4]);
					} else if (CHAIN && grp == 4 && l == 2) {
						const EchoCoef& e = a.slot[2].u.echo;
						prefetch_rows(cx.ring_echo.p - lane, 0, cx.echo_off + (it + 2) * T - e.tap2, count, e.mask);
					}
				}
			}
			if (it >= 1) {
				const int s = it - 1, buf = s % kBuffers, first = s * T, count = min(T, a.frames - first);
				for (int t = pw * FR + f; t < count; t += 2 * NP * FR) {
					const int t2 = t + NP * FR;
					const bool two = t2 < count;
					float x0[CT], x1[CT];
					typename Cx::TapsC k0, k1;
					cx.load_input(a, first + t, x0);
					cx.load_c(a, first + t, k0);
					if (two) {
						cx.load_input(a, first + t2, x1);
						cx.load_c(a, first + t2, k1);
					}
					cx.phase_c(a, sg, buf, t, first + t, x0, k0);
					if (two) {
						cx.phase_c(a, sg, buf, t2, first + t2, x1, k1);
					}
				}
			}
			if (it + 1 < nspans) {
				const int s = it + 1, buf = s % kBuffers, first = s * T, count = min(T, a.frames - first);
				for (int t = pw * FR + f; t < count; t += 2 * NP * FR) {
					const int t2 = t + NP * FR;
					const bool two = t2 < count;
					float x0[CT], x1[CT];
					typename Cx::TapsA k0, k1;
					cx.load_input(a, first + t, x0);
					cx.load_a(a, first + t, k0);
					if (two) {
						cx.load_input(a, first + t2, x1);
						cx.load_a(a, first + t2, k1);
					}
					cx.phase_a(a, sg, buf, t, x0, k0);
					if (two) {
						cx.phase_a(a, sg, buf, t2, x1, k1);
					}
				}
			}
			bar_all();
		}
		if (f == 0) {
			if (pw == 0) {
				cx.store_scalars(a);
			} else if (pw == 1) {
				cx.store_send_history(a, 0);
			} else if (pw - 2 < (CHAIN ? 4 : 1)) {
				cx.store_send_history(a, 1 + a.aux_index[pw - 2]);
			}
		}
		return;
	}

	// ---- serial warps: B(it), one recurrence each ----
	const bool active = f == 0;
	if (w < 2) {
		ShelfPair r;
		r.load(cx.st_rev, w);
		for (int it = -1; it <= nspans; ++it) {
			if (active && it >= 0 && it < nspans) {
				r.run(c, sg, it % kBuffers, min(T, a.frames - it * T), cx.ring_rev, cx.rev_off + it * T, w);
			}
			bar_all();
		}
		if (active) {
			r.store(cx.st_rev, w);
		}
	} else if (w < 4) {
		T60Pair r;
		r.load(cx.st_rev, w - 2);
		for (int it = -1; it <= nspans; ++it) {
			if (active && it >= 0 && it < nspans) {
				r.run(c, sg, it % kBuffers, min(T, a.frames - it * T), w - 2);
			}
			bar_all();
		}
		if (active) {
			r.store(cx.st_rev, w - 2);
		}
	} else if (CHAIN && w == 4) {
		EqPair r;
		r.load(cx.st_eq);
		for (int it = -1; it <= nspans; ++it) {
			if (active && it >= 0 && it < nspans) {
				r.run(a.slot[0].u.equalizer, sg, it % kBuffers, min(T, a.frames - it * T));
			}
			bar_all();
		}
		if (active) {
			r.store(cx.st_eq);
		}
	} else if (CHAIN) {
		EqSingleEcho r;
		r.load(cx.st_eq, cx.st_echo);
		for (int it = -1; it <= nspans; ++it) {
			if (active && it >= 0 && it < nspans) {
				r.run(a.slot[0].u.equalizer, a.slot[2].u.echo, sg, it % kBuffers, min(T, a.frames - it * T), cx.ring_echo, cx.echo_off + it * T);
			}
			bar_all();
		}
		if (active) {
			r.store(cx.st_eq, cx.st_echo);
		}
	}
}

#endif // __CUDACC__

} // namespace span
} // namespace oalsfx

#endif
