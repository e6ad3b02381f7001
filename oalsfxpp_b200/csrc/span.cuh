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
OALSFX_CX int serial_warps(bool chain) { return chain ? 4 : 2; }
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
	// A read `delay` frames behind the frame position in iteration offset pr against a write `wofs` frames behind it
	// that lands in iteration offsets [pw0, pw1], on a ring of `len` positions.  zero_is_new: a read of the position
	// written in the same frame sees the new value (the reference writes first), else the one written `len` frames earlier.
	void pair(int pr, int delay, int pw0, int pw1, int wofs, int len, bool zero_is_new)
	{
		int k = ((delay - wofs) % len + len) % len;  // frames between the write of a position and this read of it
		if (k == 0 && !zero_is_new) {
			k = len;
		}
		// read after write: the write has landed by iteration span(n - k) + pw1, the read happens in span(n) + pr; the
		// minimum over n of the span difference is floor(k / t)
		ok = ok && k / t > pw1 - pr;
		// write after read: the position is next written len - k frames after the read, not before span(..) + pw0
		ok = ok && (len - k) / t > pr - pw0;
	}
	// two writers of one ring: the writes to a position keep their order
	void writers(int p10, int p11, int o1, int p20, int p21, int o2, int len)
	{
		const int d = ((o2 - o1) % len + len) % len;   // writer 2 reaches a position d frames after writer 1
		ok = ok && d / t > p11 - p20;
		ok = ok && (len - d) / t > p21 - p10;
	}
};

// When the ring accesses of a span's phases take effect, as iteration offsets from the span index:
//   ra / rc  reads of phase A / C;  wb / wc  writes of phase B / C: [first, last] iteration they may land in;
//   mr / mw  the chorus ring's reads / writes
// span_kernel (loads and stores at the point of use):        A one iteration before B, C one after
// span_bulk_kernel (bulk copies issued ahead, stores behind): A's rows fetched two iterations before B, C's in B's own
//   iteration; B's row stores are issued when B is done and have landed by the end of the next iteration; C's rows
//   (C runs one iteration after B) are sent off at the start of the iteration after C and have landed by the end of the
//   one after that; the chorus taps are requested one iteration before C
struct Timing { int ra, rc, wb0, wb1, wc0, wc1, mr, mw; };
constexpr Timing kDirectTiming = {kPhA, kPhC, kPhB, kPhB, kPhC, kPhC, kPhC, kPhC};
constexpr Timing kBulkTiming = {-2, 0, 0, 1, 1, 3, 0, 1};
constexpr int kBulkFrames = 16;           // span length of span_bulk_kernel (shared memory holds every ring row of a span)

inline bool reverb_legal(const ReverbCoef& c, int t, const Timing& tm = kDirectTiming)
{
	if (c.mod_depth != 0.0F) {
		return false;
	}
	Legality g;
	g.t = t;
	const int len0 = c.mask[0] + 1;
	for (int l = 0; l < 4; ++l) {
		// main line: shelves write at the position (B), the early scatter feeds it late_feed_tap behind (C)
		g.pair(tm.rc, c.early_tap[l], tm.wb0, tm.wb1, 0, len0, true);
		g.pair(tm.rc, c.early_tap[l], tm.wc0, tm.wc1, c.late_feed_tap, len0, false);
		g.pair(tm.ra, c.late_tap[l], tm.wb0, tm.wb1, 0, len0, true);
		g.pair(tm.ra, c.late_tap[l], tm.wc0, tm.wc1, c.late_feed_tap, len0, true);
		g.pair(tm.rc, c.early_ap_off[l], tm.wc0, tm.wc1, 0, c.mask[1] + 1, false);
		g.pair(tm.rc, c.early_off[l], tm.wc0, tm.wc1, 0, c.mask[2] + 1, true);
		g.pair(tm.rc, c.late_ap_off[l], tm.wc0, tm.wc1, 0, c.mask[3] + 1, false);
		g.pair(tm.ra, c.late_off[l], tm.wc0, tm.wc1, 0, c.mask[4] + 1, true);
	}
	g.writers(tm.wb0, tm.wb1, 0, tm.wc0, tm.wc1, c.late_feed_tap, len0);
	return g.ok;
}

inline bool mod_delay_legal(const ModDelayCoef& c, int t, const Timing& tm = kDirectTiming)
{
	Legality g;
	g.t = t;
	// the LFO keeps the delay within delay +- depth (oalsfxpp.cpp:4246-4276); two frames of slack for the rounding
	const int dmin = c.delay - static_cast<int>(c.depth) - 2, dmax = c.delay + static_cast<int>(c.depth) + 2;
	if (dmin < 1 || c.lfo_range <= kMaxBlockFrames) {
		return false;
	}
	g.pair(tm.mr, dmin, tm.mw, tm.mw, 0, c.mask + 1, false);
	g.pair(tm.mr, dmax, tm.mw, tm.mw, 0, c.mask + 1, false);
	return g.ok && dmax < c.mask + 1;
}

inline bool echo_legal(const EchoCoef& c, int t, const Timing& tm = kDirectTiming)
{
	Legality g;
	g.t = t;
	g.pair(tm.rc, c.tap1, tm.wb0, tm.wb1, 0, c.mask + 1, false);
	g.pair(tm.rc, c.tap2, tm.wb0, tm.wb1, 0, c.mask + 1, false);
	g.pair(tm.ra, c.tap2, tm.wb0, tm.wb1, 0, c.mask + 1, false);
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

// span_bulk_kernel's fixed span length if its timing is legal for these coefficient blocks, or 0.
inline int plan_bulk_frames(const MixArgs& a, bool chain)
{
	bool ok = reverb_legal(a.slot[chain ? 3 : 0].u.reverb, kBulkFrames, kBulkTiming);
	if (chain) {
		ok = ok && mod_delay_legal(a.slot[1].u.mod_delay, kBulkFrames, kBulkTiming) && echo_legal(a.slot[2].u.echo, kBulkFrames, kBulkTiming);
	}
	return ok ? kBulkFrames : 0;
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

// ---- where the ring stores of phases B and C go ------------------------------------------------------------------------
// Rows of phase C (each = one ring line): early all-pass 0..3, early line 4..7, main line feed 8..11 (late_feed_tap
// behind the position), late all-pass 12..15, late line 16..19.  Rows of phase B: main line 0..3, echo ring 4.
constexpr int kOutEap = 0, kOutEline = 4, kOutFeed = 8, kOutLap = 12, kOutLline = 16, kOutRowsC = 20;
constexpr int kOutMain = 0, kOutEcho = 4, kOutRowsB = 5;

// straight to the rings
struct RingSinkC {
	const ReverbCoef& c;
	LaneMem ring;
	int pos;
	OALSFX_HD void put(int row, float v) const
	{
		const int l = row & 3;
		const int r = row < kOutEline ? 1 : row < kOutFeed ? 2 : row < kOutLap ? 0 : row < kOutLline ? 3 : 4;
		const int at = (row >= kOutFeed && row < kOutLap) ? pos - c.late_feed_tap : pos;
		ring.st(c.ring_base[r] + l * (c.mask[r] + 1) + (at & c.mask[r]), v);
	}
};
struct RingSinkB {
	const ReverbCoef& c;
	LaneMem ring_rev, ring_echo;
	int echo_mask;
	int rev_pos0, echo_pos0;      // ring positions of the span's first frame
	OALSFX_HD void put(int row, int t, float v) const
	{
		if (row < kOutEcho) {
			ring_rev.st(c.ring_base[0] + row * (c.mask[0] + 1) + ((rev_pos0 + t) & c.mask[0]), v);
		} else {
			ring_echo.st((echo_pos0 + t) & echo_mask, v);
		}
	}
};
// into shared-memory row buffers [row][frame][lane] that bulk copies move to the rings (span_bulk_kernel)
template <int T>
struct StageSinkC {
	float* base;                  // + frame * 32 + lane
	OALSFX_HD void put(int row, float v) const { base[row * (T * kLanes)] = v; }
};
// ... where phase C's store rows replace the tap rows the same thread has just read: early all-pass and early line
// rows in place of their taps, the main-line feed in place of the early taps, late all-pass in place, the late line
// in place of the echo taps + two spare rows (buffer rows: early 0..3, early all-pass 4..7, early line 8..11, late
// all-pass 12..15, echo taps 16, 17)
OALSFX_CX int buffer_row_of_out(int out_row) { return out_row < kOutFeed ? out_row + 4 : out_row < kOutLap ? out_row - kOutFeed : out_row; }
OALSFX_CX int out_of_buffer_row(int row) { return row < 4 ? kOutFeed + row : row < 12 ? row - 4 : row; }
template <int T>
struct InPlaceSinkC {
	float* base;                  // + frame * 32 + lane
	OALSFX_HD void put(int row, float v) const { base[buffer_row_of_out(row) * (T * kLanes)] = v; }
};
template <int T>
struct StageSinkB {
	float* base;                  // + lane
	OALSFX_HD void put(int row, int t, float v) const { base[(row * T + t) * kLanes] = v; }
};

// ---- per-stream context ------------------------------------------------------------------------------------------
template <int CT, bool CHAIN>
struct Context {
	using R = FxReverbTail;
	static constexpr int RP = CHAIN ? 3 : 0;  // slot position of the reverb
	const float* src;
	float* dst;
	bool io_ok;
	LaneMem ring_rev = {nullptr}, ring_mod = {nullptr}, ring_echo = {nullptr};
	uint32_t *st_rev = nullptr, *st_eq = nullptr, *st_mod = nullptr, *st_echo = nullptr, *ss = nullptr;
	int32_t rev_off = 0, mod_off = 0, echo_off = 0;  // ring offsets at the start of the block
	int32_t mod_ph[2] = {0, 0};               // chorus LFO phases at the start of the block
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
	struct TapsC { float early[4], eap[4], eline[4], lap[4], echo[2]; float t60[4], eq[3]; };  // ring taps + phase B's results
	struct TapsMod { float v[2]; int32_t pos; };

	OALSFX_HD void load_c(const MixArgs& a, int n, TapsC& k, TapsMod& md) const
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
			load_mod(a, n, md);
			const EchoCoef& e = a.slot[2].u.echo;
			k.echo[0] = ring_ld(ring_echo, (echo_off + n - e.tap1) & e.mask);
			k.echo[1] = ring_ld(ring_echo, (echo_off + n - e.tap2) & e.mask);
		}
	}

	// the chorus / flanger taps: LFO-modulated positions, always read in place
	OALSFX_HD void load_mod(const MixArgs& a, int n, TapsMod& k) const
	{
		const ModDelayCoef& m = a.slot[1].u.mod_delay;
		const int32_t mlen = m.mask + 1;
		k.pos = mod_off + n;
		OALSFX_UNROLL
		for (int side = 0; side < 2; ++side) {
			int32_t ph = mod_ph[side] + n;         // n <= 2048 < lfo_range (host-checked): at most one wrap
			ph = (ph >= m.lfo_range ? ph - m.lfo_range : ph);
			const int32_t d = FxModDelay::lfo_delay(m, ph);
			k.v[side] = ring_ld(ring_mod, side * mlen + ((k.pos - d) & m.mask));
		}
	}

	// Taps from row buffers [row][frame][LANES] filled by bulk copies (span_bulk_kernel); `rows` points at this
	// thread's element of row 0.  Row order: see tap_row_a / tap_row_c.
	template <int T, int LANES>
	OALSFX_HD void fetch_a(const float* rows, TapsA& k) const
	{
		OALSFX_UNROLL
		for (int l = 0; l < 4; ++l) {
			k.late[l] = rows[(0 + l) * (T * LANES)];
			k.lline[l] = rows[(4 + l) * (T * LANES)];
		}
		if (CHAIN) {
			k.echo2 = rows[8 * (T * LANES)];
		}
	}
	template <int T, int LANES>
	OALSFX_HD void fetch_c(const float* rows, TapsC& k) const
	{
		OALSFX_UNROLL
		for (int l = 0; l < 4; ++l) {
			k.early[l] = rows[(0 + l) * (T * LANES)];
			k.eap[l] = rows[(4 + l) * (T * LANES)];
			k.eline[l] = rows[(8 + l) * (T * LANES)];
			k.lap[l] = rows[(12 + l) * (T * LANES)];
		}
		if (CHAIN) {
			k.echo[0] = rows[16 * (T * LANES)];
			k.echo[1] = rows[17 * (T * LANES)];
		}
	}

	// vector_allpass_x with the taps already read (FxReverbT::vector_allpass2, fx_reverb.cuh); the four ring
	// stores go to rows row0 .. row0 + 3 of the sink
	template <class Sink>
	OALSFX_HD static void allpass(const ReverbCoef& c, F2& va, F2& vb, const float* tp, const Sink& out, int row0)
	{
		const F2 ta = f2(tp[0], tp[1]), tb = f2(tp[2], tp[3]);
		const F2 ina = va, inb = vb;
		va = ta - (ina * c.ap_feed_coeff);
		vb = tb - (inb * c.ap_feed_coeff);
		F2 fa = ina + (va * c.ap_feed_coeff);
		F2 fb = inb + (vb * c.ap_feed_coeff);
		R::scatter2(fa, fb, c.mix_x, c.mix_y);
		out.put(row0 + 0, f2_lo(fa));
		out.put(row0 + 1, f2_hi(fa));
		out.put(row0 + 2, f2_lo(fb));
		out.put(row0 + 3, f2_hi(fb));
	}

	// phase B's results of frame t (staged words): read before any of phase C's stores, so that the frames a thread
	// processes back to back do not wait for each other's shared-memory traffic
	template <class S>
	OALSFX_HD void fetch_staged(const S& sg, int buf, int t, TapsC& k) const
	{
		OALSFX_UNROLL
		for (int l = 0; l < 4; ++l) {
			k.t60[l] = sg.at(buf, kWL + l, t);
		}
		if (CHAIN) {
			OALSFX_UNROLL
			for (int q = 0; q < 3; ++q) {
				k.eq[q] = sg.at(buf, kWQ + q, t);
			}
		}
	}

	template <class Sink>
	OALSFX_HD void phase_c(const MixArgs& a, int n, const float* x, const TapsC& k, const TapsMod& md, const Sink& out) const
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
				const float q0 = k.eq[0], q1 = k.eq[1], q3 = k.eq[2];
				pan_add<CT, true>(acc, CT, q.gains[0], q0);
				pan_add<CT, true>(acc, CT, q.gains[1], q1);
				pan_add<CT, true>(acc, CT, q.gains[3], q3);
			}
			{
				// chorus / flanger (FxModDelay::step; every delay >= the span, so never the sample just written)
				const ModDelayCoef& m = a.slot[1].u.mod_delay;
				float wet[kWetChannels];
				encode_wet<CT>(a.aux[1], x, wet);
				const int32_t mlen = m.mask + 1, mpos = md.pos & m.mask;
				float tt[2];
				OALSFX_UNROLL
				for (int side = 0; side < 2; ++side) {
					tt[side] = md.v[side] * m.feedback;
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
		float out8[8];
		// early reflections (the EARLY half of FxReverbT::body)
		F2 fa = f2(k.early[0], k.early[1]) * f2(c.early_tap_coeff[0], c.early_tap_coeff[1]);
		F2 fb = f2(k.early[2], k.early[3]) * f2(c.early_tap_coeff[2], c.early_tap_coeff[3]);
		allpass(c, fa, fb, k.eap, out, kOutEap);
		out.put(kOutEline + 0, f2_hi(fb));   // delay_line_in4_rev: line j receives f[3 - j]
		out.put(kOutEline + 1, f2_lo(fb));
		out.put(kOutEline + 2, f2_hi(fa));
		out.put(kOutEline + 3, f2_lo(fa));
		fa = fa + (f2(k.eline[0], k.eline[1]) * f2(c.early_coeff[0], c.early_coeff[1]));
		fb = fb + (f2(k.eline[2], k.eline[3]) * f2(c.early_coeff[2], c.early_coeff[3]));
		out8[0] = f2_lo(fa);
		out8[1] = f2_hi(fa);
		out8[2] = f2_lo(fb);
		out8[3] = f2_hi(fb);
		{
			F2 ra = fa, rb = fb;
			R::scatter2_reversed(ra, rb, c.mix_x, c.mix_y);
			out.put(kOutFeed + 0, f2_hi(rb));    // main line, late_feed_tap behind the position
			out.put(kOutFeed + 1, f2_lo(rb));
			out.put(kOutFeed + 2, f2_hi(ra));
			out.put(kOutFeed + 3, f2_lo(ra));
		}
		// late reverb after the T60 filters
		fa = f2(k.t60[0], k.t60[1]);
		fb = f2(k.t60[2], k.t60[3]);
		allpass(c, fa, fb, k.lap, out, kOutLap);
		out8[4] = f2_lo(fa);
		out8[5] = f2_hi(fa);
		out8[6] = f2_lo(fb);
		out8[7] = f2_hi(fb);
		{
			F2 ra = fa, rb = fb;
			R::scatter2_reversed(ra, rb, c.mix_x, c.mix_y);
			out.put(kOutLline + 0, f2_hi(rb));
			out.put(kOutLline + 1, f2_lo(rb));
			out.put(kOutLline + 2, f2_hi(ra));
			out.put(kOutLline + 3, f2_lo(ra));
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

// Master shelves of reverb lines as packed pairs (2p, 2p + 1) -> main delay line (reverb_input_stage, fx_reverb.cuh).
// NPAIR = 2: one thread runs both pairs (a whole tile per CTA: every lane has a stream); NPAIR = 1: pair p0 only (a tile
// shared by several CTAs: the lanes that would idle take the second pair -- same instructions, half as many per step).
template <int NPAIR>
struct Shelves {
	using R = FxReverbTail;
	Biquad2 lp[NPAIR], hp[NPAIR];
	OALSFX_HD void load(const uint32_t* st, int p0)
	{
		OALSFX_UNROLL
		for (int i = 0; i < NPAIR; ++i) {
			const int p = p0 + i;
			lp[i].load(st + (R::kWLp + (2 * p) * 4) * kLanes, st + (R::kWLp + (2 * p + 1) * 4) * kLanes);
			hp[i].load(st + (R::kWHp + (2 * p) * 4) * kLanes, st + (R::kWHp + (2 * p + 1) * 4) * kLanes);
		}
	}
	OALSFX_HD void store(uint32_t* st, int p0) const
	{
		OALSFX_UNROLL
		for (int i = 0; i < NPAIR; ++i) {
			const int p = p0 + i;
			lp[i].store(st + (R::kWLp + (2 * p) * 4) * kLanes, st + (R::kWLp + (2 * p + 1) * 4) * kLanes);
			hp[i].store(st + (R::kWHp + (2 * p) * 4) * kLanes, st + (R::kWHp + (2 * p + 1) * 4) * kLanes);
		}
	}
	template <class S, class Sink>
	OALSFX_HD void run(const ReverbCoef& c, const S& sg, int buf, int count, const Sink& out, int p0)
	{
#if defined(__CUDA_ARCH__)
#pragma unroll 2
#endif
		for (int t = 0; t < count; ++t) {
			OALSFX_UNROLL
			for (int i = 0; i < NPAIR; ++i) {
				const int p = p0 + i;
				F2 v = f2(sg.at(buf, kWA + 2 * p, t), sg.at(buf, kWA + 2 * p + 1, t));
				v = lp[i].step(c.lp, v);
				if (c.is_eax) {
					v = hp[i].step(c.hp, v);
				}
				out.put(kOutMain + 2 * p, t, f2_lo(v));
				out.put(kOutMain + 2 * p + 1, t, f2_hi(v));
			}
		}
	}
};

// late_t60_filter of reverb lines as packed pairs, in place (oalsfxpp.cpp:7691-7719; FxReverbT::body).  NPAIR as above.
template <int NPAIR>
struct T60s {
	using R = FxReverbTail;
	F2 p[NPAIR][2][2];   // [pair][section][x1 / y1]
	OALSFX_HD void load(const uint32_t* st, int p0)
	{
		OALSFX_UNROLL
		for (int i = 0; i < NPAIR; ++i) {
			const int h = p0 + i;
			OALSFX_UNROLL
			for (int q = 0; q < 4; ++q) {
				p[i][q >> 1][q & 1] = f2(word_as_float(st[(R::kWT60 + (2 * h) * 4 + q) * kLanes]), word_as_float(st[(R::kWT60 + (2 * h + 1) * 4 + q) * kLanes]));
			}
		}
	}
	OALSFX_HD void store(uint32_t* st, int p0) const
	{
		OALSFX_UNROLL
		for (int i = 0; i < NPAIR; ++i) {
			const int h = p0 + i;
			OALSFX_UNROLL
			for (int q = 0; q < 4; ++q) {
				st[(R::kWT60 + (2 * h) * 4 + q) * kLanes] = float_as_word(f2_lo(p[i][q >> 1][q & 1]));
				st[(R::kWT60 + (2 * h + 1) * 4 + q) * kLanes] = float_as_word(f2_hi(p[i][q >> 1][q & 1]));
			}
		}
	}
	template <class S>
	OALSFX_HD void run(const ReverbCoef& c, const S& sg, int buf, int count, int p0)
	{
#if defined(__CUDA_ARCH__)
#pragma unroll 2
#endif
		for (int t = 0; t < count; ++t) {
			OALSFX_UNROLL
			for (int i = 0; i < NPAIR; ++i) {
				const int j = 2 * (p0 + i);
				const F2 in = f2(sg.at(buf, kWL + j, t), sg.at(buf, kWL + j + 1, t));
				const F2 o1 = (f2(c.t60_lf[j][0], c.t60_lf[j + 1][0]) * in) + (f2(c.t60_lf[j][1], c.t60_lf[j + 1][1]) * p[i][0][0]) +
					(f2(c.t60_lf[j][2], c.t60_lf[j + 1][2]) * p[i][0][1]);
				p[i][0][0] = in;
				p[i][0][1] = o1;
				const F2 o2 = (f2(c.t60_hf[j][0], c.t60_hf[j + 1][0]) * o1) + (f2(c.t60_hf[j][1], c.t60_hf[j + 1][1]) * p[i][1][0]) +
					(f2(c.t60_hf[j][2], c.t60_hf[j + 1][2]) * p[i][1][1]);
				p[i][1][0] = o1;
				p[i][1][1] = o2;
				const F2 out = f2(c.t60_mid[j], c.t60_mid[j + 1]) * o2;
				sg.at(buf, kWL + j, t) = f2_lo(out);
				sg.at(buf, kWL + j + 1, t) = f2_hi(out);
			}
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
	template <class S, class Sink>
	OALSFX_HD void run(const EqualizerCoef& c, const EchoCoef& e, const S& sg, int buf, int count, const Sink& out)
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
			const float o = biquad_step(e.filter, ef, sg.at(buf, kWX, t));
			out.put(kOutEcho, t, o * e.feed_gain);
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
	Shelves<2> shelves;
	T60s<2> t60s;
	EqPair eqp;
	EqSingleEcho eqs;
	shelves.load(cx.st_rev, 0);
	t60s.load(cx.st_rev, 0);
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
		const RingSinkB out = {c, cx.ring_rev, cx.ring_echo, CHAIN ? a.slot[2].u.echo.mask : 0, cx.rev_off + s * T, cx.echo_off + s * T};
		t60s.run(c, sg, buf, count, 0);
		shelves.run(c, sg, buf, count, out, 0);
		if (CHAIN) {
			eqs.run(a.slot[0].u.equalizer, a.slot[2].u.echo, sg, buf, count, out);
			eqp.run(a.slot[0].u.equalizer, sg, buf, count);
		}
	};
	auto run_c = [&](int s) {
		for (int t = count_of(s) - 1; t >= 0; --t) {
			float x[CT];
			typename Cx::TapsC k;
			typename Cx::TapsMod md;
			cx.load_input(a, s * T + t, x);
			cx.load_c(a, s * T + t, k, md);
			cx.fetch_staged(sg, s % kBuffers, t, k);
			const RingSinkC out = {c, cx.ring_rev, cx.rev_off + s * T + t};
			cx.phase_c(a, s * T + t, x, k, md, out);
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
	shelves.store(cx.st_rev, 0);
	t60s.store(cx.st_rev, 0);
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

// ---- span_bulk_kernel: which ring rows its bulk copies move -----------------------------------------------------------------
// A row = one ring line over the frames of a span: `behind` frames behind the span's ring position, on the reverb's
// ring region (ring 0) or the echo's (ring 1), line starting at word `word0`, positions modulo mask + 1.
struct RowJob { int ring, word0, mask, behind; };
OALSFX_CX int tap_rows_c(bool chain) { return chain ? 18 : 16; }
OALSFX_CX int tap_rows_a(bool chain) { return chain ? 9 : 8; }
OALSFX_CX int out_rows_b(bool chain) { return chain ? kOutRowsB : 4; }

template <bool CHAIN>
OALSFX_HD RowJob reverb_row(const MixArgs& a, int r, int l, int behind)
{
	const ReverbCoef& c = a.slot[CHAIN ? 3 : 0].u.reverb;
	return RowJob{0, c.ring_base[r] + l * (c.mask[r] + 1), c.mask[r], behind};
}
// taps of phase C: early 0..3, early all-pass 4..7, early line 8..11, late all-pass 12..15, echo tap1 16, tap2 17
template <bool CHAIN>
OALSFX_HD RowJob tap_row_c(const MixArgs& a, int row)
{
	const ReverbCoef& c = a.slot[CHAIN ? 3 : 0].u.reverb;
	const int l = row & 3;
	switch (row >> 2) {
	case 0: return reverb_row<CHAIN>(a, 0, l, c.early_tap[l]);
	case 1: return reverb_row<CHAIN>(a, 1, l, c.early_ap_off[l]);
	case 2: return reverb_row<CHAIN>(a, 2, l, c.early_off[l]);
	case 3: return reverb_row<CHAIN>(a, 3, l, c.late_ap_off[l]);
	default: return RowJob{1, 0, a.slot[2].u.echo.mask, l == 0 ? a.slot[2].u.echo.tap1 : a.slot[2].u.echo.tap2};
	}
}
// taps of phase A: late 0..3, late line 4..7, echo tap2 8
template <bool CHAIN>
OALSFX_HD RowJob tap_row_a(const MixArgs& a, int row)
{
	const ReverbCoef& c = a.slot[CHAIN ? 3 : 0].u.reverb;
	const int l = row & 3;
	switch (row >> 2) {
	case 0: return reverb_row<CHAIN>(a, 0, l, c.late_tap[l]);
	case 1: return reverb_row<CHAIN>(a, 4, l, c.late_off[l]);
	default: return RowJob{1, 0, a.slot[2].u.echo.mask, a.slot[2].u.echo.tap2};
	}
}
// stores of phase C (kOut* rows) and of phase B (main line 0..3, echo 4)
template <bool CHAIN>
OALSFX_HD RowJob out_row_c(const MixArgs& a, int row)
{
	const ReverbCoef& c = a.slot[CHAIN ? 3 : 0].u.reverb;
	const int l = row & 3;
	switch (row >> 2) {
	case 0: return reverb_row<CHAIN>(a, 1, l, 0);
	case 1: return reverb_row<CHAIN>(a, 2, l, 0);
	case 2: return reverb_row<CHAIN>(a, 0, l, c.late_feed_tap);
	case 3: return reverb_row<CHAIN>(a, 3, l, 0);
	default: return reverb_row<CHAIN>(a, 4, l, 0);
	}
}
template <bool CHAIN>
OALSFX_HD RowJob out_row_b(const MixArgs& a, int row)
{
	if (row < kOutEcho) {
		return reverb_row<CHAIN>(a, 0, row, 0);
	}
	return RowJob{1, 0, a.slot[2].u.echo.mask, 0};
}

// One stream through span_bulk_kernel's schedule, executed serially (CPU test build).  The bulk copies are modelled at
// the extremes of when they may take effect: `late_stores` -- loads at their issue point, stores at the END of the
// iteration they are issued in (worst case for read-after-write); otherwise stores at their issue point and loads at
// the end of the iteration they are issued in (worst case for write-after-read).
template <int CT, bool CHAIN>
inline bool emulate_stream_bulk(const MixArgs& a, int tile, int lane, bool late_stores)
{
	using Cx = Context<CT, CHAIN>;
	constexpr int W = staged_words(CHAIN), T = kBulkFrames, RC = tap_rows_c(CHAIN), RA = tap_rows_a(CHAIN), RB = out_rows_b(CHAIN);
	Cx cx;
	if (!cx.setup(a, tile, lane)) {
		return false;
	}
	std::vector<float> mem(static_cast<size_t>(kBuffers) * W * T), tap_c(2 * RC * T), tap_a(RA * T), out_c(kOutRowsC * T), out_b(RB * T);
	const Stage<1, W, T> sg = {mem.data()};
	const int nspans = (a.frames + T - 1) / T;
	const ReverbCoef& c = a.slot[Cx::RP].u.reverb;
	Shelves<2> shelves;
	T60s<2> t60s;
	EqPair eqp;
	EqSingleEcho eqs;
	shelves.load(cx.st_rev, 0);
	t60s.load(cx.st_rev, 0);
	if (CHAIN) {
		eqp.load(cx.st_eq);
		eqs.load(cx.st_eq, cx.st_echo);
	}
	auto count_of = [&](int s) { return a.frames - s * T < T ? a.frames - s * T : T; };
	std::vector<typename Cx::TapsMod> mod_taps(2 * T);   // the chorus taps are requested together with phase C's rows
	auto move_row = [&](const RowJob& j, int s, float* row, bool load) {
		const LaneMem& ring = j.ring == 0 ? cx.ring_rev : cx.ring_echo;
		const int pos0 = (j.ring == 0 ? cx.rev_off : cx.echo_off) + s * T - j.behind;
		for (int t = 0; t < count_of(s); ++t) {
			if (load) {
				row[t] = ring.ld(j.word0 + ((pos0 + t) & j.mask));
			} else {
				ring.st(j.word0 + ((pos0 + t) & j.mask), row[t]);
			}
		}
	};
	auto load_tap_c = [&](int s) {
		if (s >= 0 && s < nspans) {
			for (int r = 0; r < RC; ++r) {
				move_row(tap_row_c<CHAIN>(a, r), s, &tap_c[((s & 1) * RC + r) * T], true);
			}
			if (CHAIN) {
				for (int t = 0; t < count_of(s); ++t) {
					cx.load_mod(a, s * T + t, mod_taps[(s & 1) * T + t]);
				}
			}
		}
	};
	auto load_tap_a = [&](int s) {
		if (s >= 0 && s < nspans) {
			for (int r = 0; r < RA; ++r) {
				move_row(tap_row_a<CHAIN>(a, r), s, &tap_a[r * T], true);
			}
		}
	};
	// The row stores are issued by the warp that produced the rows, right after the phase, and may land any time up
	// to the end of the following iteration.
	auto land_c = [&](int s, const float* rows) {
		if (s >= 0 && s < nspans) {
			for (int r = 0; r < kOutRowsC; ++r) {
				move_row(out_row_c<CHAIN>(a, r), s, const_cast<float*>(rows) + r * T, false);
			}
		}
	};
	auto land_b = [&](int s, const float* rows) {
		if (s >= 0 && s < nspans) {
			for (int r = 0; r < RB; ++r) {
				move_row(out_row_b<CHAIN>(a, r), s, const_cast<float*>(rows) + r * T, false);
			}
		}
	};
	std::vector<float> pend_c[3] = {std::vector<float>(kOutRowsC * T), std::vector<float>(kOutRowsC * T), std::vector<float>(kOutRowsC * T)};
	std::vector<float> pend_b[2] = {std::vector<float>(RB * T), std::vector<float>(RB * T)};
	auto run_a = [&](int s) {
		for (int t = count_of(s) - 1; t >= 0; --t) {
			float x[CT];
			typename Cx::TapsA k;
			cx.load_input(a, s * T + t, x);
			cx.template fetch_a<T, 1>(&tap_a[t], k);
			cx.phase_a(a, sg, s % kBuffers, t, x, k);
		}
	};
	auto run_b = [&](int s) {
		const int buf = s % kBuffers, count = count_of(s);
		struct Sink1 { float* base; void put(int row, int t, float v) const { base[row * T + t] = v; } } out1 = {out_b.data()};
		t60s.run(c, sg, buf, count, 0);
		shelves.run(c, sg, buf, count, out1, 0);
		if (CHAIN) {
			eqs.run(a.slot[0].u.equalizer, a.slot[2].u.echo, sg, buf, count, out1);
			eqp.run(a.slot[0].u.equalizer, sg, buf, count);
		}
	};
	auto run_c = [&](int s) {
		for (int t = count_of(s) - 1; t >= 0; --t) {
			float x[CT];
			typename Cx::TapsC k;
			cx.load_input(a, s * T + t, x);
			cx.template fetch_c<T, 1>(&tap_c[(s & 1) * RC * T + t], k);
			cx.fetch_staged(sg, s % kBuffers, t, k);
			struct Sink1 { float* base; void put(int row, float v) const { base[row * T] = v; } } out1 = {&out_c[t]};
			cx.phase_c(a, s * T + t, x, k, mod_taps[(s & 1) * T + t], out1);
		}
	};
	load_tap_a(0);
	for (int it = -1; it <= nspans + 2; ++it) {
		if (late_stores) {
			load_tap_c(it);      // loads at the point they are issued
		}
		if (it + 1 < nspans) {
			run_a(it + 1);
		}
		if (late_stores) {
			load_tap_a(it + 2);  // issued once phase A has consumed the buffer
		}
		auto do_b = [&]() {
			if (it >= 0 && it < nspans) {
				run_b(it);
				if (late_stores) {
					pend_b[it & 1] = out_b;
				} else {
					land_b(it, out_b.data());
				}
			}
		};
		auto do_c = [&]() {
			if (it >= 1 && it - 1 < nspans) {
				run_c(it - 1);
				if (late_stores) {
					pend_c[(it + 3) % 3] = out_c;
				} else {
					land_c(it - 1, out_c.data());
				}
			}
		};
		if ((it & 1) != 0) {
			do_b();
			do_c();
		} else {
			do_c();
			do_b();
		}
		if (late_stores) {
			// landing as late as allowed: phase B's rows of span it - 1 (issued in the previous iteration), phase C's rows
			// of span it - 3 (produced in iteration it - 2, sent off in it - 1)
			land_b(it - 1, pend_b[(it - 1) & 1].data());
			land_c(it - 3, pend_c[(it - 2 + 3) % 3].data());
		} else {
			load_tap_a(it + 2);  // loads as late as they can land
			load_tap_c(it);
		}
	}
	shelves.store(cx.st_rev, 0);
	t60s.store(cx.st_rev, 0);
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
					typename Cx::TapsMod m0, m1;
					cx.load_input(a, first + t, x0);
					cx.load_c(a, first + t, k0, m0);
					cx.fetch_staged(sg, buf, t, k0);
					if (two) {
						cx.load_input(a, first + t2, x1);
						cx.load_c(a, first + t2, k1, m1);
						cx.fetch_staged(sg, buf, t2, k1);
					}
					cx.phase_c(a, first + t, x0, k0, m0, RingSinkC{c, cx.ring_rev, cx.rev_off + first + t});
					if (two) {
						cx.phase_c(a, first + t2, x1, k1, m1, RingSinkC{c, cx.ring_rev, cx.rev_off + first + t2});
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

	// ---- serial warps: B(it) ----
	// the reverb's recurrences: both line pairs per thread, or -- a tile shared among CTAs -- one pair per lane group
	constexpr int NPAIR = SL == kLanes ? 2 : 1;
	const bool active = f == 0, pair_active = f < 2 / NPAIR;
	const int p0 = NPAIR == 2 ? 0 : f;
	auto sink = [&](int it) {
		return RingSinkB{c, cx.ring_rev, cx.ring_echo, CHAIN ? a.slot[2].u.echo.mask : 0, cx.rev_off + it * T, cx.echo_off + it * T};
	};
	if (w == 0) {
		Shelves<NPAIR> r;
		r.load(cx.st_rev, pair_active ? p0 : 0);
		for (int it = -1; it <= nspans; ++it) {
			if (pair_active && it >= 0 && it < nspans) {
				r.run(c, sg, it % kBuffers, min(T, a.frames - it * T), sink(it), p0);
			}
			bar_all();
		}
		if (pair_active) {
			r.store(cx.st_rev, p0);
		}
	} else if (w == 1) {
		T60s<NPAIR> r;
		r.load(cx.st_rev, pair_active ? p0 : 0);
		for (int it = -1; it <= nspans; ++it) {
			if (pair_active && it >= 0 && it < nspans) {
				r.run(c, sg, it % kBuffers, min(T, a.frames - it * T), p0);
			}
			bar_all();
		}
		if (pair_active) {
			r.store(cx.st_rev, p0);
		}
	} else if (CHAIN && w == 2) {
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
				r.run(a.slot[0].u.equalizer, a.slot[2].u.echo, sg, it % kBuffers, min(T, a.frames - it * T), sink(it));
			}
			bar_all();
		}
		if (active) {
			r.store(cx.st_eq, cx.st_echo);
		}
	}
}

// ---- the same pipeline with bulk-asynchronous ring traffic (whole tiles) ---------------------------------------------
// A span's T = 16 positions of one ring line are ONE contiguous run of 16 x 128 bytes (the rings are line-major), so
// the ring traffic of a span is a handful of bulk copies (cp.async.bulk, the 1-D TMA path) instead of ~60 address
// computations + 4-byte requests per frame: 18 + 9 row loads global -> shared (mbarrier complete_tx) and 20 + 5 row
// stores shared -> global (bulk groups) per span, issued by the lanes of one warp; the phases then read and write
// shared-memory rows at immediate offsets.  Loads are issued ahead and stores behind the arithmetic (Timing kBulkTiming,
// checked per coefficient block by plan_bulk_frames).  Only the chorus / flanger -- LFO-modulated positions, two reads and
// two writes per frame -- stays on ordinary loads and stores.
//
// Shared memory (rows of 16 x 32 floats = 2 KB): rowsC[3][20] | tapA[9] | outB[5] | staged words [3][12] = 220 KB for
// the chain.  A phase C buffer makes one round trip per span: the copy warp fills its 18 tap rows (iteration s), phase C
// reads them and writes its 20 store rows IN PLACE -- every element is read and written by the same thread -- (iteration
// s + 1), the copy warp sends the rows to the rings (iteration s + 2); three buffers rotate, so nobody ever waits for a
// buffer.  Per iteration `it` (B of span it):
//   parallel warps:  A(it+1) from tapA; barrier P (parallel warps only); [warp 0: request tapA(it+2)]; C(it-1)
//   serial warps:    B(it); the shelves / echo warps store their own main-line / echo rows; the T60 warp first sends
//                    the rows of span it-2 to the rings and requests the tap rows of span it
//   everybody:       CTA barrier
constexpr int kBulkParallelWarps = 8;
OALSFX_CX int bulk_threads(bool chain) { return (kBulkParallelWarps + serial_warps(chain)) * kLanes; }
constexpr int kRowsC = 20;                // rows of one phase C buffer: 18 tap rows in, 20 store rows out, in place
OALSFX_CX int bulk_shared_floats(bool chain)
{
	return (3 * kRowsC + tap_rows_a(chain) + out_rows_b(chain) + kBuffers * staged_words(chain)) * kBulkFrames * kLanes;
}

__device__ __forceinline__ unsigned shared_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(shared_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(shared_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" ::"r"(shared_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(float* smem_dst, const float* gmem_src, unsigned bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		::"r"(shared_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(shared_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(float* gmem_dst, const float* smem_src, unsigned bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(shared_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// `count` rows of a ring line starting at position pos0 (modulo mask + 1) <-> a shared-memory row buffer
__device__ __forceinline__ void move_rows(bool load, float* line, int mask, int pos0, int count, float* smem_row, uint64_t* bar)
{
	const int start = pos0 & mask;
	const int n1 = min(count, mask + 1 - start);
	if (load) {
		bulk_load(smem_row, line + static_cast<size_t>(start) * kLanes, static_cast<unsigned>(n1) * kLanes * 4U, bar);
		if (n1 < count) {
			bulk_load(smem_row + n1 * kLanes, line, static_cast<unsigned>(count - n1) * kLanes * 4U, bar);
		}
	} else {
		bulk_store(line + static_cast<size_t>(start) * kLanes, smem_row, static_cast<unsigned>(n1) * kLanes * 4U);
		if (n1 < count) {
			bulk_store(line, smem_row + n1 * kLanes, static_cast<unsigned>(count - n1) * kLanes * 4U);
		}
	}
}

template <int CT, bool CHAIN>
__global__ void __launch_bounds__(bulk_threads(CHAIN), 1) span_bulk_kernel(const __grid_constant__ MixArgs a)
{
	using Cx = Context<CT, CHAIN>;
	constexpr int NS = serial_warps(CHAIN), NP = kBulkParallelWarps, T = kBulkFrames, W = staged_words(CHAIN);
	constexpr int RC = tap_rows_c(CHAIN), RA = tap_rows_a(CHAIN), RB = out_rows_b(CHAIN), ROW = T * kLanes;
	constexpr int kCopyWarp = 1;                      // the T60 warp also moves phase C's rows
	constexpr unsigned kBarP = 1, kBarPThreads = NP * kLanes;
	extern __shared__ __align__(128) float dyn[];
	float* const rows_c = dyn;                        // [3][kRowsC][T][lane]: taps in, store rows out
	float* const tap_a = rows_c + 3 * kRowsC * ROW;   // [RA][T][lane]
	float* const out_b = tap_a + RA * ROW;            // [RB][T][lane]
	float* const staged = out_b + RB * ROW;           // [3][W][T][lane]
	__shared__ __align__(8) uint64_t mbar[4];         // phase C buffers 0..2, tapA

	const int tile = a.tiles ? static_cast<int>(a.tiles[blockIdx.x].tile) : a.tile_first + static_cast<int>(blockIdx.x);
	const int lane = threadIdx.x % kLanes;
	const int w = threadIdx.x / kLanes;

	Cx cx;
	bool ok = cx.setup(a, tile, lane);
	// the copies move whole rows: every stream of the tile must sit at the same ring positions
	ok = ok && cx.rev_off == __shfl_sync(0xFFFFFFFFU, cx.rev_off, 0);
	if (CHAIN) {
		ok = ok && cx.echo_off == __shfl_sync(0xFFFFFFFFU, cx.echo_off, 0);
	}
	ok = ok || !cx.io_ok;  // lanes past the last stream: their rings exist, their results are never looked at
	if (threadIdx.x == 0) {
		mbar_init(&mbar[0], 1);
		mbar_init(&mbar[1], 1);
		mbar_init(&mbar[2], 1);
		mbar_init(&mbar[3], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (!__syncthreads_and(ok)) {
		if (w == 0 && cx.io_ok) {
			exact_stream<CT, CHAIN>(a, tile, lane);
		}
		return;
	}
	const Stage<kLanes, W, T> sg = {staged + lane};
	const int nspans = (a.frames + T - 1) / T;
	const ReverbCoef& c = a.slot[Cx::RP].u.reverb;
	float* const tile_rev = cx.ring_rev.p - lane;
	float* const tile_echo = CHAIN ? cx.ring_echo.p - lane : nullptr;
	auto count_of = [&](int s) { return min(T, a.frames - s * T); };
	auto line_of = [&](const RowJob& j) { return (j.ring == 0 ? tile_rev : tile_echo) + static_cast<size_t>(j.word0) * kLanes; };
	auto pos_of = [&](const RowJob& j, int s) { return (j.ring == 0 ? cx.rev_off : cx.echo_off) + s * T - j.behind; };

	if (w >= NS) {
		// ---- parallel warps ----
		const int pw = w - NS;
		constexpr int U = T / NP;                    // frames of a span per warp: t = pw + u * NP
		auto load_tap_a = [&](int s) {               // lanes 0 .. RA-1 move a row each
			if (s < nspans) {
				if (lane == 0) {
					mbar_expect_tx(&mbar[3], static_cast<unsigned>(RA * count_of(s) * kLanes * 4));
				}
				__syncwarp();
				if (lane < RA) {
					const RowJob j = tap_row_a<CHAIN>(a, lane);
					move_rows(true, line_of(j), j.mask, pos_of(j, s), count_of(s), tap_a + lane * ROW, &mbar[3]);
				}
			}
		};
		if (pw == 0) {
			load_tap_a(0);
		}
		// Everything a frame needs besides the ring rows -- its input and (chain) its two chorus taps, LFO-modulated
		// positions that stay on ordinary loads -- is requested one iteration ahead and parked in registers.
		float xa[U][CT], xc[U][CT];
		typename Cx::TapsMod md[U];
#pragma unroll
		for (int u = 0; u < U; ++u) {
			cx.load_input(a, min(pw + u * NP, a.frames - 1), xa[u]);   // phase A of span 0
#pragma unroll
			for (int ch = 0; ch < CT; ++ch) {
				xc[u][ch] = 0.0F;
			}
			md[u].v[0] = md[u].v[1] = 0.0F;
			md[u].pos = 0;
		}
		for (int it = -1; it <= nspans; ++it) {
			float xa_next[U][CT], xc_next[U][CT];
			typename Cx::TapsMod md_next[U];
#pragma unroll
			for (int u = 0; u < U; ++u) {
				const int na = min((it + 2) * T + pw + u * NP, a.frames - 1);   // phase A of span it + 2
				const int nc = min(max(it, 0) * T + pw + u * NP, a.frames - 1); // phase C of span it
				cx.load_input(a, na, xa_next[u]);
				cx.load_input(a, nc, xc_next[u]);
				if (CHAIN) {
					cx.load_mod(a, nc, md_next[u]);
				}
			}
			if (it + 1 < nspans) {
				const int s = it + 1, buf = s % kBuffers, count = count_of(s);
				mbar_wait(&mbar[3], static_cast<unsigned>(s & 1));
				typename Cx::TapsA k[U];
#pragma unroll
				for (int u = 0; u < U; ++u) {
					cx.template fetch_a<T, kLanes>(tap_a + (pw + u * NP) * kLanes + lane, k[u]);
				}
#pragma unroll
				for (int u = 0; u < U; ++u) {
					const int t = pw + u * NP;
					if (t < count) {
						cx.phase_a(a, sg, buf, t, xa[u], k[u]);
					}
				}
			}
			asm volatile("bar.sync %0, %1;" ::"n"(kBarP), "n"(kBarPThreads) : "memory");
			if (pw == 0) {
				load_tap_a(it + 2);                  // every warp is through with the buffer
			}
			if (it >= 1) {
				const int s = it - 1, buf = s % kBuffers, first = s * T, count = count_of(s);
				mbar_wait(&mbar[s % 3], static_cast<unsigned>((s / 3) & 1));
				float* const rows = rows_c + (s % 3) * kRowsC * ROW + lane;
				typename Cx::TapsC k[U];
#pragma unroll
				for (int u = 0; u < U; ++u) {
					cx.template fetch_c<T, kLanes>(rows + (pw + u * NP) * kLanes, k[u]);
					cx.fetch_staged(sg, buf, min(pw + u * NP, count - 1), k[u]);
				}
#pragma unroll
				for (int u = 0; u < U; ++u) {
					const int t = pw + u * NP;
					if (t < count) {
						cx.phase_c(a, first + t, xc[u], k[u], md[u], InPlaceSinkC<T>{rows + t * kLanes});
					}
				}
			}
#pragma unroll
			for (int u = 0; u < U; ++u) {
#pragma unroll
				for (int ch = 0; ch < CT; ++ch) {
					xa[u][ch] = xa_next[u][ch];
					xc[u][ch] = xc_next[u][ch];
				}
				md[u] = md_next[u];
			}
			fence_async_shared();                    // the rows written above are copied out by the copy warp after the barrier
			bar_all();
		}
		if (pw == 0) {
			cx.store_scalars(a);
		} else if (pw == 1) {
			cx.store_send_history(a, 0);
		} else if (pw - 2 < (CHAIN ? 4 : 1)) {
			cx.store_send_history(a, 1 + a.aux_index[pw - 2]);
		}
		return;
	}

	// ---- serial warps ----
	const StageSinkB<T> sink = {out_b + lane};
	// Phase B's own rows are stored by the warp that wrote them: before it writes them again it waits until the copies
	// have read shared memory, at the end of every iteration until the copies of the iteration before have landed.
	auto store_rows_b = [&](int s, int row0, int n) {
		fence_async_shared();
		__syncwarp();
		if (lane < n) {
			const RowJob j = out_row_b<CHAIN>(a, row0 + lane);
			move_rows(false, line_of(j), j.mask, pos_of(j, s), count_of(s), out_b + (row0 + lane) * ROW, nullptr);
		}
	};
	auto stores_landed_but_newest = [&]() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); };
	// phase C's tap rows of span s -> buffer s % 3 (free: its store copies had read it before the last barrier)
	auto request_taps_c = [&](int s) {
		if (s >= 0 && s < nspans) {
			if (lane == 0) {
				mbar_expect_tx(&mbar[s % 3], static_cast<unsigned>(RC * count_of(s) * kLanes * 4));
			}
			__syncwarp();
			if (lane < RC) {
				const RowJob j = tap_row_c<CHAIN>(a, lane);
				move_rows(true, line_of(j), j.mask, pos_of(j, s), count_of(s), rows_c + ((s % 3) * kRowsC + lane) * ROW, &mbar[s % 3]);
			}
		}
	};
	constexpr int kRequestWarp = CHAIN ? 2 : 0;       // the equalizer pair warp (chain) / the shelves warp requests phase C's taps
	if (w == kCopyWarp) {
		T60s<2> r;
		r.load(cx.st_rev, 0);
		for (int it = -1; it <= nspans + 1; ++it) {
			// phase C of span it - 2 ran in the previous iteration: its 20 store rows -> rings
			if (it - 2 >= 0 && it - 2 < nspans && lane < kRowsC) {
				const RowJob j = out_row_c<CHAIN>(a, out_of_buffer_row(lane));
				move_rows(false, line_of(j), j.mask, pos_of(j, it - 2), count_of(it - 2), rows_c + (((it - 2) % 3) * kRowsC + lane) * ROW, nullptr);
			}
			bulk_commit();
			if (it > nspans) {
				break;
			}
			if (it >= 0 && it < nspans) {
				r.run(c, sg, it % kBuffers, count_of(it), 0);
			}
			bulk_wait_read();                        // the buffer just sent off may be filled again after the barrier
			stores_landed_but_newest();              // ... and the rows sent off one iteration ago have landed
			bar_all();
		}
		bulk_wait_all();
		r.store(cx.st_rev, 0);
	} else if (w == 0) {
		Shelves<2> r;
		r.load(cx.st_rev, 0);
		for (int it = -1; it <= nspans; ++it) {
			if (kRequestWarp == 0) {
				request_taps_c(it);
			}
			if (it >= 0 && it < nspans) {
				bulk_wait_read();
				r.run(c, sg, it % kBuffers, count_of(it), sink, 0);
				store_rows_b(it, kOutMain, 4);
			}
			bulk_commit();
			stores_landed_but_newest();
			bar_all();
		}
		bulk_wait_all();
		r.store(cx.st_rev, 0);
	} else if (CHAIN && w == 2) {
		EqPair r;
		r.load(cx.st_eq);
		for (int it = -1; it <= nspans; ++it) {
			request_taps_c(it);
			if (it >= 0 && it < nspans) {
				r.run(a.slot[0].u.equalizer, sg, it % kBuffers, count_of(it));
			}
			bar_all();
		}
		r.store(cx.st_eq);
	} else if (CHAIN) {
		EqSingleEcho r;
		r.load(cx.st_eq, cx.st_echo);
		for (int it = -1; it <= nspans; ++it) {
			if (it >= 0 && it < nspans) {
				bulk_wait_read();
				r.run(a.slot[0].u.equalizer, a.slot[2].u.echo, sg, it % kBuffers, count_of(it), sink);
				store_rows_b(it, kOutEcho, 1);
			}
			bulk_commit();
			stores_landed_but_newest();
			bar_all();
		}
		bulk_wait_all();
		r.store(cx.st_eq, cx.st_echo);
	}
}

#endif // __CUDACC__

} // namespace span
} // namespace oalsfx

#endif
