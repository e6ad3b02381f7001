// quad.cuh -- the throughput kernels: FOUR lanes per stream.
//
// fx.cuh/mix.cuh run one thread per stream; that keeps ~220 words of recurrent state per thread
// (255 registers, 2 warps per scheduler) and serialises the ~24 ring reads of a sample behind its
// stores, so it is latency-bound (profiles/r01_v0_thread_per_stream.txt).  Here a stream is owned by
// a "quad" of 4 adjacent lanes:
//   * lane j runs reverb delay line j, equalizer / ring-modulator / compressor wet channel j,
//     chorus side j&1, and accumulates output channels j (and j+4);
//   * the 4x4 scatter matrices, the B->A format conversion and the ordered output sums exchange
//     values with quad-wide shuffles;
//   * the six ring reads a lane needs for sample n+1 are issued at the top of sample n (legal
//     whenever every delay is >= 2 samples and no cross-fade / modulation is active), so DRAM
//     latency overlaps a whole sample of arithmetic.
// Per-lane state is ~1/4 of the thread-per-stream kernel's, which quadruples the resident warps.
//
// Same state layout in HBM, same arithmetic, same per-sample operation order as fx.cuh (which the
// CPU test build checks bit for bit against the reference); only the lane that performs each
// operation differs.  A warp = 8 streams; a CTA of 128 threads = one 32-stream tile.
#ifndef OALSFX_QUAD_CUH
#define OALSFX_QUAD_CUH

#if defined(__CUDACC__)

#include "mix.cuh"

namespace oalsfx {
namespace quad {

constexpr unsigned kFull = 0xFFFFFFFFU;

#ifndef OALSFX_QUAD_MIN_CTAS
#define OALSFX_QUAD_MIN_CTAS 4   // resident CTAs per SM the register allocation is sized for (128 regs/thread)
#endif

// value of `v` held by lane `l` (0..3) of my quad
__device__ __forceinline__ float qget(float v, int l) { return __shfl_sync(kFull, v, l, 4); }

template <int CT> struct Own { static constexpr int n = (CT + 3) / 4; }; // output channels per lane

// acc[o] += v * g[o] for the channels this lane owns
template <int CT>
__device__ __forceinline__ void own_add(float* acc, const float* g, float v)
{
#pragma unroll
	for (int o = 0; o < Own<CT>::n; ++o) {
		if (audible(g[o])) {
			acc[o] += v * g[o];
		}
	}
}

// gains row -> the entries of the channels this lane owns (0 for channels >= CT)
template <int CT>
__device__ __forceinline__ void own_gains(float* dst, const float* row, int j)
{
#pragma unroll
	for (int o = 0; o < Own<CT>::n; ++o) {
		const int k = j + 4 * o;
		dst[o] = (k < CT ? row[k] : 0.0F);
	}
}

// wet channel `k` of a send for the current input frame (MixHelpers::mix with static gains)
template <int CT>
__device__ __forceinline__ float wet_channel(const float* gains_for_k, const float* x)
{
	float w = 0.0F;
#pragma unroll
	for (int c = 0; c < CT; ++c) {
		if (audible(gains_for_k[c])) {
			w += x[c] * gains_for_k[c];
		}
	}
	return w;
}

// Per-lane send gains for one wet channel: g[c] = send.gains[c][k]
template <int CT>
__device__ __forceinline__ void send_gains(float* dst, const SendCoef& s, int k)
{
#pragma unroll
	for (int c = 0; c < CT; ++c) {
		dst[c] = s.gains[c][k];
	}
}

struct Ctx {
	int j;            // lane within the quad
	uint32_t* state;  // slot state of my stream (lane offset applied)
	float* ring;      // slot ring region of my stream (lane offset applied)
	float* pf;        // this thread's column of the slot's prefetch buffer in shared memory
};

// Deep prefetch of ring reads through cp.async (LDGSTS): no registers are tied up while the loads
// are in flight, so the bytes in flight per SM (= achievable HBM bandwidth x latency) are set by the
// depth, not by the register file.
constexpr int kQuadThreads = 128;
constexpr int kPfSlots = 8;                 // power of two
constexpr int kPfDepth = kPfSlots - 1;      // positions in flight beyond the current one
constexpr int kPfTaps = 6;
constexpr int kPfFloatsPerSlotUser = kPfSlots * kPfTaps * kQuadThreads;

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src)
{
	const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- null / dedicated ------------------------------------------------------------------------------
template <int CT>
struct QNull {
	static constexpr bool kIsNull = true;
	__device__ void begin(const SlotCoef&, const SendCoef&, const Ctx&, bool, int) {}
	__device__ void step(const SlotCoef&, const float*, float*) {}
	__device__ void end(const Ctx&) {}
};

template <int CT>
struct QDedicated {
	static constexpr bool kIsNull = false;
	float g0[CT], go[Own<CT>::n];
	__device__ void begin(const SlotCoef& sc, const SendCoef& s, const Ctx& cx, bool, int)
	{
		send_gains<CT>(g0, s, 0);
		own_gains<CT>(go, sc.u.dedicated.gains, cx.j);
	}
	__device__ void step(const SlotCoef&, const float* x, float* acc) { own_add<CT>(acc, go, wet_channel<CT>(g0, x)); }
	__device__ void end(const Ctx&) {}
};

// ---- equalizer: lane j filters wet channel j -------------------------------------------------------------
template <int CT>
struct QEqualizer {
	static constexpr bool kIsNull = false;
	BiquadHist h[4];
	float gs[CT], go[4][Own<CT>::n];
	int j;
	__device__ void begin(const SlotCoef& sc, const SendCoef& s, const Ctx& cx, bool, int)
	{
		j = cx.j;
		send_gains<CT>(gs, s, j);
#pragma unroll
		for (int b = 0; b < 4; ++b) {
			load_words(h[b], cx.state + ((b * 4 + j) * 4) * kLanes); // FxEqualizer::State::h[b][j]
			own_gains<CT>(go[b], sc.u.equalizer.gains[b], j);           // reused below as gains[ft = b]
		}
	}
	__device__ void step(const SlotCoef& sc, const float* x, float* acc)
	{
		const EqualizerCoef& c = sc.u.equalizer;
		float v = wet_channel<CT>(gs, x);
#pragma unroll
		for (int b = 0; b < 4; ++b) {
			v = biquad_step(c.band[b], h[b], v);
		}
#pragma unroll
		for (int ft = 0; ft < 4; ++ft) {
			own_add<CT>(acc, go[ft], qget(v, ft));
		}
	}
	__device__ void end(const Ctx& cx)
	{
#pragma unroll
		for (int b = 0; b < 4; ++b) {
			store_words(h[b], cx.state + ((b * 4 + j) * 4) * kLanes);
		}
	}
};

// ---- chorus / flanger: lane j runs side j & 1 --------------------------------------------------------
template <int CT>
struct QModDelay {
	static constexpr bool kIsNull = false;
	int32_t offset, phase, side_base, j;
	float g0[CT], gl[Own<CT>::n], gr[Own<CT>::n];
	LaneMem ring;
	__device__ void begin(const SlotCoef& sc, const SendCoef& s, const Ctx& cx, bool, int)
	{
		const ModDelayCoef& c = sc.u.mod_delay;
		j = cx.j;
		offset = static_cast<int32_t>(cx.state[0]);
		ring.p = cx.ring;
		const int side = j & 1;
		side_base = side * (c.mask + 1);
		phase = (side == 0 ? offset % c.lfo_range : (offset + c.lfo_disp) % c.lfo_range);
		send_gains<CT>(g0, s, 0);
		own_gains<CT>(gl, c.gains[0], j);
		own_gains<CT>(gr, c.gains[1], j);
	}
	__device__ void step(const SlotCoef& sc, const float* x, float* acc)
	{
		const ModDelayCoef& c = sc.u.mod_delay;
		const float in = wet_channel<CT>(g0, x);
		int32_t d;
		if (c.waveform == 1) {
			d = static_cast<int32_t>((1.0F - fabsf(2.0F - (c.lfo_scale * phase))) * c.depth) + c.delay;
		} else {
			d = c.sin_delays[phase];
		}
		const int32_t pos = offset & c.mask;
		const int32_t rd = (offset - d) & c.mask;
		const float tapped = (rd == pos ? in : ring.ld(side_base + rd));
		const float t = tapped * c.feedback;
		if (j < 2) {
			ring.st(side_base + pos, in + t);
		}
		phase += 1;
		if (phase >= c.lfo_range) {
			phase = 0;
		}
		offset += 1;
		own_add<CT>(acc, gl, qget(t, 0));
		own_add<CT>(acc, gr, qget(t, 1));
	}
	__device__ void end(const Ctx& cx)
	{
		if (j == 0) {
			cx.state[0] = static_cast<uint32_t>(offset);
		}
	}
};

// ---- echo: computed redundantly by the four lanes (same addresses), lane 0 stores -----------------------
template <int CT>
struct QEcho {
	static constexpr bool kIsNull = false;
	FxEcho::State s;
	float g0[CT], g1[Own<CT>::n], g2[Own<CT>::n];
	LaneMem ring;
	int j;
	float n1, n2;     // taps of the NEXT sample (prefetched)
	bool have_next;
	__device__ void begin(const SlotCoef& sc, const SendCoef& snd, const Ctx& cx, bool, int)
	{
		j = cx.j;
		load_words(s, cx.state);
		ring.p = cx.ring;
		send_gains<CT>(g0, snd, 0);
		own_gains<CT>(g1, sc.u.echo.gains[0], j);
		own_gains<CT>(g2, sc.u.echo.gains[1], j);
		have_next = false;
		n1 = n2 = 0.0F;
	}
	__device__ void step(const SlotCoef& sc, const float* x, float* acc)
	{
		const EchoCoef& c = sc.u.echo;
		float t1, t2;
		if (c.tap1 >= 2) { // the next sample's taps do not alias this sample's write (tap2 >= tap1)
			if (!have_next) {
				n1 = ring.ld((s.offset - c.tap1) & c.mask);
				n2 = ring.ld((s.offset - c.tap2) & c.mask);
				have_next = true;
			}
			t1 = n1;
			t2 = n2;
			n1 = ring.ld((s.offset + 1 - c.tap1) & c.mask);
			n2 = ring.ld((s.offset + 1 - c.tap2) & c.mask);
		} else {
			t1 = ring.ld((s.offset - c.tap1) & c.mask);
			t2 = ring.ld((s.offset - c.tap2) & c.mask);
		}
		const float in = t2 + wet_channel<CT>(g0, x);
		const float out = biquad_step(c.filter, s.f, in);
		if (j == 0) {
			ring.st(s.offset & c.mask, out * c.feed_gain);
		}
		s.offset += 1;
		own_add<CT>(acc, g1, t1);
		own_add<CT>(acc, g2, t2);
	}
	__device__ void end(const Ctx& cx)
	{
		if (j == 0) {
			store_words(s, cx.state);
		}
	}
};

// ---- distortion: redundant in the four lanes ---------------------------------------------------------------
template <int CT>
struct QDistortion {
	static constexpr bool kIsNull = false;
	FxDistortion::State s;
	float g0[CT], go[Own<CT>::n];
	int j;
	__device__ void begin(const SlotCoef& sc, const SendCoef& snd, const Ctx& cx, bool, int)
	{
		j = cx.j;
		load_words(s, cx.state);
		send_gains<CT>(g0, snd, 0);
		own_gains<CT>(go, sc.u.distortion.gains, j);
	}
	__device__ void step(const SlotCoef& sc, const float* x, float* acc)
	{
		const DistortionCoef& c = sc.u.distortion;
		const float fc = c.edge_coeff;
		const float w = wet_channel<CT>(g0, x);
		float kept = 0.0F;
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const float in = (k == 0 ? w * 4.0F : 0.0F);
			float smp = biquad_step(c.low_pass, s.lp, in);
			smp = (1.0F + fc) * smp / (1.0F + (fc * fabsf(smp)));
			smp = (1.0F + fc) * smp / (1.0F + (fc * fabsf(smp))) * -1.0F;
			smp = (1.0F + fc) * smp / (1.0F + (fc * fabsf(smp)));
			const float out = biquad_step(c.band_pass, s.bp, smp);
			if (k == 0) {
				kept = out;
			}
		}
		own_add<CT>(acc, go, kept);
	}
	__device__ void end(const Ctx& cx)
	{
		if (j == 0) {
			store_words(s, cx.state);
		}
	}
};

// ---- ring modulator: lane j filters wet channel j ------------------------------------------------------------
template <int CT>
struct QRingMod {
	static constexpr bool kIsNull = false;
	BiquadHist h;
	int32_t index, j;
	float gs[CT], go[4][Own<CT>::n];
	__device__ void begin(const SlotCoef& sc, const SendCoef& snd, const Ctx& cx, bool, int)
	{
		j = cx.j;
		load_words(h, cx.state + (j * 4) * kLanes);
		index = static_cast<int32_t>(cx.state[16 * kLanes]);
		send_gains<CT>(gs, snd, j);
#pragma unroll
		for (int w = 0; w < 4; ++w) {
			own_gains<CT>(go[w], sc.u.ring_mod.gains[w], j);
		}
	}
	__device__ void step(const SlotCoef& sc, const float* x, float* acc)
	{
		const RingModCoef& c = sc.u.ring_mod;
		constexpr int32_t frac_one = 1 << 24;
		index = (index + c.step) & (frac_one - 1);
		float m;
		if (c.waveform == 0) {
			m = sinf(index * (6.28318530717958647692F / frac_one) - 3.14159265358979323846F) * 0.5F + 0.5F;
		} else if (c.waveform == 1) {
			m = static_cast<float>(index) / frac_one;
		} else {
			m = static_cast<float>((index >> 23) & 1);
		}
		const float y = biquad_step(c.filter, h, wet_channel<CT>(gs, x)) * m;
#pragma unroll
		for (int w = 0; w < 4; ++w) {
			own_add<CT>(acc, go[w], qget(y, w));
		}
	}
	__device__ void end(const Ctx& cx)
	{
		store_words(h, cx.state + (j * 4) * kLanes);
		if (j == 0) {
			cx.state[16 * kLanes] = static_cast<uint32_t>(index);
		}
	}
};

// ---- compressor: lane j scales wet channel j, the envelope is computed by all four ----------------------------
template <int CT>
struct QCompressor {
	static constexpr bool kIsNull = false;
	float gain_control;
	float gs[CT], go[4][Own<CT>::n];
	int j;
	__device__ void begin(const SlotCoef& sc, const SendCoef& snd, const Ctx& cx, bool, int)
	{
		j = cx.j;
		FxCompressor::State st;
		load_words(st, cx.state);
		gain_control = st.initialized ? st.gain_control : 1.0F;
		send_gains<CT>(gs, snd, j);
#pragma unroll
		for (int w = 0; w < 4; ++w) {
			own_gains<CT>(go[w], sc.u.compressor.gains[w], j);
		}
	}
	__device__ void step(const SlotCoef& sc, const float* x, float* acc)
	{
		const CompressorCoef& c = sc.u.compressor;
		const float w = wet_channel<CT>(gs, x);
		const float a = fabsf(w);
		const float a0 = qget(a, 0), a1 = qget(a, 1), a2 = qget(a, 2), a3 = qget(a, 3);
		float amplitude = 1.0F;
		if (c.enabled) {
			amplitude = fmaxf(a0 + a1, fmaxf(a0 + a2, a0 + a3));
		}
		if (amplitude > gain_control) {
			gain_control = fminf(gain_control + c.attack_rate, amplitude);
		} else if (amplitude < gain_control) {
			gain_control = fmaxf(gain_control - c.release_rate, amplitude);
		}
		const float y = w * (1.0F / fminf(2.0F, fmaxf(0.5F, gain_control)));
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			own_add<CT>(acc, go[q], qget(y, q));
		}
	}
	__device__ void end(const Ctx& cx)
	{
		if (j == 0) {
			FxCompressor::State st;
			st.gain_control = gain_control;
			st.initialized = 1;
			store_words(st, cx.state);
		}
	}
};

// ---- reverb / EAX reverb: lane j runs delay line j -----------------------------------------------------------------
// Word offsets of FxReverb::State members (kept identical so both kernel families share the state).
constexpr int kRvLp = 0, kRvHp = 16, kRvT60 = 32, kRvGain = 48, kRvOldEt = 48 + 64, kRvOldEap = kRvOldEt + 4,
	kRvOldEo = kRvOldEt + 8, kRvOldLt = kRvOldEt + 12, kRvOldLap = kRvOldEt + 16, kRvOldLo = kRvOldEt + 20,
	kRvScalars = kRvOldEt + 24;
static_assert(kRvScalars + 5 == FxReverb::kStateWords, "reverb state layout drifted");

template <int CT>
struct QReverb {
	static constexpr bool kIsNull = false;
	static constexpr int NO = Own<CT>::n;
	// recurrent state of line j
	BiquadHist lp, hp;
	float t60s[2][2];
	float cur[NO][8], stp[NO][8];
	uint32_t ramp, act;                 // bit o*8 + l
	int32_t old_et, old_eap, old_eo, old_lt, old_lap, old_lo;
	int32_t offset, fade_count, mod_index, mod_range;
	float mod_filter;
	// per-lane constants
	int32_t et, eap, eo, lt, lap, lo;   // new taps of line j
	int32_t b_main, b_eap, b_el, b_lap, b_ll; // first ring word of line j in each ring
	float etc, ecf, lf0, lf1, lf2, hf0, hf1, hf2, mid;
	float sg0, sg1, sg2, sg3;           // B->A row j
	int32_t o1, o2, o3, rs;             // scatter row j: the other lines in ascending order; rs = 3 - j
	float s1, s2, s3;                   // their signs
	float gs[CT];
	int32_t j;
	LaneMem ring;
	// block bookkeeping
	int32_t block_frames, base, sub_left, sub_todo;
	float fade;
	bool faded, can_pf, primed;
	int32_t pos_issue;                  // next ring position whose six reads have not been issued yet
	float* pf_slot;                     // my column of the CTA's prefetch buffer: [slot][tap][thread]

	__device__ void begin(const SlotCoef& sc, const SendCoef& snd, const Ctx& cx, bool update, int frames)
	{
		const ReverbCoef& c = sc.u.reverb;
		j = cx.j;
		ring.p = cx.ring;
		const uint32_t* st = cx.state;
		load_words(lp, st + (kRvLp + j * 4) * kLanes);
		load_words(hp, st + (kRvHp + j * 4) * kLanes);
		t60s[0][0] = __uint_as_float(st[(kRvT60 + j * 4 + 0) * kLanes]);
		t60s[0][1] = __uint_as_float(st[(kRvT60 + j * 4 + 1) * kLanes]);
		t60s[1][0] = __uint_as_float(st[(kRvT60 + j * 4 + 2) * kLanes]);
		t60s[1][1] = __uint_as_float(st[(kRvT60 + j * 4 + 3) * kLanes]);
#pragma unroll
		for (int o = 0; o < NO; ++o) {
			const int k = j + 4 * o;
#pragma unroll
			for (int l = 0; l < 8; ++l) {
				cur[o][l] = (k < CT ? __uint_as_float(st[(kRvGain + l * kMaxChannels + k) * kLanes]) : 0.0F);
				stp[o][l] = 0.0F;
			}
		}
		old_et = static_cast<int32_t>(st[(kRvOldEt + j) * kLanes]);
		old_eap = static_cast<int32_t>(st[(kRvOldEap + j) * kLanes]);
		old_eo = static_cast<int32_t>(st[(kRvOldEo + j) * kLanes]);
		old_lt = static_cast<int32_t>(st[(kRvOldLt + j) * kLanes]);
		old_lap = static_cast<int32_t>(st[(kRvOldLap + j) * kLanes]);
		old_lo = static_cast<int32_t>(st[(kRvOldLo + j) * kLanes]);
		offset = static_cast<int32_t>(st[(kRvScalars + 0) * kLanes]);
		fade_count = static_cast<int32_t>(st[(kRvScalars + 1) * kLanes]);
		mod_index = static_cast<int32_t>(st[(kRvScalars + 2) * kLanes]);
		mod_range = static_cast<int32_t>(st[(kRvScalars + 3) * kLanes]);
		mod_filter = __uint_as_float(st[(kRvScalars + 4) * kLanes]);
		if (mod_range == 0) {
			mod_range = 1;
		}

		et = c.early_tap[j]; eap = c.early_ap_off[j]; eo = c.early_off[j];
		lt = c.late_tap[j]; lap = c.late_ap_off[j]; lo = c.late_off[j];
		etc = c.early_tap_coeff[j]; ecf = c.early_coeff[j];
		lf0 = c.t60_lf[j][0]; lf1 = c.t60_lf[j][1]; lf2 = c.t60_lf[j][2];
		hf0 = c.t60_hf[j][0]; hf1 = c.t60_hf[j][1]; hf2 = c.t60_hf[j][2];
		mid = c.t60_mid[j];
		b_main = c.ring_base[0] + j * (c.mask[0] + 1);
		b_eap = c.ring_base[1] + j * (c.mask[1] + 1);
		b_el = c.ring_base[2] + j * (c.mask[2] + 1);
		b_lap = c.ring_base[3] + j * (c.mask[3] + 1);
		b_ll = c.ring_base[4] + j * (c.mask[4] + 1);
		// B-format -> A-format row j (oalsfxpp.cpp:6377-6383)
		constexpr float q = 0.288675134595F;
		sg0 = q;
		sg1 = (j == 1 || j == 3) ? -q : q;
		sg2 = (j == 1 || j == 2) ? -q : q;
		sg3 = (j == 2 || j == 3) ? -q : q;
		// scatter row j (oalsfxpp.cpp:7510-7521)
		o1 = (j == 0 ? 1 : 0);
		o2 = (j <= 1 ? 2 : 1);
		o3 = (j == 3 ? 2 : 3);
		s1 = (j == 0 || j == 2) ? 1.0F : -1.0F;
		s2 = (j == 1) ? 1.0F : -1.0F;
		s3 = (j == 3) ? -1.0F : 1.0F;
		rs = 3 - j;
		send_gains<CT>(gs, snd, j);

		if (update) {
			mod_index = static_cast<int32_t>(mod_index * static_cast<int64_t>(c.mod_range) / mod_range);
			mod_range = c.mod_range;
			const bool mine = et != old_et || eap != old_eap || eo != old_eo || lt != old_lt || lap != old_lap || lo != old_lo;
			// "any of the four lines differs" -- a quad-wide OR
			const unsigned ballot = __ballot_sync(kFull, mine);
			const unsigned quad_bits = (ballot >> ((threadIdx.x & 31) & ~3)) & 0xFU;
			if (quad_bits != 0) {
				fade_count = 0;
			}
		}
		block_frames = frames;
		base = 0;
		sub_left = 0;
		sub_todo = 0;
		fade = static_cast<float>(fade_count) / FxReverb::kFadeSamples;
		faded = false;
		primed = false;
		pos_issue = 0;
		can_pf = false;
		ramp = act = 0;
		pf_slot = cx.pf;
	}

	__device__ void begin_sub(const ReverbCoef& c)
	{
		int todo = block_frames - base;
		if (todo > FxReverb::kMaxUpdate) {
			todo = FxReverb::kMaxUpdate;
		}
		if (FxReverb::kFadeSamples - fade_count > 0 && todo > FxReverb::kFadeSamples - fade_count) {
			todo = FxReverb::kFadeSamples - fade_count;
		}
		sub_todo = todo;
		sub_left = todo;
		faded = fade < 1.0F;
		// Reading kPfDepth samples ahead is legal when nothing written during those samples can be what
		// the prefetch reads: every delay > kPfDepth, late tap that far beyond the late feed write, no
		// cross-fade (reads both tap sets), no modulation (the late line read position moves).
		can_pf = !faded && c.mod_depth == 0.0F && mod_filter == 0.0F && et > kPfDepth && eap > kPfDepth && eo > kPfDepth &&
			lt > c.late_feed_tap + kPfDepth && lap > kPfDepth && lo > kPfDepth;
		const float delta = 1.0F / static_cast<float>(block_frames - base);
		ramp = act = 0;
#pragma unroll
		for (int o = 0; o < NO; ++o) {
			const int k = j + 4 * o;
#pragma unroll
			for (int l = 0; l < 8; ++l) {
				const float target = (k < CT ? (l < 4 ? c.pan_early[l][k] : c.pan_late[l - 4][k]) : 0.0F);
				const float step = (target - cur[o][l]) * delta;
				if (k < CT && fabsf(step) > FLT_EPSILON) {
					ramp |= 1U << (o * 8 + l);
					stp[o][l] = step;
				} else {
					stp[o][l] = 0.0F;
					if (k < CT && audible(cur[o][l])) {
						act |= 1U << (o * 8 + l);
					}
				}
			}
		}
	}

	__device__ void end_sub(const ReverbCoef& c)
	{
		if (faded) {
			fade = fminf(1.0F, fade);
		}
		if (fade_count < FxReverb::kFadeSamples) {
			fade_count += sub_todo;
			if (fade_count >= FxReverb::kFadeSamples) {
				fade_count = FxReverb::kFadeSamples;
				fade = 1.0F;
				old_et = et; old_eap = eap; old_eo = eo; old_lt = lt; old_lap = lap; old_lo = lo;
			}
		}
		if (sub_todo == block_frames - base) {
#pragma unroll
			for (int o = 0; o < NO; ++o) {
				const int k = j + 4 * o;
#pragma unroll
				for (int l = 0; l < 8; ++l) {
					if ((ramp >> (o * 8 + l)) & 1U) {
						cur[o][l] = (l < 4 ? c.pan_early[l][k < CT ? k : 0] : c.pan_late[l - 4][k < CT ? k : 0]);
					}
				}
			}
		}
		base += sub_todo;
	}

	__device__ __forceinline__ float rd(int base_word, int mask, int pos, int old_d, int new_d, float mu) const
	{
		if (!faded) {
			return ring.ld(base_word + ((pos - new_d) & mask));
		}
		const float a = ring.ld(base_word + ((pos - old_d) & mask));
		const float b = ring.ld(base_word + ((pos - new_d) & mask));
		return a + ((b - a) * mu);
	}

	// row j of vector_partial_scatter applied to the vector whose element i sits in lane src(i)
	__device__ __forceinline__ float scatter_row(float self, float a, float b, float c3, float x, float y) const
	{
		return (x * self) + (y * (((a * s1) + (b * s2)) + (c3 * s3)));
	}

	__device__ void step(const SlotCoef& sc, const float* x, float* acc)
	{
		const ReverbCoef& c = sc.u.reverb;
		if (sub_left == 0) {
			begin_sub(c);
		}
		const int pos = offset;
		const float mu = fade;
		const int m_main = c.mask[0], m_eap = c.mask[1], m_el = c.mask[2], m_lap = c.mask[3], m_ll = c.mask[4];

		// The whole warp takes one path so that the shuffles below stay convergent.
		const bool pf = __all_sync(kFull, can_pf);
		float v_et = 0.0F, v_eap = 0.0F, v_eo = 0.0F, v_lt = 0.0F, v_lo = 0.0F, v_lap = 0.0F;
		if (pf) {
			// cp.async pipeline: the reads of positions pos .. pos + kPfDepth are in flight or landed;
			// one commit group per position, kPfSlots = kPfDepth + 1 shared-memory slots.
			if (!primed) {
				pos_issue = pos;
				primed = true;
			}
			while (pos_issue - pos <= kPfDepth) {
				float* slot = pf_slot + (pos_issue & (kPfSlots - 1)) * (kPfTaps * kQuadThreads);
				cp_async4(slot + 0 * kQuadThreads, ring.p + static_cast<unsigned>(b_main + ((pos_issue - et) & m_main)) * kLanes);
				cp_async4(slot + 1 * kQuadThreads, ring.p + static_cast<unsigned>(b_eap + ((pos_issue - eap) & m_eap)) * kLanes);
				cp_async4(slot + 2 * kQuadThreads, ring.p + static_cast<unsigned>(b_el + ((pos_issue - eo) & m_el)) * kLanes);
				cp_async4(slot + 3 * kQuadThreads, ring.p + static_cast<unsigned>(b_main + ((pos_issue - lt) & m_main)) * kLanes);
				cp_async4(slot + 4 * kQuadThreads, ring.p + static_cast<unsigned>(b_ll + ((pos_issue - lo) & m_ll)) * kLanes);
				cp_async4(slot + 5 * kQuadThreads, ring.p + static_cast<unsigned>(b_lap + ((pos_issue - lap) & m_lap)) * kLanes);
				cp_async_commit();
				pos_issue += 1;
			}
			cp_async_wait<kPfDepth>();
			const float* slot = pf_slot + (pos & (kPfSlots - 1)) * (kPfTaps * kQuadThreads);
			v_et = slot[0 * kQuadThreads];
			v_eap = slot[1 * kQuadThreads];
			v_eo = slot[2 * kQuadThreads];
			v_lt = slot[3 * kQuadThreads];
			v_lo = slot[4 * kQuadThreads];
			v_lap = slot[5 * kQuadThreads];
		} else {
			primed = false;
		}

		// input: wet_j -> A-format line j -> shelf filter(s) -> main line
		const float w = wet_channel<CT>(gs, x);
		float a = 0.0F;
		a += qget(w, 0) * sg0;
		a += qget(w, 1) * sg1;
		a += qget(w, 2) * sg2;
		a += qget(w, 3) * sg3;
		float v = biquad_step(c.lp, lp, a);
		if (c.is_eax) {
			v = biquad_step(c.hp, hp, v);
		}
		ring.st(b_main + (pos & m_main), v);

		// ---- early reflections ----
		float f = (pf ? v_et : rd(b_main, m_main, pos, old_et, et, mu)) * etc;
		{
			const float in = f;
			const float z = (pf ? v_eap : rd(b_eap, m_eap, pos, old_eap, eap, mu));
			f = z - (c.ap_feed_coeff * in);
			const float g = in + (c.ap_feed_coeff * f);
			const float ga = qget(g, o1), gb = qget(g, o2), gc = qget(g, o3);
			ring.st(b_eap + (pos & m_eap), scatter_row(g, ga, gb, gc, c.mix_x, c.mix_y));
		}
		ring.st(b_el + (pos & m_el), qget(f, rs));
		f += (pf ? v_eo : rd(b_el, m_el, pos, old_eo, eo, mu)) * ecf;
		const float early_out = f;
		{
			// reversed vector r_i = f_(3-i): element i lives in lane 3 - i
			const float self = qget(f, rs), ra = qget(f, 3 - o1), rb = qget(f, 3 - o2), rc = qget(f, 3 - o3);
			ring.st(b_main + ((pos - c.late_feed_tap) & m_main), scatter_row(self, ra, rb, rc, c.mix_x, c.mix_y));
		}

		// ---- late reverb ----
		int mod_delay = 0;
		{
			const bool quiet = (c.mod_depth == 0.0F && mod_filter == 0.0F);
			const float sinus = (quiet ? 0.0F : c.mod_sinus[mod_index]);
			mod_index += 1;
			if (mod_index >= mod_range) {
				mod_index = 0;
			}
			if (!quiet) {
				mod_filter = mod_filter + ((c.mod_depth - mod_filter) * c.mod_coeff);
				mod_delay = static_cast<int>(lroundf(mod_filter * sinus));
			}
		}
		f = (pf ? v_lt : rd(b_main, m_main, pos, old_lt, lt, mu)) * c.density_gain;
		f += (pf ? v_lo : rd(b_ll, m_ll, pos - mod_delay, old_lo, lo, mu));
		{
			const float o1v = (lf0 * f) + (lf1 * t60s[0][0]) + (lf2 * t60s[0][1]);
			t60s[0][0] = f;
			t60s[0][1] = o1v;
			const float o2v = (hf0 * o1v) + (hf1 * t60s[1][0]) + (hf2 * t60s[1][1]);
			t60s[1][0] = o1v;
			t60s[1][1] = o2v;
			f = mid * o2v;
		}
		{
			const float in = f;
			const float z = (pf ? v_lap : rd(b_lap, m_lap, pos, old_lap, lap, mu));
			f = z - (c.ap_feed_coeff * in);
			const float g = in + (c.ap_feed_coeff * f);
			const float ga = qget(g, o1), gb = qget(g, o2), gc = qget(g, o3);
			ring.st(b_lap + (pos & m_lap), scatter_row(g, ga, gb, gc, c.mix_x, c.mix_y));
		}
		const float late_out = f;
		{
			const float self = qget(f, rs), ra = qget(f, 3 - o1), rb = qget(f, 3 - o2), rc = qget(f, 3 - o3);
			ring.st(b_ll + (pos & m_ll), scatter_row(self, ra, rb, rc, c.mix_x, c.mix_y));
		}

		offset += 1;
		if (faded) {
			fade += 1.0F / FxReverb::kFadeSamples;
		}

		// ---- pan: lane k sums the 8 line outputs into its output channel(s), in line order ----
#pragma unroll
		for (int l = 0; l < 8; ++l) {
			const float d = qget(l < 4 ? early_out : late_out, l & 3);
#pragma unroll
			for (int o = 0; o < NO; ++o) {
				const uint32_t bit = 1U << (o * 8 + l);
				if (ramp & bit) {
					acc[o] += d * cur[o][l];
					cur[o][l] += stp[o][l];
				} else if (act & bit) {
					acc[o] += d * cur[o][l];
				}
			}
		}

		sub_left -= 1;
		if (sub_left == 0) {
			end_sub(c);
		}
	}

	__device__ void end(const Ctx& cx)
	{
		uint32_t* st = cx.state;
		store_words(lp, st + (kRvLp + j * 4) * kLanes);
		store_words(hp, st + (kRvHp + j * 4) * kLanes);
		st[(kRvT60 + j * 4 + 0) * kLanes] = __float_as_uint(t60s[0][0]);
		st[(kRvT60 + j * 4 + 1) * kLanes] = __float_as_uint(t60s[0][1]);
		st[(kRvT60 + j * 4 + 2) * kLanes] = __float_as_uint(t60s[1][0]);
		st[(kRvT60 + j * 4 + 3) * kLanes] = __float_as_uint(t60s[1][1]);
#pragma unroll
		for (int o = 0; o < NO; ++o) {
			const int k = j + 4 * o;
			if (k < CT) {
#pragma unroll
				for (int l = 0; l < 8; ++l) {
					st[(kRvGain + l * kMaxChannels + k) * kLanes] = __float_as_uint(cur[o][l]);
				}
			}
		}
		st[(kRvOldEt + j) * kLanes] = static_cast<uint32_t>(old_et);
		st[(kRvOldEap + j) * kLanes] = static_cast<uint32_t>(old_eap);
		st[(kRvOldEo + j) * kLanes] = static_cast<uint32_t>(old_eo);
		st[(kRvOldLt + j) * kLanes] = static_cast<uint32_t>(old_lt);
		st[(kRvOldLap + j) * kLanes] = static_cast<uint32_t>(old_lap);
		st[(kRvOldLo + j) * kLanes] = static_cast<uint32_t>(old_lo);
		if (j == 0) {
			st[(kRvScalars + 0) * kLanes] = static_cast<uint32_t>(offset);
			st[(kRvScalars + 1) * kLanes] = static_cast<uint32_t>(fade_count);
			st[(kRvScalars + 2) * kLanes] = static_cast<uint32_t>(mod_index);
			st[(kRvScalars + 3) * kLanes] = static_cast<uint32_t>(mod_range);
			st[(kRvScalars + 4) * kLanes] = __float_as_uint(mod_filter);
		}
	}
};

// ---- the fused quad kernel ---------------------------------------------------------------------------------
// Requirements checked by the host: no send shelf filter active, frames >= 2, every tile of the
// launch takes part with all its lanes (ragged tail streams only skip I/O; the arenas are padded
// to whole tiles).  CTA = 128 threads = one tile.
template <int CT, template <int> class Q0, template <int> class Q1, template <int> class Q2, template <int> class Q3>
__global__ void __launch_bounds__(kQuadThreads, OALSFX_QUAD_MIN_CTAS) quad_kernel(const __grid_constant__ MixArgs a)
{
	__shared__ float pf_smem[kPfFloatsPerSlotUser]; // one prefetching slot user (the reverb) per kernel
	const int tile = a.tiles ? static_cast<int>(a.tiles[blockIdx.x].tile) : a.tile_first + static_cast<int>(blockIdx.x);
	const int lane_in_tile = threadIdx.x >> 2;           // stream within the tile
	const int j = threadIdx.x & 3;
	const bool io_ok = tile * kLanes + lane_in_tile < a.num_streams;
	const float* src = a.src + tile * a.io_ts + lane_in_tile * a.io_ls;
	float* dst = a.dst + tile * a.io_ts + lane_in_tile * a.io_ls;

	Ctx cx[kMaxSlots];
#pragma unroll
	for (int p = 0; p < kMaxSlots; ++p) {
		cx[p].j = j;
		cx[p].state = a.slot_state[p] ? a.slot_state[p] + (static_cast<long long>(tile) * kSlotStateWords) * kLanes + lane_in_tile : nullptr;
		cx[p].ring = a.ring[p] ? a.ring[p] + static_cast<long long>(tile) * a.ring_tile_stride[p] + lane_in_tile : nullptr;
		cx[p].pf = pf_smem + threadIdx.x;
	}
	Q0<CT> q0;
	Q1<CT> q1;
	Q2<CT> q2;
	Q3<CT> q3;
	q0.begin(a.slot[0], a.aux[0], cx[0], (a.update_mask >> 0) & 1U, a.frames);
	q1.begin(a.slot[1], a.aux[1], cx[1], (a.update_mask >> 1) & 1U, a.frames);
	q2.begin(a.slot[2], a.aux[2], cx[2], (a.update_mask >> 2) & 1U, a.frames);
	q3.begin(a.slot[3], a.aux[3], cx[3], (a.update_mask >> 3) & 1U, a.frames);

	constexpr int NO = Own<CT>::n;
	float gd[CT][NO]; // direct-send gains into the channels this lane owns
#pragma unroll
	for (int c = 0; c < CT; ++c) {
		own_gains<CT>(gd[c], a.direct.gains[c], j);
	}

	float x[CT], xn[CT];
#pragma unroll
	for (int c = 0; c < CT; ++c) {
		xn[c] = io_ok ? src[c * a.io_cs] : 0.0F;
	}
	for (int i = 0; i < a.frames; ++i) {
		float acc[NO];
#pragma unroll
		for (int c = 0; c < CT; ++c) {
			x[c] = xn[c];
		}
		if (i + 1 < a.frames) {
#pragma unroll
			for (int c = 0; c < CT; ++c) {
				xn[c] = io_ok ? src[(i + 1) * a.io_fs + c * a.io_cs] : 0.0F;
			}
		}
#pragma unroll
		for (int o = 0; o < NO; ++o) {
			acc[o] = 0.0F;
		}
#pragma unroll
		for (int c = 0; c < CT; ++c) {
			own_add<CT>(acc, gd[c], x[c]);
		}
		q0.step(a.slot[0], x, acc);
		q1.step(a.slot[1], x, acc);
		q2.step(a.slot[2], x, acc);
		q3.step(a.slot[3], x, acc);
#pragma unroll
		for (int o = 0; o < NO; ++o) {
			const int k = j + 4 * o;
			if (k < CT && io_ok) {
				dst[i * a.io_fs + k * a.io_cs] = acc[o];
			}
		}
	}

	q0.end(cx[0]);
	q1.end(cx[1]);
	q2.end(cx[2]);
	q3.end(cx[3]);

	// Send filter histories: with no shelf filter active every processed send just remembers the last
	// two input samples (oalsfxpp.cpp:1038-1056).  Lane j writes the sends j and j + 4.
	uint32_t* ss = a.send_state + (static_cast<long long>(tile) * kSendStateWords) * kLanes + lane_in_tile;
	const bool live[kMaxSlots] = {!Q0<CT>::kIsNull, !Q1<CT>::kIsNull, !Q2<CT>::kIsNull, !Q3<CT>::kIsNull};
	for (int send = j; send < kSendCount; send += 4) {
		bool on = (send == 0);
#pragma unroll
		for (int p = 0; p < kMaxSlots; ++p) {
			if (send == 1 + a.aux_index[p] && live[p]) {
				on = true;
			}
		}
		if (!on) {
			continue;
		}
#pragma unroll
		for (int c = 0; c < CT; ++c) {
			// the last two input samples of the block (frames >= 2), re-read instead of carried in registers
			const float last1 = io_ok ? src[(a.frames - 1) * a.io_fs + c * a.io_cs] : 0.0F;
			const float last2 = io_ok ? src[(a.frames - 2) * a.io_fs + c * a.io_cs] : 0.0F;
			SendHist h;
			h.lp.x0 = h.lp.y0 = h.hp.x0 = h.hp.y0 = last1;
			h.lp.x1 = h.lp.y1 = h.hp.x1 = h.hp.y1 = last2;
			store_words(h, ss + (send * kMaxChannels + c) * 8 * kLanes);
		}
	}
}

} // namespace quad
} // namespace oalsfx

#endif // __CUDACC__
#endif
