// span.cuh -- the single-reverb-slot signature, BLOCK-PARALLEL in time (few streams, long blocks).
//
// With a thread per stream a block of F frames is F dependent iterations of a ~900-instruction sample
// body: ~0.5 us per frame however few streams there are (cfg1: 1024 streams = 32 tiles on 148 SMs).  But
// the reverb is a feedback delay network: every feedback path runs through a delay line, and the only
// sample-to-sample recurrences that do NOT are eight short IIR filters (the input shelves, oalsfxpp.cpp:
// 7821-7832, and the late lines' T60 filters, :7691-7719).  So, in the steady state of a preset, a SPAN of
// T consecutive frames (T <= the shortest delay) is processed in three phases by the 16 warps of a CTA,
// lanes = streams of the tile (all 32, or 16 / 8 of them when the tile is shared by 2 / 4 CTAs), warps = time:
//
//   A  (parallel over frames)  dry mix, wet encode, B->A conversion -> shared memory;
//                              the whole early-reflection stage (reads only data older than the span);
//                              late taps + late line reads -> shared memory
//   B  (serial over frames)    warps 0,1: the lp/hp shelf pairs of lines (0,1) / (2,3) -> main delay line
//                              warps 2,3: the T60 filter pairs of lines (0,1) / (2,3), in place
//   C  (parallel over frames)  late all-pass + scatter + late line feed, pan of the 8 line outputs, output
//
// Every value is computed by the same expression as in fx_reverb.cuh's sample body (same helpers, same
// order of additions), only the schedule differs -- the result is bit-identical.  What makes the parallel
// phases legal is checked on the host per coefficient block (span_frames_for): for every ring read with
// delay d against every write position of the same ring, T <= d (nothing written inside the span is read
// inside it) and d + T <= ring length (nothing read inside the span is overwritten inside it).
//
// Steady state = no parameter update pending, tap cross-fade finished (fade_count = 128), modulator quiet,
// no pan-gain ramp in any sub-chunk of the block.  The host only launches this kernel when no update is
// pending; the device verifies the rest per tile and otherwise runs the exact thread-per-stream body.
#ifndef OALSFX_SPAN_CUH
#define OALSFX_SPAN_CUH

#include "mix.cuh"
#if defined(__CUDACC__)
#include "duo.cuh"
#endif

namespace oalsfx {
namespace span {

constexpr int kWarps = 16;
constexpr int kMaxFrames = 64;            // frames per span (shared memory is sized for this)
constexpr int kThreads = kWarps * kLanes;
constexpr int shared_floats(int channels, int stream_lanes) { return kMaxFrames * (12 + channels) * stream_lanes; }

// Largest span length (a multiple of 16, <= kMaxFrames) that is legal for this coefficient block, or 0.
inline int span_frames_for(const ReverbCoef& c)
{
	if (c.mod_depth != 0.0F) {
		return 0;
	}
	for (int t = kMaxFrames; t >= 16; t -= 16) {
		bool ok = true;
		auto check = [&](int delay, int write_offset, int ring) {
			const int len = c.mask[ring] + 1;
			const int k = ((delay - write_offset) % len + len) % len;
			ok = ok && k >= t && k <= len - t;
		};
		for (int l = 0; l < 4; ++l) {
			check(c.early_tap[l], 0, 0);
			check(c.early_tap[l], c.late_feed_tap, 0);
			check(c.late_tap[l], 0, 0);
			check(c.late_tap[l], c.late_feed_tap, 0);
			check(c.early_ap_off[l], 0, 1);
			check(c.early_off[l], 0, 2);
			check(c.late_ap_off[l], 0, 3);
			check(c.late_off[l], 0, 4);
		}
		if (ok) {
			return t;
		}
	}
	return 0;
}

#if defined(__CUDACC__)

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Captures the wet bus SlotRunner::step encodes for slot position 0 (same code path as every other kernel).
struct FxWetProbe {
	static constexpr bool kIsNull = false;
	float wet[kWetChannels];
	template <int CT, bool FAST>
	__device__ __forceinline__ void step(const SlotCoef&, const float* w, float*, int)
	{
#pragma unroll
		for (int k = 0; k < kWetChannels; ++k) {
			wet[k] = w[k];
		}
	}
};

// SL: streams of the tile one CTA handles (32, 16 or 8).  With SL < 32 a tile is shared by 32 / SL CTAs -- on as many
// SMs -- and a warp covers FR = 32 / SL consecutive frames of those streams: lane = frame * SL + stream.  Few tiles
// are instruction-bound on the one SM each of them gets (profiles/r01_ncu_span_top_stalls.txt); streams are
// independent, so the split needs no communication.
template <int CT, int SL>
__global__ void __launch_bounds__(kThreads, 1) span_reverb_kernel(const __grid_constant__ MixArgs a)
{
	using R = FxReverbTail;
	constexpr int FR = kLanes / SL, SPLIT = kLanes / SL;
	extern __shared__ __align__(16) float dyn[];  // shared_floats(CT, SL) floats, each array [line][frame][stream]
	float* const sA = dyn;                          // A-format lines before the shelves (A -> B)
	float* const sL = dyn + 4 * kMaxFrames * SL;    // late lines before / after the T60 filters (A -> B -> C)
	float* const sE = dyn + 8 * kMaxFrames * SL;    // early line outputs (A -> C)
	float* const sO = dyn + 12 * kMaxFrames * SL;   // bus after the dry mix (A -> C)
#define OALSFX_SPAN_AT(base, line, t) base[((line) * kMaxFrames + (t)) * SL + sl]

	const int tile_slot = static_cast<int>(blockIdx.x) / SPLIT;
	const int tile = a.tiles ? static_cast<int>(a.tiles[tile_slot].tile) : a.tile_first + tile_slot;
	const int sl = static_cast<int>(threadIdx.x % kLanes) % SL;          // stream within this CTA's share
	const int f = static_cast<int>(threadIdx.x % kLanes) / SL;           // frame within the warp's FR frames
	const int lane = (static_cast<int>(blockIdx.x) % SPLIT) * SL + sl;   // stream within the tile
	const int w = threadIdx.x / kLanes;
	const bool io_ok = tile * kLanes + lane < a.num_streams;
	const ReverbCoef& c = a.slot[0].u.reverb;
	const float* src = a.src + tile * a.io_ts + lane * a.io_ls;
	float* dst = a.dst + tile * a.io_ts + lane * a.io_ls;
	uint32_t* st = a.slot_state[0] + (static_cast<long long>(tile) * kSlotStateWords) * kLanes + lane;
	uint32_t* ss = a.send_state + (static_cast<long long>(tile) * kSendStateWords) * kLanes + lane;
	const int T = a.span_frames;

	// ---- steady state?  (per lane, then the whole tile) ----
	const int32_t offset0 = static_cast<int32_t>(st[(R::kWScalars + 0) * kLanes]);
	const int32_t fade_count = static_cast<int32_t>(st[(R::kWScalars + 1) * kLanes]);
	int32_t mod_index = static_cast<int32_t>(st[(R::kWScalars + 2) * kLanes]);
	int32_t mod_range = static_cast<int32_t>(st[(R::kWScalars + 3) * kLanes]);
	const float mod_filter = word_as_float(st[(R::kWScalars + 4) * kLanes]);
	if (mod_range == 0) {
		mod_range = 1;
	}
	float gain[8][CT];
	bool ok = T >= 1 && a.update_mask == 0 && fade_count >= R::kFadeSamples && c.mod_depth == 0.0F && mod_filter == 0.0F;
#pragma unroll
	for (int l = 0; l < 8; ++l) {
#pragma unroll
		for (int k = 0; k < CT; ++k) {
			gain[l][k] = word_as_float(st[(R::kWGain + l * kMaxChannels + k) * kLanes]);
		}
	}
	// no pan-gain ramp in any sub-chunk (begin_sub: sub-chunks of <= 256 frames, step = (target - gain) / frames left)
	for (int base = 0; base < a.frames; base += R::kMaxUpdate) {
		const float delta = 1.0F / static_cast<float>(a.frames - base);
#pragma unroll
		for (int l = 0; l < 8; ++l) {
			const float* target = (l < 4 ? c.pan_early[l] : c.pan_late[l - 4]);
#pragma unroll
			for (int k = 0; k < CT; ++k) {
				ok = ok && !(fabsf((target[k] - gain[l][k]) * delta) > FLT_EPSILON);
			}
		}
	}
	if (!__syncthreads_and(ok || !io_ok)) {
		if (w == 0 && f == 0 && io_ok) {
			mix_stream<CT, false, FxReverb, FxNull, FxNull, FxNull>(a, tile, lane, nullptr);
		}
		return;
	}

	R fx;
	fx.ring.p = a.ring[0] + static_cast<long long>(tile) * a.ring_tile_stride[0] + lane;
	fx.faded = false;
	fx.st_mem = st;
	const LaneMem ring = fx.ring;
	const int main_len = c.mask[0] + 1, main0 = c.ring_base[0], main_mask = c.mask[0];
	const int eline_len = c.mask[2] + 1, eline0 = c.ring_base[2], eline_mask = c.mask[2];
	const int lline_len = c.mask[4] + 1, lline0 = c.ring_base[4], lline_mask = c.mask[4];

	// serial-phase state: warps 0,1 own the shelves of lines (2w, 2w+1), warps 2,3 the T60 filters of lines (2h, 2h+1)
	BiquadHist lp[2], hp[2];
	F2 t60p[2][2];
	if (w < 2) {
#pragma unroll
		for (int i = 0; i < 2; ++i) {
			load_words(lp[i], st + (R::kWLp + (2 * w + i) * 4) * kLanes);
			load_words(hp[i], st + (R::kWHp + (2 * w + i) * 4) * kLanes);
		}
	} else if (w < 4) {
		const int h = w - 2;
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			t60p[q >> 1][q & 1] = f2(word_as_float(st[(R::kWT60 + (2 * h) * 4 + q) * kLanes]),
				word_as_float(st[(R::kWT60 + (2 * h + 1) * 4 + q) * kLanes]));
		}
	}

	// All ring reads of a frame, requested before any of its arithmetic (and before its ring stores, which the
	// compiler must assume to alias): one memory round trip per frame instead of five.
	struct Taps { float early[4], eap[4], eline[4], late[4], lline[4]; };
	auto load_taps = [&](int pos, Taps& k) {
		const int eap_len = c.mask[1] + 1, eap0 = c.ring_base[1], eap_mask = c.mask[1];
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			k.early[l] = ring.ld(main0 + l * main_len + ((pos - c.early_tap[l]) & main_mask));
			k.eap[l] = ring.ld(eap0 + l * eap_len + ((pos - c.early_ap_off[l]) & eap_mask));
			k.eline[l] = ring.ld(eline0 + l * eline_len + ((pos - c.early_off[l]) & eline_mask));
			k.late[l] = ring.ld(main0 + l * main_len + ((pos - c.late_tap[l]) & main_mask));
			k.lline[l] = ring.ld(lline0 + l * lline_len + ((pos - c.late_off[l]) & lline_mask));
		}
	};
	// vector_allpass_x with the taps already read (FxReverbT::vector_allpass2, fx_reverb.cuh)
	auto allpass = [&](F2& va, F2& vb, const float* tp, int ring_idx, int pos) {
		const int len = c.mask[ring_idx] + 1, word0 = c.ring_base[ring_idx], mask = c.mask[ring_idx];
		const F2 ta = f2(tp[0], tp[1]), tb = f2(tp[2], tp[3]);
		const F2 ina = va, inb = vb;
		va = ta - (ina * c.ap_feed_coeff);
		vb = tb - (inb * c.ap_feed_coeff);
		F2 fa = ina + (va * c.ap_feed_coeff);
		F2 fb = inb + (vb * c.ap_feed_coeff);
		R::scatter2(fa, fb, c.mix_x, c.mix_y);
		ring.st(word0 + 0 * len + (pos & mask), f2_lo(fa));
		ring.st(word0 + 1 * len + (pos & mask), f2_hi(fa));
		ring.st(word0 + 2 * len + (pos & mask), f2_lo(fb));
		ring.st(word0 + 3 * len + (pos & mask), f2_hi(fb));
	};
	// Phase A of one frame, from its input and taps.
	SlotRunner<CT, false, FxWetProbe> probe;
	auto phase_a = [&](int t, int pos, const float* x, const Taps& k) {
		float acc[CT];
#pragma unroll
		for (int ch = 0; ch < CT; ++ch) {
			acc[ch] = 0.0F;
		}
#pragma unroll
		for (int ch = 0; ch < CT; ++ch) {
			pan_add<CT, true>(acc, CT, a.direct.gains[ch], x[ch]); // direct send (oalsfxpp.cpp:2924-2950)
		}
		probe.step(a, 0, x, acc);
		const float* wet = probe.fx.wet;
#pragma unroll
		for (int ch = 0; ch < CT; ++ch) {
			OALSFX_SPAN_AT(sO, ch, t) = acc[ch];
		}
		{
			// B-format -> A-format, as reverb_input_stage (fx_reverb.cuh)
			constexpr float q = 0.288675134595F;
			const F2 zero = f2(0.0F, 0.0F);
			const F2 p0 = f2_bcast(wet[0] * q), p3 = f2_bcast(wet[3] * q);
			const F2 p1 = f2_bcast(wet[1]) * f2(q, -q);
			const F2 p2 = f2_bcast(wet[2]) * f2(q, -q);
			const F2 a01 = (((zero + p0) + p1) + p2) + p3;
			const F2 a23 = (((zero + p0) + p1) - p2) - p3;
			OALSFX_SPAN_AT(sA, 0, t) = f2_lo(a01);
			OALSFX_SPAN_AT(sA, 1, t) = f2_hi(a01);
			OALSFX_SPAN_AT(sA, 2, t) = f2_lo(a23);
			OALSFX_SPAN_AT(sA, 3, t) = f2_hi(a23);
		}
		// early reflections (the EARLY half of FxReverbT::body)
		F2 fa = f2(k.early[0], k.early[1]) * f2(c.early_tap_coeff[0], c.early_tap_coeff[1]);
		F2 fb = f2(k.early[2], k.early[3]) * f2(c.early_tap_coeff[2], c.early_tap_coeff[3]);
		allpass(fa, fb, k.eap, 1, pos);
		ring.st(eline0 + 0 * eline_len + (pos & eline_mask), f2_hi(fb));
		ring.st(eline0 + 1 * eline_len + (pos & eline_mask), f2_lo(fb));
		ring.st(eline0 + 2 * eline_len + (pos & eline_mask), f2_hi(fa));
		ring.st(eline0 + 3 * eline_len + (pos & eline_mask), f2_lo(fa));
		fa = fa + (f2(k.eline[0], k.eline[1]) * f2(c.early_coeff[0], c.early_coeff[1]));
		fb = fb + (f2(k.eline[2], k.eline[3]) * f2(c.early_coeff[2], c.early_coeff[3]));
		OALSFX_SPAN_AT(sE, 0, t) = f2_lo(fa);
		OALSFX_SPAN_AT(sE, 1, t) = f2_hi(fa);
		OALSFX_SPAN_AT(sE, 2, t) = f2_lo(fb);
		OALSFX_SPAN_AT(sE, 3, t) = f2_hi(fb);
		{
			F2 ra = fa, rb = fb;
			R::scatter2_reversed(ra, rb, c.mix_x, c.mix_y);
			const int feed = (pos - c.late_feed_tap) & main_mask;
			ring.st(main0 + 0 * main_len + feed, f2_hi(rb));
			ring.st(main0 + 1 * main_len + feed, f2_lo(rb));
			ring.st(main0 + 2 * main_len + feed, f2_hi(ra));
			ring.st(main0 + 3 * main_len + feed, f2_lo(ra));
		}
		// late reverb up to the T60 filters (modulation delay 0: steady state)
		fa = f2(k.late[0], k.late[1]) * c.density_gain;
		fb = f2(k.late[2], k.late[3]) * c.density_gain;
		fa = fa + f2(k.lline[0], k.lline[1]);
		fb = fb + f2(k.lline[2], k.lline[3]);
		OALSFX_SPAN_AT(sL, 0, t) = f2_lo(fa);
		OALSFX_SPAN_AT(sL, 1, t) = f2_hi(fa);
		OALSFX_SPAN_AT(sL, 2, t) = f2_lo(fb);
		OALSFX_SPAN_AT(sL, 3, t) = f2_hi(fb);
	};
	// Next span's ring rows and input rows -> L2, requested by the warps that idle during phase B.
	auto prefetch_span = [&](int first_next) {
		const int count = min(T, a.frames - first_next);
		const int lap_len = c.mask[3] + 1, lap0 = c.ring_base[3], lap_mask = c.mask[3];
		const int eap_len = c.mask[1] + 1, eap0 = c.ring_base[1], eap_mask = c.mask[1];
		for (int t = (w - 4) * FR + f; t < count; t += (kWarps - 4) * FR) {
			const int pos = offset0 + first_next + t;
#pragma unroll
			for (int l = 0; l < 4; ++l) {
				prefetch_l2(ring.p + static_cast<unsigned>(main0 + l * main_len + ((pos - c.early_tap[l]) & main_mask)) * kLanes);
				prefetch_l2(ring.p + static_cast<unsigned>(eap0 + l * eap_len + ((pos - c.early_ap_off[l]) & eap_mask)) * kLanes);
				prefetch_l2(ring.p + static_cast<unsigned>(eline0 + l * eline_len + ((pos - c.early_off[l]) & eline_mask)) * kLanes);
				prefetch_l2(ring.p + static_cast<unsigned>(main0 + l * main_len + ((pos - c.late_tap[l]) & main_mask)) * kLanes);
				prefetch_l2(ring.p + static_cast<unsigned>(lline0 + l * lline_len + ((pos - c.late_off[l]) & lline_mask)) * kLanes);
				prefetch_l2(ring.p + static_cast<unsigned>(lap0 + l * lap_len + ((pos - c.late_ap_off[l]) & lap_mask)) * kLanes);
			}
			if (io_ok) {
				prefetch_l2(src + (first_next + t) * a.io_fs);
			}
		}
	};

	for (int first = 0; first < a.frames; first += T) {
		const int count = min(T, a.frames - first);
		// ---- A: everything that only reads data older than the span (two frames per iteration) ----
		for (int t = w * FR + f; t < count; t += 2 * kWarps * FR) {
			const int t2 = t + kWarps * FR;
			const bool two = t2 < count;
			const int pos = offset0 + first + t;
			float x0[CT], x1[CT];
			Taps k0, k1;
#pragma unroll
			for (int ch = 0; ch < CT; ++ch) {
				x0[ch] = io_ok ? src[(first + t) * a.io_fs + ch * a.io_cs] : 0.0F;
				x1[ch] = (io_ok && two) ? src[(first + t2) * a.io_fs + ch * a.io_cs] : 0.0F;
			}
			load_taps(pos, k0);
			if (two) {
				load_taps(pos + kWarps * FR, k1);
			}
			phase_a(t, pos, x0, k0);
			if (two) {
				phase_a(t2, pos + kWarps * FR, x1, k1);
			}
		}
		__syncthreads();
		// ---- B: the recurrences, one warp per line pair; the other warps pull the next span into L2 ----
		if (w < 2 && f == 0) {
#pragma unroll 4
			for (int t = 0; t < count; ++t) {
				const int pos = offset0 + first + t;
				F2 v = f2(OALSFX_SPAN_AT(sA, 2 * w, t), OALSFX_SPAN_AT(sA, 2 * w + 1, t));
				v = biquad_step2(c.lp, lp[0], lp[1], v);
				if (c.is_eax) {
					v = biquad_step2(c.hp, hp[0], hp[1], v);
				}
				ring.st(main0 + (2 * w) * main_len + (pos & main_mask), f2_lo(v));
				ring.st(main0 + (2 * w + 1) * main_len + (pos & main_mask), f2_hi(v));
			}
		} else if (w < 4 && f == 0) {
			const int j = 2 * (w - 2);
#pragma unroll 4
			for (int t = 0; t < count; ++t) {
				// late_t60_filter (oalsfxpp.cpp:7691-7719), as in FxReverbT::body
				const F2 in = f2(OALSFX_SPAN_AT(sL, j, t), OALSFX_SPAN_AT(sL, j + 1, t));
				const F2 o1 = (f2(c.t60_lf[j][0], c.t60_lf[j + 1][0]) * in) + (f2(c.t60_lf[j][1], c.t60_lf[j + 1][1]) * t60p[0][0]) +
					(f2(c.t60_lf[j][2], c.t60_lf[j + 1][2]) * t60p[0][1]);
				t60p[0][0] = in;
				t60p[0][1] = o1;
				const F2 o2 = (f2(c.t60_hf[j][0], c.t60_hf[j + 1][0]) * o1) + (f2(c.t60_hf[j][1], c.t60_hf[j + 1][1]) * t60p[1][0]) +
					(f2(c.t60_hf[j][2], c.t60_hf[j + 1][2]) * t60p[1][1]);
				t60p[1][0] = o1;
				t60p[1][1] = o2;
				const F2 out = f2(c.t60_mid[j], c.t60_mid[j + 1]) * o2;
				OALSFX_SPAN_AT(sL, j, t) = f2_lo(out);
				OALSFX_SPAN_AT(sL, j + 1, t) = f2_hi(out);
			}
		} else if (w >= 4 && first + T < a.frames) {
			prefetch_span(first + T);
		}
		__syncthreads();
		// ---- C: the rest of the late reverb, pan, output (two frames per iteration) ----
		for (int t = w * FR + f; t < count; t += 2 * kWarps * FR) {
			const int lap_len = c.mask[3] + 1, lap0 = c.ring_base[3], lap_mask = c.mask[3];
			float tp[2][4];
#pragma unroll
			for (int u = 0; u < 2; ++u) {
#pragma unroll
				for (int l = 0; l < 4; ++l) {
					const int pos = offset0 + first + t + u * kWarps * FR;
					tp[u][l] = (u == 0 || t + kWarps * FR < count) ? ring.ld(lap0 + l * lap_len + ((pos - c.late_ap_off[l]) & lap_mask)) : 0.0F;
				}
			}
#pragma unroll
			for (int u = 0; u < 2; ++u) {
				const int tt = t + u * kWarps * FR;
				if (tt < count) {
					const int n = first + tt;
					const int pos = offset0 + n;
					F2 fa = f2(OALSFX_SPAN_AT(sL, 0, tt), OALSFX_SPAN_AT(sL, 1, tt)), fb = f2(OALSFX_SPAN_AT(sL, 2, tt), OALSFX_SPAN_AT(sL, 3, tt));
					allpass(fa, fb, tp[u], 3, pos);
					float out8[8];
#pragma unroll
					for (int l = 0; l < 4; ++l) {
						out8[l] = OALSFX_SPAN_AT(sE, l, tt);
					}
					out8[4] = f2_lo(fa);
					out8[5] = f2_hi(fa);
					out8[6] = f2_lo(fb);
					out8[7] = f2_hi(fb);
					{
						F2 ra = fa, rb = fb;
						R::scatter2_reversed(ra, rb, c.mix_x, c.mix_y);
						ring.st(lline0 + 0 * lline_len + (pos & lline_mask), f2_hi(rb));
						ring.st(lline0 + 1 * lline_len + (pos & lline_mask), f2_lo(rb));
						ring.st(lline0 + 2 * lline_len + (pos & lline_mask), f2_hi(ra));
						ring.st(lline0 + 3 * lline_len + (pos & lline_mask), f2_lo(ra));
					}
					// pan with static gains (oalsfxpp.cpp:6142-6166, 2752-2798): inaudible gains are skipped
					float acc[CT];
#pragma unroll
					for (int ch = 0; ch < CT; ++ch) {
						acc[ch] = OALSFX_SPAN_AT(sO, ch, tt);
					}
#pragma unroll
					for (int l = 0; l < 8; ++l) {
#pragma unroll
						for (int ch = 0; ch < CT; ++ch) {
							if (audible(gain[l][ch])) {
								acc[ch] += out8[l] * gain[l][ch];
							}
						}
					}
					if (io_ok) {
#pragma unroll
						for (int ch = 0; ch < CT; ++ch) {
							dst[n * a.io_fs + ch * a.io_cs] = acc[ch];
						}
					}
				}
			}
		}
		__syncthreads();
	}

	// ---- state ----
	if (f != 0) {
		return;
	}
	if (w < 2) {
#pragma unroll
		for (int i = 0; i < 2; ++i) {
			store_words(lp[i], st + (R::kWLp + (2 * w + i) * 4) * kLanes);
			store_words(hp[i], st + (R::kWHp + (2 * w + i) * 4) * kLanes);
		}
	} else if (w < 4) {
		const int h = w - 2;
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			st[(R::kWT60 + (2 * h) * 4 + q) * kLanes] = float_as_word(f2_lo(t60p[q >> 1][q & 1]));
			st[(R::kWT60 + (2 * h + 1) * 4 + q) * kLanes] = float_as_word(f2_hi(t60p[q >> 1][q & 1]));
		}
	} else if (w == 4) {
		st[(R::kWScalars + 0) * kLanes] = static_cast<uint32_t>(offset0 + a.frames);
		// the quiet modulator only advances its index (FxReverbT::body): +1 per frame, wrapping at the range
		mod_index = static_cast<int32_t>((static_cast<long long>(mod_index) + a.frames) % mod_range);
		st[(R::kWScalars + 2) * kLanes] = static_cast<uint32_t>(mod_index);
		st[(R::kWScalars + 3) * kLanes] = static_cast<uint32_t>(mod_range);
	} else if (w == 5) {
		duo::store_passthrough_history<CT>(ss, 0, src, a, io_ok);
		duo::store_passthrough_history<CT>(ss, 1 + a.aux_index[0], src, a, io_ok);
	}
}

#undef OALSFX_SPAN_AT

#endif // __CUDACC__

} // namespace span
} // namespace oalsfx

#endif
