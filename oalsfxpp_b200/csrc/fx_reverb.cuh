// fx_reverb.cuh -- Reverb / EAX reverb, one thread per stream (included by fx.cuh).
//
// reference: do_process oalsfxpp.cpp:6078-6170, (eax_)verb_pass :7814-7903,
// early_reflection_x :7625-7672, late_reverb_x :7735-7794, vector_allpass_x :7533-7562,
// vector_partial_scatter :7510-7521, late_t60_filter :7691-7719, calc_modulation_delays :7443-7470,
// MixHelpers::mix :2752-2798.
//
// The reference runs each <=256-sample sub-chunk in phases (input filter, early, late, pan-mix);
// here the phases are interleaved per sample, which yields identical values because every ring
// read in a phase targets a position that no *later* sample of an earlier phase writes (all main
// line taps are >= 0, the late taps are >= the late feed tap, and the main ring carries 256 spare
// frames, oalsfxpp.cpp:6573).
//
// Memory system (device build): the 24 ring reads of a sample are the only long-latency operations
// of the path.  When every delay exceeds the prefetch depth and no cross-fade / modulation is
// active (the steady state of every preset), the reads are BATCHED: one 16-byte cp.async.cg per lane
// moves, for one tap, four consecutive ring positions of the whole tile (a contiguous 512-byte
// run: lane = position * 8 + stream quad) into a per-warp shared-memory window, one batch (4
// samples x 24 taps) ahead of the arithmetic.  That is a quarter of the copy instructions and
// address arithmetic of per-sample requests, longer DRAM bursts, no registers held by loads in
// flight, and -- .cg bypasses L1 -- no L1 lines tied up as landing buffers.  Otherwise (first 128
// samples after a tap change, modulated presets, tiny delays) the rings are read directly at the
// point of use.  Both paths read the same values.
#ifndef OALSFX_FX_REVERB_CUH
#define OALSFX_FX_REVERB_CUH

namespace oalsfx {

// Prefetch window geometry: [ring position & 7][tap 0..23][lane]; taps = {early, early all-pass,
// early line, late, late line, late all-pass} x 4 lines.  Batches of kPfBatch positions: the one
// being consumed and the one in flight.
constexpr int kPfBatch = 4;                 // ring positions per batched copy (32 lanes x 16 B = 4 lines)
#ifndef OALSFX_PF_SLOTS
#define OALSFX_PF_SLOTS 6
#endif
// Window rows (ring positions).  8 = two whole batches, the next one requested when the current one
// starts.  6 (7) = the next batch is requested once two (one) positions of the current one have been
// consumed -- their rows plus the spare ones receive it: less shared memory per warp, i.e. more L1,
// for a shorter lead over the arithmetic.
constexpr int kPfSlots = OALSFX_PF_SLOTS;
static_assert(kPfSlots >= 6 && kPfSlots <= 8, "window of 6, 7 or 8 ring positions");
constexpr int kPfIssueAt = 8 - kPfSlots;    // position within the batch at which the next batch is requested
constexpr int kPfDepth = 7;                 // bound on how far beyond the current position a request may reach
constexpr int kPfTaps = 24;                 // whole effect; a half (FxReverbT<.., EARLY, LATE>) uses 12
constexpr int kPfWarpFloats = kPfSlots * kPfTaps * kLanes; // 24 KiB per reverb warp

// Input stage of the reverb, shared by the whole effect (FxReverbT<true>) and by FxReverbInput,
// which lets another warp run it ahead of the rest (duo.cuh): B-format -> A-format (mix_row with
// the b2a matrix, oalsfxpp.cpp:6099-6113, 6377-6383), the master shelf filter(s), and the feed of the
// main delay line (oalsfxpp.cpp:7821-7832 / 7867-7879).
OALSFX_HD F2 biquad_step2(const Biquad& c, BiquadHist& ha, BiquadHist& hb, F2 x)
{
	// Two filters with the same coefficients (FilterState::process, oalsfxpp.cpp:984-1036), one per half.
	const F2 x0 = f2(ha.x0, hb.x0), x1 = f2(ha.x1, hb.x1), y0 = f2(ha.y0, hb.y0), y1 = f2(ha.y1, hb.y1);
	const F2 y = (x * c.b0) + (x0 * c.b1) + (x1 * c.b2) - (y0 * c.a1) - (y1 * c.a2);
	ha.x1 = ha.x0;
	hb.x1 = hb.x0;
	ha.x0 = f2_lo(x);
	hb.x0 = f2_hi(x);
	ha.y1 = ha.y0;
	hb.y1 = hb.y0;
	ha.y0 = f2_lo(y);
	hb.y0 = f2_hi(y);
	return y;
}

OALSFX_HD void reverb_input_stage(const ReverbCoef& c, const float* wet, BiquadHist* lp, BiquadHist* hp,
	const LaneMem& ring, int pos)
{
	const int main_len = c.mask[0] + 1, main0 = c.ring_base[0], main_mask = c.mask[0];
	// Every b2a entry is +-q, and w * (-q) == -(w * q) bit for bit, so the four products are formed
	// once and added with the row's signs, in the reference's order k = 0..3 starting from 0:
	//   line 0: + + + +    line 1: + - - +    line 2: + + - -    line 3: + - + -
	// Lines (0,1) and (2,3) are computed as pairs; a row's +-p_k is the matching half of (p_k, -p_k).
	constexpr float q = 0.288675134595F;
	const F2 zero = f2(0.0F, 0.0F);
	const F2 p0 = f2_bcast(wet[0] * q), p3 = f2_bcast(wet[3] * q);
	const F2 p1 = f2_bcast(wet[1]) * f2(q, -q);   // (+p1, -p1)
	const F2 p2 = f2_bcast(wet[2]) * f2(q, -q);   // (+p2, -p2)
	F2 a01 = (((zero + p0) + p1) + p2) + p3;      // (p0+p1+p2+p3, p0-p1-p2+p3)
	F2 a23 = (((zero + p0) + p1) - p2) - p3;      // (p0+p1-p2-p3, p0-p1+p2-p3)
	a01 = biquad_step2(c.lp, lp[0], lp[1], a01);
	a23 = biquad_step2(c.lp, lp[2], lp[3], a23);
	if (c.is_eax) {
		a01 = biquad_step2(c.hp, hp[0], hp[1], a01);
		a23 = biquad_step2(c.hp, hp[2], hp[3], a23);
	}
	ring.st(main0 + 0 * main_len + (pos & main_mask), f2_lo(a01));
	ring.st(main0 + 1 * main_len + (pos & main_mask), f2_hi(a01));
	ring.st(main0 + 2 * main_len + (pos & main_mask), f2_lo(a23));
	ring.st(main0 + 3 * main_len + (pos & main_mask), f2_hi(a23));
}

// INPUT = false: the input stage (and the lp/hp filter history words of the state) belong to a
// FxReverbInput running in another warp; `wet` is not read.
// EARLY / LATE: which half of the sample body this instance runs.  The halves only meet in the main
// delay line (early -> late feed) and in the order of the pan adds (early lines first), so they can
// run in two warps, the late one behind the early one (quartet.cuh).  Each half keeps its own copy of
// the block-partition / cross-fade bookkeeping (it evolves identically in both), owns its lines' pan
// gains, its three OLD tap groups and -- the late half -- the T60 / modulator state and the scalars.
template <bool INPUT, bool EARLY = true, bool LATE = true>
struct FxReverbT {
	static_assert(EARLY || LATE, "a reverb instance runs at least one half");
	static_assert(!INPUT || (EARLY && LATE), "the input stage is only bundled with the whole effect");
	static constexpr int kLine0 = EARLY ? 0 : 4, kLine1 = LATE ? 8 : 4;     // pan lines [kLine0, kLine1)
	static constexpr int kTap0 = EARLY ? 0 : 12;                            // first window tap of this instance
	static constexpr int kTaps = (EARLY ? 12 : 0) + (LATE ? 12 : 0);        // window taps per ring position
	static constexpr int kWindowFloats = kPfSlots * kTaps * kLanes;
	// Layout of the slot state in HBM (words, per lane).  Only the hot part lives in registers.
	struct State {
		BiquadHist lp[4], hp[4];
		float t60[4][2][2];
		float cur_gain[8][kMaxChannels]; // early 0..3, late 4..7: running pan gains (oalsfxpp.cpp:6142-6166)
		int32_t old_early_tap[4], old_early_ap[4], old_early_off[4];
		int32_t old_late_tap[4], old_late_ap[4], old_late_off[4];
		int32_t offset, fade_count, mod_index, mod_range;
		float mod_filter;
	};
	static constexpr int kStateWords = sizeof(State) / 4;
	static constexpr int kWLp = 0, kWHp = 16, kWT60 = 32, kWGain = 48, kWOld = 48 + 8 * kMaxChannels, kWScalars = kWOld + 24;
	static_assert(kWScalars + 5 == kStateWords, "state layout");
	static constexpr bool kIsNull = false;
	static constexpr int kFadeSamples = 128;  // oalsfxpp.cpp:6187
	static constexpr int kMaxUpdate = 256;    // oalsfxpp.cpp:6181

	// hot state
	BiquadHist lp[4], hp[4];
	F2 t60p[2][2][2];            // late T60 filter states of lines (2h, 2h+1): [h][section][x1 / y1]
	float cur_gain[8][kMaxChannels];
	int32_t offset, fade_count, mod_index, mod_range;
	float mod_filter;
	// cold state stays in memory: the OLD tap sets (read only while cross-fading / at an update)
	uint32_t* st_mem;
	LaneMem ring;
	int32_t block_frames, base, sub_left, sub_todo;
	float fade;
	bool faded;
	float step_gain[8][kMaxChannels];
	uint32_t ramp_mask[2], active_mask[2]; // bit (line % 4) * 8 + k, word = line / 4
	// prefetch pipeline (device build; pf_col == nullptr disables it)
	float* pf_col;
	const float* pf_cur;
	bool primed, can_pf;
	int32_t pf_row4;   // window row of the first position of the batch holding the current position
	bool pan_static;   // this sub-chunk: no gain ramps and every one of the 8 x C pan gains is audible or an exact zero
	bool pan_ramp_all; // this sub-chunk: every one of this instance's pan gains ramps (a gain change: all of them scale)
	uint32_t direct_groups; // this sub-chunk: tap groups (bit = OLD-tap group id) the window cannot serve, read in place

	unsigned pf_s;     // shared-space address of the warp's window (lane offset removed), device build
	OALSFX_HD void set_prefetch(float* column)
	{
		pf_col = column;
#if defined(__CUDA_ARCH__)
		pf_s = column ? smem_addr(column) - (threadIdx.x % kLanes) * 4U : 0U;
#endif
	}
	OALSFX_HD void prefetch_issue(const SlotCoef&, int) {} // the reverb runs its own pipeline inside step()
	OALSFX_HD void prefetch_next(const SlotCoef&) {}

	OALSFX_HD int32_t old_tap(int group, int line) const
	{
		return static_cast<int32_t>(st_mem[(kWOld + group * 4 + line) * kLanes]);
	}

	template <int CT>
	OALSFX_HD void begin(const SlotCoef& sc, uint32_t* st, float* ring_p, bool update, int frames, int channels)
	{
		const ReverbCoef& c = sc.u.reverb;
		st_mem = st;
		ring.p = ring_p;
		OALSFX_UNROLL
		for (int l = 0; l < 4; ++l) {
			if (INPUT) {
				load_words(lp[l], st + (kWLp + l * 4) * kLanes);
				load_words(hp[l], st + (kWHp + l * 4) * kLanes);
			}
		}
		OALSFX_UNROLL
		for (int h = 0; h < 2; ++h) {
			OALSFX_UNROLL
			for (int w = 0; w < 4; ++w) {
				if (LATE) {
					t60p[h][w >> 1][w & 1] = f2(word_as_float(st[(kWT60 + (2 * h) * 4 + w) * kLanes]),
						word_as_float(st[(kWT60 + (2 * h + 1) * 4 + w) * kLanes]));
				}
			}
		}
		OALSFX_UNROLL
		for (int l = kLine0; l < kLine1; ++l) {
			OALSFX_UNROLL
			for (int k = 0; k < (CT ? CT : kMaxChannels); ++k) {
				if (CT || k < channels) {
					cur_gain[l][k] = word_as_float(st[(kWGain + l * kMaxChannels + k) * kLanes]);
				}
			}
		}
		offset = static_cast<int32_t>(st[(kWScalars + 0) * kLanes]);
		fade_count = static_cast<int32_t>(st[(kWScalars + 1) * kLanes]);
		mod_index = static_cast<int32_t>(st[(kWScalars + 2) * kLanes]);
		mod_range = static_cast<int32_t>(st[(kWScalars + 3) * kLanes]);
		mod_filter = word_as_float(st[(kWScalars + 4) * kLanes]);
		if (mod_range == 0) { // do_construct: mod_.range_ = 1 (oalsfxpp.cpp:5879); state memory is zero-filled
			mod_range = 1;
		}
		if (update) {
			// update_modulator (oalsfxpp.cpp:7028-7030)
			if (LATE) {
				mod_index = static_cast<int32_t>(mod_index * static_cast<int64_t>(c.mod_range) / mod_range);
				mod_range = c.mod_range;
			}
			// "Determine if delay-line cross-fading is required" (oalsfxpp.cpp:6061-6075)
			bool differs = false;
			for (int i = 0; i < 4; ++i) {
				differs = differs || c.early_tap[i] != old_tap(0, i) || c.early_ap_off[i] != old_tap(1, i) ||
					c.early_off[i] != old_tap(2, i) || c.late_tap[i] != old_tap(3, i) ||
					c.late_ap_off[i] != old_tap(4, i) || c.late_off[i] != old_tap(5, i);
			}
			if (differs) {
				fade_count = 0;
			}
		}
		block_frames = frames;
		base = 0;
		sub_left = 0;
		sub_todo = 0;
		fade = static_cast<float>(fade_count) / kFadeSamples;
		faded = false;
		pf_cur = nullptr;
		primed = false;
		can_pf = false;
#if defined(__CUDA_ARCH__)
		// A window is only handed to whole tiles (all 32 lanes running); the batches also need every
		// stream of the tile at the same ring position (streams created together, the normal case).
		if (pf_col != nullptr && !__all_sync(0xFFFFFFFFU, offset == __shfl_sync(0xFFFFFFFFU, offset, 0))) {
			pf_col = nullptr;
		}
#endif
	}

	// Sub-chunk prologue: size (oalsfxpp.cpp:6088-6096) and pan-gain stepping (MixHelpers::mix,
	// oalsfxpp.cpp:2762-2768) for the 8 line outputs.
	template <int CT>
	OALSFX_HD void begin_sub(const ReverbCoef& c, int channels)
	{
		int todo = block_frames - base;
		if (todo > kMaxUpdate) {
			todo = kMaxUpdate;
		}
		if (kFadeSamples - fade_count > 0 && todo > kFadeSamples - fade_count) {
			todo = kFadeSamples - fade_count;
		}
		sub_todo = todo;
		sub_left = todo;
		faded = fade < 1.0F;
		// Reading kPfDepth samples ahead is legal when nothing written during those samples can be what
		// the prefetch reads: every delay > kPfDepth, the late taps that far beyond the late feed write,
		// no cross-fade (reads both tap sets), no modulation (the late line read position moves).
		// A tap group that breaks the rule (a zero reflections / late delay, the modulated late line of the "underwater"
		// kind of preset) is read in place instead, the other groups keep the window ("mixed" body).
		const bool ok = pf_col != nullptr && !faded;
		uint32_t direct = (LATE && !(c.mod_depth == 0.0F && mod_filter == 0.0F)) ? (1U << 5) : 0U;
		OALSFX_UNROLL
		for (int l = 0; l < 4; ++l) {
			if (EARLY) {
				direct |= (c.early_tap[l] > kPfDepth ? 0U : 1U << 0) | (c.early_ap_off[l] > kPfDepth ? 0U : 1U << 1) |
					(c.early_off[l] > kPfDepth ? 0U : 1U << 2);
			}
			if (LATE) {
				direct |= (c.late_tap[l] > c.late_feed_tap + kPfDepth ? 0U : 1U << 3) | (c.late_ap_off[l] > kPfDepth ? 0U : 1U << 4) |
					(c.late_off[l] > kPfDepth ? 0U : 1U << 5);
			}
		}
		can_pf = ok;
		direct_groups = direct;
		pan_static = true;
		pan_ramp_all = true;
		const int counter = block_frames - base;
		const float delta = 1.0F / static_cast<float>(counter);
		ramp_mask[0] = ramp_mask[1] = 0;
		active_mask[0] = active_mask[1] = 0;
		OALSFX_UNROLL
		for (int l = kLine0; l < kLine1; ++l) {
			const float* target = (l < 4 ? c.pan_early[l] : c.pan_late[l - 4]);
			OALSFX_UNROLL
			for (int k = 0; k < (CT ? CT : kMaxChannels); ++k) {
				if (CT || k < channels) {
					const float gain = cur_gain[l][k];
					const float step = (target[k] - gain) * delta;
					const uint32_t bit = 1U << ((l & 3) * 8 + k);
					if (fabsf(step) > FLT_EPSILON) {
						ramp_mask[l >> 2] |= bit;
						step_gain[l][k] = step;
						pan_static = false;
					} else {
						step_gain[l][k] = 0.0F;
						pan_ramp_all = false;
						if (audible(gain)) {
							active_mask[l >> 2] |= bit;
						} else if (gain != 0.0F) {
							// (an inaudible gain that is an exact zero -- the LFE column of every 5.1 / 6.1 / 7.1 pan -- keeps
							// the static path: its product is +-0, and adding that changes nothing: the bus starts at +0 and
							// never becomes -0; finite samples assumed, as for the host's sanitized gains, fx.cuh pan_add)
							pan_static = false;
						}
					}
				}
			}
		}
	}

	// Sub-chunk epilogue: fade bookkeeping (oalsfxpp.cpp:6118-6138) and ramp snap (oalsfxpp.cpp:2778-2783).
	template <int CT>
	OALSFX_HD void end_sub(const ReverbCoef& c, int channels)
	{
		// (eax_)verb_pass tail: fade = min(1, fade + todo * fade_step); per-sample increments of the
		// exactly representable 1/128 give the same value.
		if (faded) {
			fade = fminf(1.0F, fade);
		}
		if (fade_count < kFadeSamples) {
			fade_count += sub_todo;
			if (fade_count >= kFadeSamples) {
				fade_count = kFadeSamples;
				fade = 1.0F;
				for (int i = 0; i < 4; ++i) { // commit the new tap sets (each half its own groups)
					if (EARLY) {
						st_mem[(kWOld + 0 + i) * kLanes] = static_cast<uint32_t>(c.early_tap[i]);
						st_mem[(kWOld + 4 + i) * kLanes] = static_cast<uint32_t>(c.early_ap_off[i]);
						st_mem[(kWOld + 8 + i) * kLanes] = static_cast<uint32_t>(c.early_off[i]);
					}
					if (LATE) {
						st_mem[(kWOld + 12 + i) * kLanes] = static_cast<uint32_t>(c.late_tap[i]);
						st_mem[(kWOld + 16 + i) * kLanes] = static_cast<uint32_t>(c.late_ap_off[i]);
						st_mem[(kWOld + 20 + i) * kLanes] = static_cast<uint32_t>(c.late_off[i]);
					}
				}
			}
		}
		const bool ramp_done = (sub_todo == block_frames - base); // `pos == counter`
		if (ramp_done) {
			OALSFX_UNROLL
			for (int l = kLine0; l < kLine1; ++l) {
				const float* target = (l < 4 ? c.pan_early[l] : c.pan_late[l - 4]);
				OALSFX_UNROLL
				for (int k = 0; k < (CT ? CT : kMaxChannels); ++k) {
					if ((CT || k < channels) && (ramp_mask[l >> 2] >> ((l & 3) * 8 + k)) & 1U) {
						cur_gain[l][k] = target[k];
					}
				}
			}
		}
		base += sub_todo;
	}

	// Delay read (oalsfxpp.cpp:7358-7406): prefetched value, direct read, or old/new cross-fade.
	template <bool PF, bool MIXED = false>
	OALSFX_HD float tap(int tap_index, int ring_word0, int mask, int pos, int group, int line, int new_d, float mu, uint32_t dg = 0) const
	{
		if (PF && (!MIXED || ((dg >> group) & 1U) == 0)) {
			return pf_cur[(tap_index - kTap0) * kLanes];
		}
		if (PF || !faded) {
			return ring.ld(ring_word0 + ((pos - new_d) & mask)); // committed: old == new
		}
		const float a = ring.ld(ring_word0 + ((pos - old_tap(group, line)) & mask));
		const float b = ring.ld(ring_word0 + ((pos - new_d) & mask));
		return a + ((b - a) * mu);
	}

	// vector_partial_scatter (oalsfxpp.cpp:7510-7521) on lines held as pairs a = (f0, f1), b = (f2, f3):
	//   v0 = x*f0 + y*(f1 + -f2 + f3)    v1 = x*f1 + y*(-f0 + f2 + f3)
	//   v2 = x*f2 + y*(f0 + -f1 + f3)    v3 = x*f3 + y*(-f0 + -f1 + -f2)
	// The four inner sums mix the halves, so they stay scalar; the products and the outer sums are packed.
	OALSFX_HD static void scatter2(F2& a, F2& b, float x, float y)
	{
		const float f0 = f2_lo(a), f1 = f2_hi(a), f2v = f2_lo(b), f3 = f2_hi(b);
		const float i0 = (f1 + -f2v) + f3, i1 = (-f0 + f2v) + f3, i2 = (f0 + -f1) + f3, i3 = (-f0 + -f1) + -f2v;
		a = (a * x) + (f2(i0, i1) * y);
		b = (b * x) + (f2(i2, i3) * y);
	}

	// The same scatter applied to the REVERSED vector r = (f3, f2, f1, f0) (vector_reverse,
	// oalsfxpp.cpp:7652, 7781), returned reversed as well -- a = (v3, v2), b = (v1, v0) -- so that no half
	// ever has to change sides:  v0 = x*f3 + y*(f2 + -f1 + f0)   v1 = x*f2 + y*(-f3 + f1 + f0)
	//                            v2 = x*f1 + y*(f3 + -f2 + f0)   v3 = x*f0 + y*(-f3 + -f2 + -f1)
	OALSFX_HD static void scatter2_reversed(F2& a, F2& b, float x, float y)
	{
		const float f0 = f2_lo(a), f1 = f2_hi(a), f2v = f2_lo(b), f3 = f2_hi(b);
		const float i0 = (f2v + -f1) + f0, i1 = (-f3 + f1) + f0, i2 = (f3 + -f2v) + f0, i3 = (-f3 + -f2v) + -f1;
		a = (a * x) + (f2(i3, i2) * y);
		b = (b * x) + (f2(i1, i0) * y);
	}

	// vector_allpass_x (oalsfxpp.cpp:7533-7562) on line pairs.
	template <bool PF, bool MIXED = false>
	OALSFX_HD void vector_allpass2(const ReverbCoef& c, F2& va, F2& vb, int ring_idx, int tap_base, int group,
		const int32_t* new_off, int pos, float mu, uint32_t dg = 0) const
	{
		const int len = c.mask[ring_idx] + 1;
		const int word0 = c.ring_base[ring_idx];
		const int mask = c.mask[ring_idx];
		const F2 ta = f2(tap<PF, MIXED>(tap_base + 0, word0 + 0 * len, mask, pos, group, 0, new_off[0], mu, dg),
			tap<PF, MIXED>(tap_base + 1, word0 + 1 * len, mask, pos, group, 1, new_off[1], mu, dg));
		const F2 tb = f2(tap<PF, MIXED>(tap_base + 2, word0 + 2 * len, mask, pos, group, 2, new_off[2], mu, dg),
			tap<PF, MIXED>(tap_base + 3, word0 + 3 * len, mask, pos, group, 3, new_off[3], mu, dg));
		const F2 ina = va, inb = vb;
		va = ta - (ina * c.ap_feed_coeff);
		vb = tb - (inb * c.ap_feed_coeff);
		F2 fa = ina + (va * c.ap_feed_coeff);
		F2 fb = inb + (vb * c.ap_feed_coeff);
		scatter2(fa, fb, c.mix_x, c.mix_y);
		ring.st(word0 + 0 * len + (pos & mask), f2_lo(fa));
		ring.st(word0 + 1 * len + (pos & mask), f2_hi(fa));
		ring.st(word0 + 2 * len + (pos & mask), f2_lo(fb));
		ring.st(word0 + 3 * len + (pos & mask), f2_hi(fb));
	}

#if defined(__CUDA_ARCH__)
	// Request ring positions p4 .. p4+3 (p4 a multiple of kPfBatch) of all 24 taps: lane = s * 8 + q
	// copies streams 4q .. 4q+3 of position p4 + s, i.e. the warp moves one contiguous 512-byte run
	// per tap (the rings are line-major: consecutive positions of a line are consecutive 128-byte rows).
	// `row0`: window row of position p4; rows wrap modulo kPfSlots.
	__device__ __forceinline__ void issue_batch(const ReverbCoef& c, int p4, int row0) const
	{
		const int lane = threadIdx.x % kLanes;
		const int q4 = (lane & 7) * 4;
		const int ps = p4 + (lane >> 3);
		int row = row0 + (lane >> 3);
		row = (row >= kPfSlots ? row - kPfSlots : row);
		const unsigned dst = pf_s + static_cast<unsigned>((row * (kTaps * kLanes) + q4) * 4);
		const float* src = ring.p - lane + q4;
		const int len0 = c.mask[0] + 1, len1 = c.mask[1] + 1, len2 = c.mask[2] + 1, len3 = c.mask[3] + 1, len4 = c.mask[4] + 1;
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			if (EARLY) {
				cp_async_16_s(dst + (0 - kTap0 + l) * kLanes * 4, src + static_cast<unsigned>(c.ring_base[0] + l * len0 + ((ps - c.early_tap[l]) & c.mask[0])) * kLanes);
				cp_async_16_s(dst + (4 - kTap0 + l) * kLanes * 4, src + static_cast<unsigned>(c.ring_base[1] + l * len1 + ((ps - c.early_ap_off[l]) & c.mask[1])) * kLanes);
				cp_async_16_s(dst + (8 - kTap0 + l) * kLanes * 4, src + static_cast<unsigned>(c.ring_base[2] + l * len2 + ((ps - c.early_off[l]) & c.mask[2])) * kLanes);
			}
			if (LATE) {
				cp_async_16_s(dst + (12 - kTap0 + l) * kLanes * 4, src + static_cast<unsigned>(c.ring_base[0] + l * len0 + ((ps - c.late_tap[l]) & c.mask[0])) * kLanes);
				cp_async_16_s(dst + (16 - kTap0 + l) * kLanes * 4, src + static_cast<unsigned>(c.ring_base[4] + l * len4 + ((ps - c.late_off[l]) & c.mask[4])) * kLanes);
				cp_async_16_s(dst + (20 - kTap0 + l) * kLanes * 4, src + static_cast<unsigned>(c.ring_base[3] + l * len3 + ((ps - c.late_ap_off[l]) & c.mask[3])) * kLanes);
			}
		}
		cp_async_commit_group();
	}
#endif

	template <int CT, bool FAST = false>
	OALSFX_HD void step(const SlotCoef& sc, const float* wet, float* acc, int channels)
	{
		const ReverbCoef& c = sc.u.reverb;
		if (OALSFX_UNLIKELY(sub_left == 0)) {
			begin_sub<CT>(c, channels);
		}
		const int pos = offset;
#if defined(__CUDA_ARCH__)
		// The batched copies are a whole-warp affair (a lane fetches other lanes' streams), so the
		// decision is a vote: every lane of the tile must be in the prefetchable state.
		const bool window_all = pf_col != nullptr && __all_sync(0xFFFFFFFFU, can_pf && direct_groups == 0);
		// (second vote only off the steady state) some lane needs a tap group read in place: the union of the groups
		const bool window_some = !window_all && pf_col != nullptr && __all_sync(0xFFFFFFFFU, can_pf);
		if (OALSFX_LIKELY(window_all || window_some)) {
			// Invariant while primed, at position r of the batch whose first position sits in window row
			// pf_row4: that batch has landed, and the next one has been requested iff r >= kPfIssueAt.
			const int r = pos & (kPfBatch - 1);
			if (!primed) {
				__syncwarp(); // every lane is done with the window rows about to be overwritten
				pf_row4 = 0;
				issue_batch(c, pos - r, 0);
				if (kPfSlots == 8) {
					issue_batch(c, pos - r + kPfBatch, kPfBatch);
					cp_async_wait_group<1>();
					__syncwarp();
				} else {
					cp_async_wait_group<0>();
					__syncwarp();
					if (r >= kPfIssueAt) { // rows 0, 1 are re-used by the next batch: only once this one has landed
						issue_batch(c, pos - r + kPfBatch, kPfBatch);
					}
				}
				primed = true;
			} else {
				if (r == 0) {
					pf_row4 += kPfBatch;
					pf_row4 = (pf_row4 >= kPfSlots ? pf_row4 - kPfSlots : pf_row4);
				}
				if (r == kPfIssueAt) {
					int next = pf_row4 + kPfBatch;
					next = (next >= kPfSlots ? next - kPfSlots : next);
					__syncwarp(); // the rows about to be overwritten have been consumed by every lane
					issue_batch(c, pos - r + kPfBatch, next);
				}
				if (r == 0) {
					cp_async_wait_group<(kPfIssueAt == 0 ? 1 : 0)>(); // the batch holding `pos` has landed
					__syncwarp();                                      // ... for every lane's share of it
				}
			}
			int row = pf_row4 + r;
			row = (row >= kPfSlots ? row - kPfSlots : row);
			pf_cur = pf_col + row * (kTaps * kLanes);
			if (OALSFX_LIKELY(window_all)) {
				body<CT, true>(c, wet, acc, channels, pos); // branch-free: every read is a shared-memory load
			} else {
				body<CT, true, true>(c, wet, acc, channels, pos, __reduce_or_sync(0xFFFFFFFFU, direct_groups));
			}
		} else {
			if (primed) {
				// leaving the batched mode: nothing of the window may still be in flight when it is primed
				// again (an old copy landing after a new one would put stale rows into the window)
				cp_async_wait_group<0>();
				primed = false;
			}
			body<CT, false>(c, wet, acc, channels, pos);
		}
#else
		body<CT, false>(c, wet, acc, channels, pos);
#endif
		sub_left -= 1;
		if (OALSFX_UNLIKELY(sub_left == 0)) {
			end_sub<CT>(c, channels);
		}
	}

	// One sample.  PF: all 24 ring reads come from the prefetch window (steady state).
	template <int CT, bool PF, bool MIXED = false>
	OALSFX_HD void body(const ReverbCoef& c, const float* wet, float* acc, int channels, const int pos, const uint32_t dg = 0)
	{
		const float mu = fade;

		const int main_len = c.mask[0] + 1, main0 = c.ring_base[0], main_mask = c.mask[0];
		const int eline_len = c.mask[2] + 1, eline0 = c.ring_base[2], eline_mask = c.mask[2];
		const int lline_len = c.mask[4] + 1, lline0 = c.ring_base[4], lline_mask = c.mask[4];

		if (INPUT) {
			reverb_input_stage(c, wet, lp, hp, ring, pos);
		}

		// The four lines are carried as two pairs: a = lines (0, 1), b = lines (2, 3).
		F2 fa, fb;
		float early_out[4] = {0.0F, 0.0F, 0.0F, 0.0F}, late_out[4] = {0.0F, 0.0F, 0.0F, 0.0F};

		if (EARLY) {
		// ---- early reflections (oalsfxpp.cpp:7625-7672) ----
		fa = f2(tap<PF, MIXED>(0, main0 + 0 * main_len, main_mask, pos, 0, 0, c.early_tap[0], mu, dg),
			tap<PF, MIXED>(1, main0 + 1 * main_len, main_mask, pos, 0, 1, c.early_tap[1], mu, dg)) * f2(c.early_tap_coeff[0], c.early_tap_coeff[1]);
		fb = f2(tap<PF, MIXED>(2, main0 + 2 * main_len, main_mask, pos, 0, 2, c.early_tap[2], mu, dg),
			tap<PF, MIXED>(3, main0 + 3 * main_len, main_mask, pos, 0, 3, c.early_tap[3], mu, dg)) * f2(c.early_tap_coeff[2], c.early_tap_coeff[3]);
		vector_allpass2<PF, MIXED>(c, fa, fb, 1, 4, 1, c.early_ap_off, pos, mu, dg);
		// delay_line_in4_rev: line j receives f[3 - j]
		ring.st(eline0 + 0 * eline_len + (pos & eline_mask), f2_hi(fb));
		ring.st(eline0 + 1 * eline_len + (pos & eline_mask), f2_lo(fb));
		ring.st(eline0 + 2 * eline_len + (pos & eline_mask), f2_hi(fa));
		ring.st(eline0 + 3 * eline_len + (pos & eline_mask), f2_lo(fa));
		fa = fa + (f2(tap<PF, MIXED>(8, eline0 + 0 * eline_len, eline_mask, pos, 2, 0, c.early_off[0], mu, dg),
			tap<PF, MIXED>(9, eline0 + 1 * eline_len, eline_mask, pos, 2, 1, c.early_off[1], mu, dg)) * f2(c.early_coeff[0], c.early_coeff[1]));
		fb = fb + (f2(tap<PF, MIXED>(10, eline0 + 2 * eline_len, eline_mask, pos, 2, 2, c.early_off[2], mu, dg),
			tap<PF, MIXED>(11, eline0 + 3 * eline_len, eline_mask, pos, 2, 3, c.early_off[3], mu, dg)) * f2(c.early_coeff[2], c.early_coeff[3]));
		early_out[0] = f2_lo(fa);
		early_out[1] = f2_hi(fa);
		early_out[2] = f2_lo(fb);
		early_out[3] = f2_hi(fb);
		{
			// vector_reverse + scatter, fed to the late reverb through the main line
			F2 ra = fa, rb = fb;
			scatter2_reversed(ra, rb, c.mix_x, c.mix_y); // ra = (v3, v2), rb = (v1, v0)
			const int feed = (pos - c.late_feed_tap) & main_mask;
			ring.st(main0 + 0 * main_len + feed, f2_hi(rb));
			ring.st(main0 + 1 * main_len + feed, f2_lo(rb));
			ring.st(main0 + 2 * main_len + feed, f2_hi(ra));
			ring.st(main0 + 3 * main_len + feed, f2_lo(ra));
		}
		} // EARLY

		if (LATE) {
		// ---- late reverb (oalsfxpp.cpp:7735-7794) ----
		// calc_modulation_delays (oalsfxpp.cpp:7443-7470); when depth and filter are both zero the
		// product range*sinus is +-0 and the delay is 0 whatever the sinus is.
		int mod_delay = 0;
		if (PF && !MIXED) { // every group served by the window implies depth == 0 and filter == 0: the delay is 0, only the index moves
			mod_index += 1;
			if (mod_index >= mod_range) {
				mod_index = 0;
			}
		} else {
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
		fa = f2(tap<PF, MIXED>(12, main0 + 0 * main_len, main_mask, pos, 3, 0, c.late_tap[0], mu, dg),
			tap<PF, MIXED>(13, main0 + 1 * main_len, main_mask, pos, 3, 1, c.late_tap[1], mu, dg)) * c.density_gain;
		fb = f2(tap<PF, MIXED>(14, main0 + 2 * main_len, main_mask, pos, 3, 2, c.late_tap[2], mu, dg),
			tap<PF, MIXED>(15, main0 + 3 * main_len, main_mask, pos, 3, 3, c.late_tap[3], mu, dg)) * c.density_gain;
		const int mod_pos = pos - mod_delay;
		fa = fa + f2(tap<PF, MIXED>(16, lline0 + 0 * lline_len, lline_mask, mod_pos, 5, 0, c.late_off[0], mu, dg),
			tap<PF, MIXED>(17, lline0 + 1 * lline_len, lline_mask, mod_pos, 5, 1, c.late_off[1], mu, dg));
		fb = fb + f2(tap<PF, MIXED>(18, lline0 + 2 * lline_len, lline_mask, mod_pos, 5, 2, c.late_off[2], mu, dg),
			tap<PF, MIXED>(19, lline0 + 3 * lline_len, lline_mask, mod_pos, 5, 3, c.late_off[3], mu, dg));
		OALSFX_UNROLL
		for (int h = 0; h < 2; ++h) {
			// late_t60_filter: two first-order sections and the mid gain (oalsfxpp.cpp:7691-7719), lines 2h, 2h+1
			const int j = 2 * h;
			const F2 in = (h == 0 ? fa : fb);
			const F2 o1 = (f2(c.t60_lf[j][0], c.t60_lf[j + 1][0]) * in) + (f2(c.t60_lf[j][1], c.t60_lf[j + 1][1]) * t60p[h][0][0]) +
				(f2(c.t60_lf[j][2], c.t60_lf[j + 1][2]) * t60p[h][0][1]);
			t60p[h][0][0] = in;
			t60p[h][0][1] = o1;
			const F2 o2 = (f2(c.t60_hf[j][0], c.t60_hf[j + 1][0]) * o1) + (f2(c.t60_hf[j][1], c.t60_hf[j + 1][1]) * t60p[h][1][0]) +
				(f2(c.t60_hf[j][2], c.t60_hf[j + 1][2]) * t60p[h][1][1]);
			t60p[h][1][0] = o1;
			t60p[h][1][1] = o2;
			const F2 out = f2(c.t60_mid[j], c.t60_mid[j + 1]) * o2;
			if (h == 0) {
				fa = out;
			} else {
				fb = out;
			}
		}
		vector_allpass2<PF, MIXED>(c, fa, fb, 3, 20, 4, c.late_ap_off, pos, mu, dg);
		late_out[0] = f2_lo(fa);
		late_out[1] = f2_hi(fa);
		late_out[2] = f2_lo(fb);
		late_out[3] = f2_hi(fb);
		{
			F2 ra = fa, rb = fb;
			scatter2_reversed(ra, rb, c.mix_x, c.mix_y);
			ring.st(lline0 + 0 * lline_len + (pos & lline_mask), f2_hi(rb));
			ring.st(lline0 + 1 * lline_len + (pos & lline_mask), f2_lo(rb));
			ring.st(lline0 + 2 * lline_len + (pos & lline_mask), f2_hi(ra));
			ring.st(lline0 + 3 * lline_len + (pos & lline_mask), f2_lo(ra));
		}
		} // LATE

		offset += 1;
		if (faded) {
			fade += 1.0F / kFadeSamples; // fade_step (exactly representable, so the running sum is exact)
		}

		// ---- pan the 8 line outputs to the bus with stepped gains (oalsfxpp.cpp:6142-6166, 2752-2798) ----
		if (OALSFX_LIKELY(pan_static) && CT == 2) {
			// acc[k] += d_l * gain[l][k], both output channels at once, lines in the reference's order
			F2 accp = f2(acc[0], acc[1]);
			OALSFX_UNROLL
			for (int l = kLine0; l < kLine1; ++l) {
				const float d = (l < 4 ? early_out[l] : late_out[l - 4]);
				accp = accp + (f2_bcast(d) * f2(cur_gain[l][0], cur_gain[l][1]));
			}
			acc[0] = f2_lo(accp);
			acc[1] = f2_hi(accp);
			return;
		}
		if (pan_static) {
			OALSFX_UNROLL
			for (int l = kLine0; l < kLine1; ++l) {
				const float d = (l < 4 ? early_out[l] : late_out[l - 4]);
				OALSFX_UNROLL
				for (int k = 0; k < (CT ? CT : kMaxChannels); ++k) {
					if (CT || k < channels) {
						acc[k] += d * cur_gain[l][k];
					}
				}
			}
			return;
		}
		if (pan_ramp_all && CT != 0) {
			// every gain ramps (MixHelpers::mix, oalsfxpp.cpp:2769-2777): the general loop below without its tests
			OALSFX_UNROLL
			for (int l = kLine0; l < kLine1; ++l) {
				const float d = (l < 4 ? early_out[l] : late_out[l - 4]);
				OALSFX_UNROLL
				for (int k = 0; k < CT; ++k) {
					acc[k] += d * cur_gain[l][k];
					cur_gain[l][k] += step_gain[l][k];
				}
			}
			return;
		}
		OALSFX_UNROLL
		for (int l = kLine0; l < kLine1; ++l) {
			const float d = (l < 4 ? early_out[l] : late_out[l - 4]);
			OALSFX_UNROLL
			for (int k = 0; k < (CT ? CT : kMaxChannels); ++k) {
				if (CT || k < channels) {
					const uint32_t bit = 1U << ((l & 3) * 8 + k);
					if (ramp_mask[l >> 2] & bit) {
						acc[k] += d * cur_gain[l][k];
						cur_gain[l][k] += step_gain[l][k];
					} else if (active_mask[l >> 2] & bit) {
						acc[k] += d * cur_gain[l][k];
					}
				}
			}
		}

	}

	template <int CT>
	OALSFX_HD void end_ct(const SlotCoef&, uint32_t* st, int channels)
	{
		OALSFX_UNROLL
		for (int l = 0; l < 4; ++l) {
			if (INPUT) {
				store_words(lp[l], st + (kWLp + l * 4) * kLanes);
				store_words(hp[l], st + (kWHp + l * 4) * kLanes);
			}
		}
		OALSFX_UNROLL
		for (int h = 0; h < 2; ++h) {
			OALSFX_UNROLL
			for (int w = 0; w < 4; ++w) {
				if (LATE) {
					st[(kWT60 + (2 * h) * 4 + w) * kLanes] = float_as_word(f2_lo(t60p[h][w >> 1][w & 1]));
					st[(kWT60 + (2 * h + 1) * 4 + w) * kLanes] = float_as_word(f2_hi(t60p[h][w >> 1][w & 1]));
				}
			}
		}
		OALSFX_UNROLL
		for (int l = kLine0; l < kLine1; ++l) {
			OALSFX_UNROLL
			for (int k = 0; k < (CT ? CT : kMaxChannels); ++k) {
				if (CT || k < channels) {
					st[(kWGain + l * kMaxChannels + k) * kLanes] = float_as_word(cur_gain[l][k]);
				}
			}
		}
		if (LATE) { // the late half finishes last and owns the scalars
			st[(kWScalars + 0) * kLanes] = static_cast<uint32_t>(offset);
			st[(kWScalars + 1) * kLanes] = static_cast<uint32_t>(fade_count);
			st[(kWScalars + 2) * kLanes] = static_cast<uint32_t>(mod_index);
			st[(kWScalars + 3) * kLanes] = static_cast<uint32_t>(mod_range);
			st[(kWScalars + 4) * kLanes] = float_as_word(mod_filter);
		}
#if defined(__CUDA_ARCH__)
		cp_async_wait_group<0>(); // nothing of this warp's window may still be in flight when the CTA exits
#endif
	}
};

using FxReverb = FxReverbT<true>;
using FxReverbTail = FxReverbT<false>;
using FxReverbEarly = FxReverbT<false, true, false>;
using FxReverbLate = FxReverbT<false, false, true>;

// The reverb's input stage alone, for a warp that runs ahead of the FxReverbTail of the same slot.
// Shares the slot's state region: owns the lp/hp history words, reads (never writes) the offset.
struct FxReverbInput {
	static constexpr bool kIsNull = false;
	BiquadHist lp[4], hp[4];
	int32_t offset;
	LaneMem ring;

	OALSFX_HD void set_prefetch(float*) {}
	OALSFX_HD void prefetch_issue(const SlotCoef&, int) {}
	OALSFX_HD void prefetch_next(const SlotCoef&) {}

	template <int CT>
	OALSFX_HD void begin(const SlotCoef&, uint32_t* st, float* ring_p, bool, int, int)
	{
		ring.p = ring_p;
		OALSFX_UNROLL
		for (int l = 0; l < 4; ++l) {
			load_words(lp[l], st + (FxReverb::kWLp + l * 4) * kLanes);
			load_words(hp[l], st + (FxReverb::kWHp + l * 4) * kLanes);
		}
		offset = static_cast<int32_t>(st[(FxReverb::kWScalars + 0) * kLanes]);
	}

	template <int CT, bool FAST = false>
	OALSFX_HD void step(const SlotCoef& sc, const float* wet, float*, int)
	{
		reverb_input_stage(sc.u.reverb, wet, lp, hp, ring, offset);
		offset += 1;
	}

	template <int CT>
	OALSFX_HD void end_ct(const SlotCoef&, uint32_t* st, int)
	{
		OALSFX_UNROLL
		for (int l = 0; l < 4; ++l) {
			store_words(lp[l], st + (FxReverb::kWLp + l * 4) * kLanes);
			store_words(hp[l], st + (FxReverb::kWHp + l * 4) * kLanes);
		}
	}
};

} // namespace oalsfx

#endif
