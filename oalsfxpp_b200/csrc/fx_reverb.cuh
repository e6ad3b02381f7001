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
// early line, late, late line, late all-pass} x 4 lines.  Two batches of kPfBatch positions: the one
// being consumed and the one in flight.
constexpr int kPfBatch = 4;                 // ring positions per batched copy (32 lanes x 16 B = 4 lines)
constexpr int kPfSlots = 2 * kPfBatch;      // power of two
constexpr int kPfDepth = kPfSlots - 1;      // furthest position requested beyond the current one
constexpr int kPfTaps = 24;
constexpr int kPfWarpFloats = kPfSlots * kPfTaps * kLanes; // 24 KiB per reverb warp

// Input stage of the reverb, shared by the whole effect (FxReverbT<true>) and by FxReverbInput,
// which lets another warp run it ahead of the rest (duo.cuh): B-format -> A-format (mix_row with
// the b2a matrix, oalsfxpp.cpp:6099-6113, 6377-6383), the master shelf filter(s), and the feed of the
// main delay line (oalsfxpp.cpp:7821-7832 / 7867-7879).
OALSFX_HD void reverb_input_stage(const ReverbCoef& c, const float* wet, BiquadHist* lp, BiquadHist* hp,
	const LaneMem& ring, int pos)
{
	const int main_len = c.mask[0] + 1, main0 = c.ring_base[0], main_mask = c.mask[0];
	// Every b2a entry is +-q, and w * (-q) == -(w * q) bit for bit, so the four products are formed
	// once and added with the row's signs, in the reference's order k = 0..3 starting from 0.
	constexpr float q = 0.288675134595F;
	const float p[4] = {wet[0] * q, wet[1] * q, wet[2] * q, wet[3] * q};
	const bool neg[4][4] = {{false, false, false, false}, {false, true, true, false}, {false, false, true, true},
		{false, true, false, true}};
	OALSFX_UNROLL
	for (int l = 0; l < 4; ++l) {
		float a = 0.0F;
		OALSFX_UNROLL
		for (int k = 0; k < 4; ++k) {
			a = neg[l][k] ? a - p[k] : a + p[k];
		}
		float v = biquad_step(c.lp, lp[l], a);
		if (c.is_eax) {
			v = biquad_step(c.hp, hp[l], v);
		}
		ring.st(main0 + l * main_len + (pos & main_mask), v);
	}
}

// INPUT = false: the input stage (and the lp/hp filter history words of the state) belong to a
// FxReverbInput running in another warp; `wet` is not read.
template <bool INPUT>
struct FxReverbT {
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
	float t60[4][2][2];
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
	bool pan_static;   // this sub-chunk: no gain ramps and every one of the 8 x C pan gains is audible

	OALSFX_HD void set_prefetch(float* column) { pf_col = column; }
	OALSFX_HD void prefetch_issue(const SlotCoef&, int) {} // the reverb runs its own pipeline inside step()

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
			t60[l][0][0] = word_as_float(st[(kWT60 + l * 4 + 0) * kLanes]);
			t60[l][0][1] = word_as_float(st[(kWT60 + l * 4 + 1) * kLanes]);
			t60[l][1][0] = word_as_float(st[(kWT60 + l * 4 + 2) * kLanes]);
			t60[l][1][1] = word_as_float(st[(kWT60 + l * 4 + 3) * kLanes]);
		}
		OALSFX_UNROLL
		for (int l = 0; l < 8; ++l) {
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
			mod_index = static_cast<int32_t>(mod_index * static_cast<int64_t>(c.mod_range) / mod_range);
			mod_range = c.mod_range;
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
		bool ok = pf_col != nullptr && !faded && c.mod_depth == 0.0F && mod_filter == 0.0F;
		OALSFX_UNROLL
		for (int l = 0; l < 4; ++l) {
			ok = ok && c.early_tap[l] > kPfDepth && c.early_ap_off[l] > kPfDepth && c.early_off[l] > kPfDepth &&
				c.late_tap[l] > c.late_feed_tap + kPfDepth && c.late_ap_off[l] > kPfDepth && c.late_off[l] > kPfDepth;
		}
		can_pf = ok;
		pan_static = true;
		const int counter = block_frames - base;
		const float delta = 1.0F / static_cast<float>(counter);
		ramp_mask[0] = ramp_mask[1] = 0;
		active_mask[0] = active_mask[1] = 0;
		OALSFX_UNROLL
		for (int l = 0; l < 8; ++l) {
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
						if (audible(gain)) {
							active_mask[l >> 2] |= bit;
						} else {
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
				for (int i = 0; i < 4; ++i) { // commit the new tap sets
					st_mem[(kWOld + 0 + i) * kLanes] = static_cast<uint32_t>(c.early_tap[i]);
					st_mem[(kWOld + 4 + i) * kLanes] = static_cast<uint32_t>(c.early_ap_off[i]);
					st_mem[(kWOld + 8 + i) * kLanes] = static_cast<uint32_t>(c.early_off[i]);
					st_mem[(kWOld + 12 + i) * kLanes] = static_cast<uint32_t>(c.late_tap[i]);
					st_mem[(kWOld + 16 + i) * kLanes] = static_cast<uint32_t>(c.late_ap_off[i]);
					st_mem[(kWOld + 20 + i) * kLanes] = static_cast<uint32_t>(c.late_off[i]);
				}
			}
		}
		const bool ramp_done = (sub_todo == block_frames - base); // `pos == counter`
		if (ramp_done) {
			OALSFX_UNROLL
			for (int l = 0; l < 8; ++l) {
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
	template <bool PF>
	OALSFX_HD float tap(int tap_index, int ring_word0, int mask, int pos, int group, int line, int new_d, float mu) const
	{
		if (PF) {
			return pf_cur[tap_index * kLanes];
		}
		if (!faded) {
			return ring.ld(ring_word0 + ((pos - new_d) & mask)); // committed: old == new
		}
		const float a = ring.ld(ring_word0 + ((pos - old_tap(group, line)) & mask));
		const float b = ring.ld(ring_word0 + ((pos - new_d) & mask));
		return a + ((b - a) * mu);
	}

	OALSFX_HD static void scatter(float* v, float x, float y)
	{
		const float f0 = v[0], f1 = v[1], f2 = v[2], f3 = v[3];
		v[0] = (x * f0) + (y * (f1 + -f2 + f3));
		v[1] = (x * f1) + (y * (-f0 + f2 + f3));
		v[2] = (x * f2) + (y * (f0 + -f1 + f3));
		v[3] = (x * f3) + (y * (-f0 + -f1 + -f2));
	}

	template <bool PF>
	OALSFX_HD void vector_allpass(const ReverbCoef& c, float* vec, int ring_idx, int tap_base, int group,
		const int32_t* new_off, int pos, float mu) const
	{
		const int len = c.mask[ring_idx] + 1;
		const int word0 = c.ring_base[ring_idx];
		float f[4];
		OALSFX_UNROLL
		for (int i = 0; i < 4; ++i) {
			const float input = vec[i];
			vec[i] = tap<PF>(tap_base + i, word0 + i * len, c.mask[ring_idx], pos, group, i, new_off[i], mu) - (c.ap_feed_coeff * input);
			f[i] = input + (c.ap_feed_coeff * vec[i]);
		}
		scatter(f, c.mix_x, c.mix_y);
		OALSFX_UNROLL
		for (int i = 0; i < 4; ++i) {
			ring.st(word0 + i * len + (pos & c.mask[ring_idx]), f[i]);
		}
	}

#if defined(__CUDA_ARCH__)
	// Request ring positions p4 .. p4+3 (p4 a multiple of kPfBatch) of all 24 taps: lane = s * 8 + q
	// copies streams 4q .. 4q+3 of position p4 + s, i.e. the warp moves one contiguous 512-byte run
	// per tap (the rings are line-major: consecutive positions of a line are consecutive 128-byte rows).
	__device__ __forceinline__ void issue_batch(const ReverbCoef& c, int p4) const
	{
		const int lane = threadIdx.x % kLanes;
		const int q4 = (lane & 7) * 4;
		const int ps = p4 + (lane >> 3);
		float* dst = pf_col - lane + (ps & (kPfSlots - 1)) * (kPfTaps * kLanes) + q4;
		const float* src = ring.p - lane + q4;
		const int len0 = c.mask[0] + 1, len1 = c.mask[1] + 1, len2 = c.mask[2] + 1, len3 = c.mask[3] + 1, len4 = c.mask[4] + 1;
#pragma unroll
		for (int l = 0; l < 4; ++l) {
			cp_async_16(dst + (0 + l) * kLanes, src + static_cast<unsigned>(c.ring_base[0] + l * len0 + ((ps - c.early_tap[l]) & c.mask[0])) * kLanes);
			cp_async_16(dst + (4 + l) * kLanes, src + static_cast<unsigned>(c.ring_base[1] + l * len1 + ((ps - c.early_ap_off[l]) & c.mask[1])) * kLanes);
			cp_async_16(dst + (8 + l) * kLanes, src + static_cast<unsigned>(c.ring_base[2] + l * len2 + ((ps - c.early_off[l]) & c.mask[2])) * kLanes);
			cp_async_16(dst + (12 + l) * kLanes, src + static_cast<unsigned>(c.ring_base[0] + l * len0 + ((ps - c.late_tap[l]) & c.mask[0])) * kLanes);
			cp_async_16(dst + (16 + l) * kLanes, src + static_cast<unsigned>(c.ring_base[4] + l * len4 + ((ps - c.late_off[l]) & c.mask[4])) * kLanes);
			cp_async_16(dst + (20 + l) * kLanes, src + static_cast<unsigned>(c.ring_base[3] + l * len3 + ((ps - c.late_ap_off[l]) & c.mask[3])) * kLanes);
		}
		cp_async_commit_group();
	}
#endif

	template <int CT, bool FAST = false>
	OALSFX_HD void step(const SlotCoef& sc, const float* wet, float* acc, int channels)
	{
		const ReverbCoef& c = sc.u.reverb;
		if (sub_left == 0) {
			begin_sub<CT>(c, channels);
		}
		const int pos = offset;
#if defined(__CUDA_ARCH__)
		// The batched copies are a whole-warp affair (a lane fetches other lanes' streams), so the
		// decision is a vote: every lane of the tile must be in the prefetchable state.
		if (pf_col != nullptr && __all_sync(0xFFFFFFFFU, can_pf)) {
			// Invariant while primed: the batch holding `pos` and the one after it have been requested.
			const bool batch_start = (pos & (kPfBatch - 1)) == 0;
			if (!primed || batch_start) {
				__syncwarp(); // every lane is done with the window rows about to be overwritten
				if (!primed) {
					issue_batch(c, pos & ~(kPfBatch - 1));
					issue_batch(c, (pos & ~(kPfBatch - 1)) + kPfBatch);
					primed = true;
				} else {
					issue_batch(c, pos + kPfBatch);
				}
				cp_async_wait_group<1>(); // all but the newest batch: the one holding `pos` has landed
				__syncwarp();             // ... for every lane's share of it
			}
			pf_cur = pf_col + (pos & (kPfSlots - 1)) * (kPfTaps * kLanes);
			body<CT, true>(c, wet, acc, channels, pos); // branch-free: every read is a shared-memory load
		} else {
			primed = false;
			body<CT, false>(c, wet, acc, channels, pos);
		}
#else
		body<CT, false>(c, wet, acc, channels, pos);
#endif
		sub_left -= 1;
		if (sub_left == 0) {
			end_sub<CT>(c, channels);
		}
	}

	// One sample.  PF: all 24 ring reads come from the prefetch window (steady state).
	template <int CT, bool PF>
	OALSFX_HD void body(const ReverbCoef& c, const float* wet, float* acc, int channels, const int pos)
	{
		const float mu = fade;

		const int main_len = c.mask[0] + 1, main0 = c.ring_base[0], main_mask = c.mask[0];
		const int eline_len = c.mask[2] + 1, eline0 = c.ring_base[2], eline_mask = c.mask[2];
		const int lline_len = c.mask[4] + 1, lline0 = c.ring_base[4], lline_mask = c.mask[4];

		if (INPUT) {
			reverb_input_stage(c, wet, lp, hp, ring, pos);
		}

		float f[4];
		float early_out[4], late_out[4];

		// ---- early reflections (oalsfxpp.cpp:7625-7672) ----
		OALSFX_UNROLL
		for (int j = 0; j < 4; ++j) {
			f[j] = tap<PF>(0 + j, main0 + j * main_len, main_mask, pos, 0, j, c.early_tap[j], mu) * c.early_tap_coeff[j];
		}
		vector_allpass<PF>(c, f, 1, 4, 1, c.early_ap_off, pos, mu);
		OALSFX_UNROLL
		for (int j = 0; j < 4; ++j) {
			ring.st(eline0 + j * eline_len + (pos & eline_mask), f[3 - j]); // delay_line_in4_rev
		}
		OALSFX_UNROLL
		for (int j = 0; j < 4; ++j) {
			f[j] += tap<PF>(8 + j, eline0 + j * eline_len, eline_mask, pos, 2, j, c.early_off[j], mu) * c.early_coeff[j];
			early_out[j] = f[j];
		}
		{
			float r[4] = {f[3], f[2], f[1], f[0]}; // vector_reverse
			scatter(r, c.mix_x, c.mix_y);
			OALSFX_UNROLL
			for (int j = 0; j < 4; ++j) {
				ring.st(main0 + j * main_len + ((pos - c.late_feed_tap) & main_mask), r[j]);
			}
		}

		// ---- late reverb (oalsfxpp.cpp:7735-7794) ----
		// calc_modulation_delays (oalsfxpp.cpp:7443-7470); when depth and filter are both zero the
		// product range*sinus is +-0 and the delay is 0 whatever the sinus is.
		int mod_delay = 0;
		if (PF) { // can_pf implies depth == 0 and filter == 0: the delay is 0, only the index moves
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
		OALSFX_UNROLL
		for (int j = 0; j < 4; ++j) {
			f[j] = tap<PF>(12 + j, main0 + j * main_len, main_mask, pos, 3, j, c.late_tap[j], mu) * c.density_gain;
		}
		const int mod_pos = pos - mod_delay;
		OALSFX_UNROLL
		for (int j = 0; j < 4; ++j) {
			f[j] += tap<PF>(16 + j, lline0 + j * lline_len, lline_mask, mod_pos, 5, j, c.late_off[j], mu);
		}
		OALSFX_UNROLL
		for (int j = 0; j < 4; ++j) {
			// late_t60_filter: two first-order sections and the mid gain (oalsfxpp.cpp:7691-7719)
			const float in = f[j];
			const float o1 = (c.t60_lf[j][0] * in) + (c.t60_lf[j][1] * t60[j][0][0]) + (c.t60_lf[j][2] * t60[j][0][1]);
			t60[j][0][0] = in;
			t60[j][0][1] = o1;
			const float o2 = (c.t60_hf[j][0] * o1) + (c.t60_hf[j][1] * t60[j][1][0]) + (c.t60_hf[j][2] * t60[j][1][1]);
			t60[j][1][0] = o1;
			t60[j][1][1] = o2;
			f[j] = c.t60_mid[j] * o2;
		}
		vector_allpass<PF>(c, f, 3, 20, 4, c.late_ap_off, pos, mu);
		OALSFX_UNROLL
		for (int j = 0; j < 4; ++j) {
			late_out[j] = f[j];
		}
		{
			float r[4] = {f[3], f[2], f[1], f[0]};
			scatter(r, c.mix_x, c.mix_y);
			OALSFX_UNROLL
			for (int j = 0; j < 4; ++j) {
				ring.st(lline0 + j * lline_len + (pos & lline_mask), r[j]);
			}
		}

		offset += 1;
		if (faded) {
			fade += 1.0F / kFadeSamples; // fade_step (exactly representable, so the running sum is exact)
		}

		// ---- pan the 8 line outputs to the bus with stepped gains (oalsfxpp.cpp:6142-6166, 2752-2798) ----
		if (pan_static) {
			OALSFX_UNROLL
			for (int l = 0; l < 8; ++l) {
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
		OALSFX_UNROLL
		for (int l = 0; l < 8; ++l) {
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
			st[(kWT60 + l * 4 + 0) * kLanes] = float_as_word(t60[l][0][0]);
			st[(kWT60 + l * 4 + 1) * kLanes] = float_as_word(t60[l][0][1]);
			st[(kWT60 + l * 4 + 2) * kLanes] = float_as_word(t60[l][1][0]);
			st[(kWT60 + l * 4 + 3) * kLanes] = float_as_word(t60[l][1][1]);
		}
		OALSFX_UNROLL
		for (int l = 0; l < 8; ++l) {
			OALSFX_UNROLL
			for (int k = 0; k < (CT ? CT : kMaxChannels); ++k) {
				if (CT || k < channels) {
					st[(kWGain + l * kMaxChannels + k) * kLanes] = float_as_word(cur_gain[l][k]);
				}
			}
		}
		st[(kWScalars + 0) * kLanes] = static_cast<uint32_t>(offset);
		st[(kWScalars + 1) * kLanes] = static_cast<uint32_t>(fade_count);
		st[(kWScalars + 2) * kLanes] = static_cast<uint32_t>(mod_index);
		st[(kWScalars + 3) * kLanes] = static_cast<uint32_t>(mod_range);
		st[(kWScalars + 4) * kLanes] = float_as_word(mod_filter);
#if defined(__CUDA_ARCH__)
		cp_async_wait_group<0>(); // nothing of this warp's window may still be in flight when the CTA exits
#endif
	}
};

using FxReverb = FxReverbT<true>;
using FxReverbTail = FxReverbT<false>;

// The reverb's input stage alone, for a warp that runs ahead of the FxReverbTail of the same slot.
// Shares the slot's state region: owns the lp/hp history words, reads (never writes) the offset.
struct FxReverbInput {
	static constexpr bool kIsNull = false;
	BiquadHist lp[4], hp[4];
	int32_t offset;
	LaneMem ring;

	OALSFX_HD void set_prefetch(float*) {}
	OALSFX_HD void prefetch_issue(const SlotCoef&, int) {}

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
