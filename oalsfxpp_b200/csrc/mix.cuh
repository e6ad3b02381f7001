// mix.cuh -- the fused per-block mixer: source encode + dry mix + up to four effect slots + output
// interleave in ONE pass over the block, intermediate buses held in registers (the data path of
// the reference's Api::Impl::mix_data / mix_source / write_f32, oalsfxpp.cpp:2984-3037, 2917-2982,
// 3414-3431).
//
// Slots are PARALLEL aux sends of the same source, summed onto the output bus in slot order
// (SURVEY.md section 0, fact 1); per output sample the additions happen in exactly the order the
// reference performs them: dry (input channel 0..C-1), then slot 0..3, each effect adding its
// contributions in its own fixed order.
#ifndef OALSFX_MIX_CUH
#define OALSFX_MIX_CUH

#include <cstddef>
#include <type_traits>

#include "fx.cuh"

namespace oalsfx {

constexpr int kSlotStateWords = 160;           // >= FxReverb::kStateWords (per slot, per lane)
constexpr int kSendCount = 1 + kMaxSlots;       // direct + aux
constexpr int kSendStateWords = kSendCount * kMaxChannels * 8; // lp + hp history per send per input channel

static_assert(FxReverb::kStateWords <= kSlotStateWords, "slot state too small");

// One tile covered by a launch and the lanes (streams) of it that take part.
struct TileRef { uint32_t tile, mask; };

// Kernel arguments of one launch: all streams covered by a launch share these coefficient blocks
// ("parameter class"), so they sit in the constant bank and cost no registers or loads.
struct MixArgs {
	int32_t frames;            // <= kMaxBlockFrames: the block every effect sees (oalsfxpp.cpp:3820, 2993)
	int32_t channels;
	int32_t num_streams;
	int32_t with_dry;          // this pass zero-fills the bus and mixes the direct send
	int32_t accumulate;        // this pass starts from the bus already in dst (multi-pass fallback)
	uint32_t update_mask;      // bit p: slot position p runs its `update` prologue (oalsfxpp.cpp:3145-3157)
	int32_t aux_index[kMaxSlots]; // slot position p -> engine slot (selects the aux send state)
	// I/O addressing: element (tile, lane, frame i, channel c) = base + tile*ts + lane*ls + i*fs + c*cs
	const float* src;
	float* dst;
	long long io_ts, io_ls, io_fs, io_cs;
	// Tiles covered by this launch: (tile index, lane mask); null = tiles 0..tile_count-1, all lanes.
	const TileRef* tiles;
	int32_t tile_count;
	int32_t tile_first;                 // identity mapping only: first tile of this launch (host-buffer pipelining slices the tiles)
	float* ring[kMaxSlots];          // per slot position: [tile][word][lane]
	long long ring_tile_stride[kMaxSlots];
	uint32_t* slot_state[kMaxSlots]; // per slot position: [tile][kSlotStateWords][lane]
	uint32_t* send_state;            // [tile][kSendStateWords][lane]
	SendCoef direct;
	SendCoef aux[kMaxSlots];
	SlotCoef slot[kMaxSlots];
	// Table mode (kTab* kernels): streams of one launch have DIFFERENT parameter sets.  The coefficient
	// blocks then live in HBM, one per parameter class, and every stream carries the index of its own:
	//   slot_table[lane_class[p][stream]]                      instead of slot[p]
	//   send_table[lane_send[stream] * kSendCount + 0 / 1 + i] instead of direct / aux[.] of engine slot i
	const SlotCoef* slot_table;
	const int32_t* lane_class[kMaxSlots];
	const SendCoef* send_table;
	const int32_t* lane_send;
	// Relay kernels (relay.cuh): slot position p = pipeline stage p (non-null slots, compacted in slot order).
	int32_t relay_count;              // stages
	int32_t relay_kind[kMaxSlots];    // FxKind of stage p
	int32_t relay_win[kMaxSlots];     // stage p's prefetch window: float offset into the dynamic shared memory
	int32_t relay_smem_floats;        // dynamic shared memory of the launch
	// Span kernels (span.cuh): frames per block-parallel span (host-checked against every delay of slot[0])
	int32_t span_frames;
	// Class-per-tile kernels (duo_multi_kernel): tile t runs with class_table[tile_class[t]]
	const struct MixClassEntry* class_table;
	const int32_t* tile_class;
};

// The per-class part of MixArgs: direct, aux[], slot[] (contiguous), plus the class's pending-update bits.
constexpr size_t kMixCoefOffset = offsetof(MixArgs, direct);
constexpr size_t kMixCoefBytes = offsetof(MixArgs, slot_table) - offsetof(MixArgs, direct);
static_assert(kMixCoefOffset % 4 == 0 && kMixCoefBytes % 4 == 0 && sizeof(MixArgs) % 4 == 0, "word copies");
struct MixClassEntry {
	uint32_t pending;
	uint32_t reserved[3];
	uint32_t coefs[kMixCoefBytes / 4];
};

// Send shelf filters (reference: apply_filters, oalsfxpp.cpp:3101-3143).  Pass-through still
// tracks the last two samples in both histories (oalsfxpp.cpp:1038-1056).
struct SendHist { BiquadHist lp, hp; };

OALSFX_HD void hist_pass(BiquadHist& h, float v)
{
	h.x1 = h.x0;
	h.x0 = v;
	h.y1 = h.y0;
	h.y0 = v;
}

OALSFX_HD float send_filter_step(const SendCoef& sc, SendHist& h, float x)
{
	switch (sc.filter_type) {
	case 1: {
		const float d = biquad_step(sc.lp, h.lp, x);
		hist_pass(h.hp, d);
		return d;
	}
	case 2: {
		hist_pass(h.lp, x);
		return biquad_step(sc.hp, h.hp, x);
	}
	case 3: {
		const float d = biquad_step(sc.lp, h.lp, x);
		return biquad_step(sc.hp, h.hp, d);
	}
	default:
		hist_pass(h.lp, x);
		hist_pass(h.hp, x);
		return x;
	}
}

// One slot position: encode the source into the slot's 4-channel wet bus (MixHelpers::mix with
// static gains, oalsfxpp.cpp:2952-2980) and run the effect on it.
// TABLE: the coefficient blocks come from the per-stream tables in HBM (MixArgs::slot_table ...) instead of
// the kernel arguments; everything else is identical.
template <int CT, bool SF, class Fx, bool TABLE = false>
struct SlotRunner {
	Fx fx;
	SendHist hist[SF ? (CT ? CT : kMaxChannels) : 1];   // indexed by unrolled loops only: registers, not local memory
	const SlotCoef* tab_slot = nullptr;
	const SendCoef* tab_send = nullptr;

	OALSFX_HD const SlotCoef& coef(const MixArgs& a, int p) const { return TABLE ? *tab_slot : a.slot[p]; }
	OALSFX_HD const SendCoef& send(const MixArgs& a, int p) const { return TABLE ? *tab_send : a.aux[p]; }

	OALSFX_HD void begin(const MixArgs& a, int p, int tile, int lane, float* prefetch_column)
	{
		if (Fx::kIsNull) {
			return;
		}
		fx.set_prefetch(prefetch_column);
		if (TABLE) {
			const long long stream = static_cast<long long>(tile) * kLanes + lane;
			tab_slot = a.slot_table + a.lane_class[p][stream];
			tab_send = a.send_table + static_cast<long long>(a.lane_send[stream]) * kSendCount + 1 + a.aux_index[p];
		}
		uint32_t* st = a.slot_state[p] + (static_cast<long long>(tile) * kSlotStateWords) * kLanes + lane;
		float* ring = a.ring[p] ? a.ring[p] + static_cast<long long>(tile) * a.ring_tile_stride[p] + lane : nullptr;
		fx.template begin<CT>(coef(a, p), st, ring, (a.update_mask >> p) & 1U, a.frames, a.channels);
		if (SF) {
			const uint32_t* ss = a.send_state + (static_cast<long long>(tile) * kSendStateWords +
				(1 + a.aux_index[p]) * kMaxChannels * 8) * kLanes + lane;
			OALSFX_UNROLL
			for (int c = 0; c < (CT ? CT : kMaxChannels); ++c) {
				if (CT || c < a.channels) {
					load_words(hist[SF ? c : 0], ss + c * 8 * kLanes);
				}
			}
		}
	}

	OALSFX_HD void step(const MixArgs& a, int p, const float* x, float* acc)
	{
		if (Fx::kIsNull) {
			return;
		}
		const SendCoef& sc = send(a, p);
		float wet[kWetChannels] = {0.0F, 0.0F, 0.0F, 0.0F};
		if (CT == 2 && !SF) {
			// wet[k] = (0 + x0 * g[0][k]) + x1 * g[1][k], wet channels (0,1) and (2,3) as pairs
			const F2 zero = f2(0.0F, 0.0F), x0 = f2_bcast(x[0]), x1 = f2_bcast(x[1]);
			const F2 wa = (zero + (x0 * f2(sc.gains[0][0], sc.gains[0][1]))) + (x1 * f2(sc.gains[1][0], sc.gains[1][1]));
			const F2 wb = (zero + (x0 * f2(sc.gains[0][2], sc.gains[0][3]))) + (x1 * f2(sc.gains[1][2], sc.gains[1][3]));
			wet[0] = f2_lo(wa);
			wet[1] = f2_hi(wa);
			wet[2] = f2_lo(wb);
			wet[3] = f2_hi(wb);
			fx.template step<CT, true>(coef(a, p), wet, acc, a.channels);
		} else {
			OALSFX_UNROLL
			for (int c = 0; c < (CT ? CT : kMaxChannels); ++c) {
				if (CT || c < a.channels) {
					const float v = SF ? send_filter_step(sc, hist[SF ? c : 0], x[c]) : x[c];
					OALSFX_UNROLL
					for (int k = 0; k < kWetChannels; ++k) {
						if (!SF || audible(sc.gains[c][k])) {
							wet[k] += v * sc.gains[c][k];
						}
					}
				}
			}
			fx.template step<CT, !SF>(coef(a, p), wet, acc, a.channels);
		}
	}

	// Effect state only (the caller writes the send filter history itself).
	OALSFX_HD void end_state_only(const MixArgs& a, int p, int tile, int lane)
	{
		if (Fx::kIsNull) {
			return;
		}
		uint32_t* st = a.slot_state[p] + (static_cast<long long>(tile) * kSlotStateWords) * kLanes + lane;
		fx.template end_ct<CT>(coef(a, p), st, a.channels);
	}

	OALSFX_HD void end(const MixArgs& a, int p, int tile, int lane, const float* last1, const float* last2)
	{
		if (Fx::kIsNull) {
			return;
		}
		uint32_t* st = a.slot_state[p] + (static_cast<long long>(tile) * kSlotStateWords) * kLanes + lane;
		fx.template end_ct<CT>(coef(a, p), st, a.channels);
		uint32_t* ss = a.send_state + (static_cast<long long>(tile) * kSendStateWords +
			(1 + a.aux_index[p]) * kMaxChannels * 8) * kLanes + lane;
		OALSFX_UNROLL
		for (int c = 0; c < (CT ? CT : kMaxChannels); ++c) {
			if (CT || c < a.channels) {
				if (SF) {
					store_words(hist[SF ? c : 0], ss + c * 8 * kLanes);
				} else {
					SendHist h;
					h.lp.x0 = h.lp.y0 = h.hp.x0 = h.hp.y0 = last1[c];
					h.lp.x1 = h.lp.y1 = h.hp.x1 = h.hp.y1 = last2[c];
					store_words(h, ss + c * 8 * kLanes);
				}
			}
		}
	}
};

// Whole block for one stream.  SF = false ("fast" kernels) requires: no send has an active shelf
// filter, frames >= 2 (then every processed send's filter history is simply the last two input
// samples) and static gains sanitized by the host (inaudible -> exact 0, see pan_add in fx.cuh).
// Which slot position (if any) owns the warp's prefetch window: the first reverb.
template <class F0, class F1, class F2, class F3>
struct PrefetchUser {
	static constexpr int value = std::is_same<F0, FxReverb>::value ? 0 : std::is_same<F1, FxReverb>::value ? 1 :
		std::is_same<F2, FxReverb>::value ? 2 : std::is_same<F3, FxReverb>::value ? 3 : -1;
};

// `prefetch_column`: this thread's column of its warp's shared-memory prefetch window (kPfWarpFloats
// floats per warp), or null (CPU test build, kernels without a window) to read the rings directly.
template <int CT, bool SF, class F0, class F1, class F2, class F3, bool TABLE = false>
OALSFX_HD void mix_stream(const MixArgs& a, int tile, int lane, float* prefetch_column = nullptr)
{
	constexpr int pf_user = PrefetchUser<F0, F1, F2, F3>::value;
	const int channels = CT ? CT : a.channels;
	const float* src = a.src + tile * a.io_ts + lane * a.io_ls;
	float* dst = a.dst + tile * a.io_ts + lane * a.io_ls;

	SlotRunner<CT, SF, F0, TABLE> r0;
	SlotRunner<CT, SF, F1, TABLE> r1;
	SlotRunner<CT, SF, F2, TABLE> r2;
	SlotRunner<CT, SF, F3, TABLE> r3;
	const SendCoef& direct = TABLE ? a.send_table[static_cast<long long>(a.lane_send[static_cast<long long>(tile) * kLanes + lane]) * kSendCount] : a.direct;
	r0.begin(a, 0, tile, lane, pf_user == 0 ? prefetch_column : nullptr);
	r1.begin(a, 1, tile, lane, pf_user == 1 ? prefetch_column : nullptr);
	r2.begin(a, 2, tile, lane, pf_user == 2 ? prefetch_column : nullptr);
	r3.begin(a, 3, tile, lane, pf_user == 3 ? prefetch_column : nullptr);

	SendHist dhist[SF ? kMaxChannels : 1];
	uint32_t* dss = a.send_state + (static_cast<long long>(tile) * kSendStateWords) * kLanes + lane;
	if (SF && a.with_dry) {
		for (int c = 0; c < channels; ++c) {
			load_words(dhist[c], dss + c * 8 * kLanes);
		}
	}

	float last1[CT ? CT : kMaxChannels], last2[CT ? CT : kMaxChannels];
	OALSFX_UNROLL
	for (int c = 0; c < (CT ? CT : kMaxChannels); ++c) {
		last1[c] = last2[c] = 0.0F;
	}

	for (int i = 0; i < a.frames; ++i) {
		float x[CT ? CT : kMaxChannels];
		float acc[CT ? CT : kMaxChannels];
		OALSFX_UNROLL
		for (int c = 0; c < (CT ? CT : kMaxChannels); ++c) {
			if (CT || c < channels) {
				x[c] = src[i * a.io_fs + c * a.io_cs];
				acc[c] = a.accumulate ? dst[i * a.io_fs + c * a.io_cs] : 0.0F;
				last2[c] = last1[c];
				last1[c] = x[c];
			}
		}
		if (a.with_dry) {
			// Direct send: bus[k] += x_c * g[c][k] for c = 0..C-1 (oalsfxpp.cpp:2924-2950)
			OALSFX_UNROLL
			for (int c = 0; c < (CT ? CT : kMaxChannels); ++c) {
				if (CT || c < channels) {
					const float v = SF ? send_filter_step(direct, dhist[SF ? c : 0], x[c]) : x[c];
					OALSFX_UNROLL
					for (int k = 0; k < (CT ? CT : kMaxChannels); ++k) {
						if ((CT || k < channels) && (!SF || audible(direct.gains[c][k]))) {
							acc[k] += v * direct.gains[c][k];
						}
					}
				}
			}
		}
		r0.step(a, 0, x, acc);
		r1.step(a, 1, x, acc);
		r2.step(a, 2, x, acc);
		r3.step(a, 3, x, acc);
		OALSFX_UNROLL
		for (int c = 0; c < (CT ? CT : kMaxChannels); ++c) {
			if (CT || c < channels) {
				dst[i * a.io_fs + c * a.io_cs] = acc[c];
			}
		}
	}

	r0.end(a, 0, tile, lane, last1, last2);
	r1.end(a, 1, tile, lane, last1, last2);
	r2.end(a, 2, tile, lane, last1, last2);
	r3.end(a, 3, tile, lane, last1, last2);
	if (a.with_dry) {
		for (int c = 0; c < channels; ++c) {
			if (SF) {
				store_words(dhist[c], dss + c * 8 * kLanes);
			} else {
				SendHist h;
				h.lp.x0 = h.lp.y0 = h.hp.x0 = h.hp.y0 = last1[c];
				h.lp.x1 = h.lp.y1 = h.hp.x1 = h.hp.y1 = last2[c];
				store_words(h, dss + c * 8 * kLanes);
			}
		}
	}
}

#if defined(__CUDACC__)
template <int CT, bool SF, class F0, class F1, class F2, class F3, bool TABLE = false>
__global__ void __launch_bounds__(64) mix_kernel(const __grid_constant__ MixArgs a)
{
	// Table mode: taps differ from lane to lane, so the reverb's warp-wide batched prefetch is off.
	constexpr bool has_window = !TABLE && PrefetchUser<F0, F1, F2, F3>::value >= 0;
	__shared__ __align__(16) float window[has_window ? (64 / kLanes) * kPfWarpFloats : 1];
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
	const int lane = threadIdx.x % kLanes;
	if (warp >= a.tile_count) {
		return;
	}
	int tile = a.tile_first + warp;
	uint32_t mask = 0xFFFFFFFFU;
	if (a.tiles) {
		const TileRef t = a.tiles[warp];
		tile = static_cast<int>(t.tile);
		mask = t.mask;
	}
	if (!((mask >> lane) & 1U) || tile * kLanes + lane >= a.num_streams) {
		return;
	}
	// The reverb's batched prefetch is a whole-warp operation: only complete tiles get a window.
	const bool whole_tile = mask == 0xFFFFFFFFU && (tile + 1) * kLanes <= a.num_streams;
	mix_stream<CT, SF, F0, F1, F2, F3, TABLE>(a, tile, lane,
		has_window && whole_tile ? window + (threadIdx.x / kLanes) * kPfWarpFloats + lane : nullptr);
}
#endif

} // namespace oalsfx

#endif
