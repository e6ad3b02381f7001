// relay.cuh -- ANY slot signature as a warp pipeline: one warp per non-null effect slot of a 32-stream tile.
//
// The fused duo / quartet kernels are compiled per signature (equalizer + chorus + echo + reverb, a single
// reverb ...).  Every other combination of the ten effect types used to run as a chain of single-effect
// passes: one launch per slot, each re-reading the input and read-modify-writing the output in HBM, and one
// warp per tile doing a long dependent instruction sequence per sample.  The relay kernel covers all of them
// in ONE launch and ONE pass over the I/O buffers:
//
//   stage 0          : input frame (cp.async window), dry mix (direct send), first non-null slot
//   stage 1 .. n-1   : the next non-null slots, in slot order
//   last stage       : + output rows
//
// The effect a stage runs is chosen at run time (MixArgs::relay_kind, one warp-uniform switch per launch),
// the stage POSITION is a template parameter so that its coefficient block a.slot[P] stays a constant-bank
// operand.  A stage hands the input frame and the running output bus to the next one through a double-
// buffered shared-memory exchange of kChunk frames guarded by named barriers (the protocol of quartet.cuh),
// so per output sample the additions happen in the reference's order: dry, then slot by slot
// (oalsfxpp.cpp:2984-3037).  Slots are parallel aux sends of the same source, so a
// stage needs nothing from its predecessor but those two vectors.
//
// LIGHT kernels leave the reverb out (128 registers, 4 CTAs per SM); HEAVY ones (any reverb among the slots)
// run the unsplit reverb in one warp with its batched cp.async window (up to 255 registers).  Windows live in
// dynamic shared memory at host-computed offsets (MixArgs::relay_win), so a launch only pays for the windows
// its own effects use.
//
// Host-checked requirements (as for the duo kernels): no send shelf filter active, no unstable filter,
// frames >= 2, every tile of the launch takes part with all of its lanes, sanitized static gains,
// one or two device channels.
#ifndef OALSFX_RELAY_CUH
#define OALSFX_RELAY_CUH

#if defined(__CUDACC__)

#include "kernel_table.h"
#include "quartet.cuh"

namespace oalsfx {
namespace relay {

#ifndef OALSFX_RELAY_CHUNK
#define OALSFX_RELAY_CHUNK 4
#endif
constexpr int kChunk = 4;                 // frames per hand-off of the wide kernels (their exchange is 68 KB as it is)
template <int CT>
struct ChunkOf { static constexpr int value = CT ? OALSFX_RELAY_CHUNK : kChunk; }; // frames per hand-off (a multiple of the output row batch of 4)
constexpr int kThreads = kMaxSlots * kLanes;

// CT = 0: the device channel count is a run-time value (quad, 5.1, 6.1, 7.1: MixArgs::channels), arrays are sized for
// kMaxChannels.
template <int CT>
struct Shared {
	static constexpr int NC = CT ? CT : kMaxChannels;
	static constexpr int K = ChunkOf<CT>::value;
	static constexpr int WIN = K > kFwSlots ? K : kFwSlots; // (a reverb in stage 0 parks a whole hand-off of input frames here)
	float xch[kMaxSlots][2][K][2 * NC][kLanes];      // hand-off P -> P+1: x_0..x_C-1, bus_0..bus_C-1 (the last stage parks finished frames in its own)
	float win_in[WIN][NC][kLanes];                   // stage 0: input frames in flight
};

// Where the exchange lives: static shared memory for one and two channels (16 / 28 KB), the head of the dynamic
// allocation for the wide kernels (68 KB, beyond the static limit).
template <int CT>
struct SharedPlace {
	static __device__ __forceinline__ Shared<CT>& get(float*&)
	{
		__shared__ __align__(16) Shared<CT> sh;
		return sh;
	}
};
template <>
struct SharedPlace<0> {
	static __device__ __forceinline__ Shared<0>& get(float*& dyn)
	{
		Shared<0>* sh = reinterpret_cast<Shared<0>*>(dyn);
		dyn += sizeof(Shared<0>) / sizeof(float);
		return *sh;
	}
};

// No shelf filter active: a processed send's filter histories are the last two input samples (oalsfxpp.cpp:1038-1056).
template <int CT>
__device__ __forceinline__ void store_passthrough_history(uint32_t* ss, int send, const float* src, const MixArgs& a, bool io_ok)
{
	if (CT) {
		duo::store_passthrough_history<CT ? CT : 1>(ss, send, src, a, io_ok);
	} else {
		for (int c = 0; c < a.channels; ++c) {
			const float last1 = io_ok ? src[(a.frames - 1) * a.io_fs + c * a.io_cs] : 0.0F;
			const float last2 = io_ok ? src[(a.frames - 2) * a.io_fs + c * a.io_cs] : 0.0F;
			SendHist h;
			h.lp.x0 = h.lp.y0 = h.hp.x0 = h.hp.y0 = last1;
			h.lp.x1 = h.lp.y1 = h.hp.x1 = h.hp.y1 = last2;
			store_words(h, ss + (send * kMaxChannels + c) * 8 * kLanes);
		}
	}
}

// SF: the sends' shelf filters are compiled in (apply_filters, oalsfxpp.cpp:3101-3143): every stage filters the input
// frame with ITS aux send's filters before the wet encode, stage 0 the direct send's as well, and the filter histories
// are state (oalsfxpp.cpp:1097-1113) instead of "the last two input samples".
template <int CT, int P, class Fx, bool SF = false>
__device__ __forceinline__ void stage(const MixArgs& a, Shared<CT>& sh, float* dyn, const int tile, const int lane)
{
	constexpr bool kReverb = std::is_same<Fx, FxReverb>::value;
	constexpr bool kMod = std::is_same<Fx, FxModDelay>::value;
	constexpr bool kEcho = std::is_same<Fx, FxEcho>::value;
	constexpr int NC = CT ? CT : kMaxChannels;      // array extents
	constexpr int kChunk = ChunkOf<CT>::value;      // frames per hand-off
	constexpr int kWinMask = (kReverb ? Shared<CT>::WIN : kFwSlots) - 1;
	constexpr int CTD = CT ? CT : 1;                // divisor of the vector-store path (one and two channels only)
	const int nch = CT ? CT : a.channels;
	constexpr int HP = P > 0 ? P - 1 : 0;           // hand-off this stage reads
	constexpr int HN = P < kMaxSlots - 1 ? P : 0;   // hand-off this stage writes (unless it is the last)
	const bool last = P + 1 == a.relay_count;       // warp-uniform
	const bool io_ok = tile * kLanes + lane < a.num_streams;
	const float* src = a.src + tile * a.io_ts + lane * a.io_ls;
	float* dst = a.dst + tile * a.io_ts + lane * a.io_ls;
	uint32_t* ss = a.send_state + (static_cast<long long>(tile) * kSendStateWords) * kLanes + lane;
	const int chunks = (a.frames + kChunk - 1) / kChunk;
	const bool fast_out = CT > 0 && a.io_cs == 1 && a.io_fs == CT && (a.frames % kChunk) == 0 && (a.io_ls % 4) == 0 &&
		(a.io_ts % 4) == 0 && (reinterpret_cast<unsigned long long>(a.dst) & 15ULL) == 0;

	SlotRunner<CT, SF, Fx> r;
	float* win = dyn + a.relay_win[P];
	SendHist dhist[SF ? NC : 1];   // stage 0: the direct send's filter histories
	if (SF && P == 0) {
#pragma unroll
		for (int c = 0; c < NC; ++c) {
			if (CT || c < nch) {
				load_words(dhist[SF ? c : 0], ss + c * 8 * kLanes);
			}
		}
	}
	r.begin(a, P, tile, lane, (kReverb || kMod) ? win + lane : kEcho ? win + lane + 2 * kLanes : nullptr);

	// Stage 0 input.  Through the cp.async window, kFwDepth frames ahead, in the same commit groups as the
	// effect's own ring prefetches; a reverb runs its own cp.async pipeline inside step(), so there the frames
	// of the next hand-off are loaded into registers one hand-off early and parked in the same window.
	const unsigned win_s = smem_addr(&sh.win_in[0][0][lane]);
	const float* in = src;
	auto issue_input = [&](int frame) {
		if (P == 0 && !kReverb) {
			if (io_ok && frame < a.frames) {
				const unsigned slot = win_s + static_cast<unsigned>((frame & (kFwSlots - 1)) * NC * kLanes * 4);
#pragma unroll
				for (int c = 0; c < NC; ++c) {
					if (CT || c < nch) {
						cp_async_f32_s(slot + c * kLanes * 4, in + c * a.io_cs);
					}
				}
			}
			in += a.io_fs;
		}
	};
	float nx[kChunk][NC];
	auto load_rows = [&](int first) {
#pragma unroll
		for (int f = 0; f < kChunk; ++f) {
#pragma unroll
			for (int c = 0; c < NC; ++c) {
				nx[f][c] = (io_ok && (CT || c < nch) && first + f < a.frames) ? src[(first + f) * a.io_fs + c * a.io_cs] : 0.0F;
			}
		}
	};
	if (P == 0 && kReverb) {
		load_rows(0);
	}
	if (!kReverb) {
		for (int k = 0; k < kFwDepth; ++k) {
			r.fx.prefetch_issue(a.slot[P], k);
			issue_input(k);
			cp_async_commit_group();
		}
	}
	if (P > 0) {
		quartet::signal_empty<HP>(0);
		quartet::signal_empty<HP>(1);
	}
	for (int ci = 0; ci < chunks; ++ci) {
		const int b = ci & 1;
		const int first = ci * kChunk;
		const int count = min(kChunk, a.frames - first);
		if (P == 0 && kReverb) {
			// this hand-off's frames go to the thread's own window column, the next one's are requested
#pragma unroll
			for (int f = 0; f < kChunk; ++f) {
#pragma unroll
				for (int c = 0; c < NC; ++c) {
					if (CT || c < nch) {
						sh.win_in[(first + f) & kWinMask][c][lane] = nx[f][c];
					}
				}
			}
			load_rows(first + kChunk);
		}
		if (P > 0) {
			quartet::wait_full<HP>(b);
		}
		if (P < kMaxSlots - 1 && !last) {
			quartet::wait_empty<HN>(b);
		}
		for (int f = 0; f < count; ++f) {
			const int i = first + f;
			float x[NC], acc[NC];
			if (!kReverb) {
				r.fx.prefetch_next(a.slot[P]);
				issue_input(i + kFwDepth);
				cp_async_commit_group();
				cp_async_wait_group<kFwDepth>();
			}
			if (P == 0) {
#pragma unroll
				for (int c = 0; c < NC; ++c) {
					x[c] = (io_ok && (CT || c < nch)) ? sh.win_in[i & kWinMask][c][lane] : 0.0F;
					acc[c] = 0.0F;
				}
				// direct send (oalsfxpp.cpp:2924-2950); gains sanitized by the host
#pragma unroll
				for (int c = 0; c < NC; ++c) {
					if (CT || c < nch) {
						pan_add<CT, true>(acc, nch, a.direct.gains[c], SF ? send_filter_step(a.direct, dhist[SF ? c : 0], x[c]) : x[c]);
					}
				}
			} else {
#pragma unroll
				for (int c = 0; c < NC; ++c) {
					if (CT || c < nch) {
						x[c] = sh.xch[HP][b][f][c][lane];
						acc[c] = sh.xch[HP][b][f][NC + c][lane];
					}
				}
			}
			r.step(a, P, x, acc);
			if (last && !fast_out) {
				if (io_ok) {
#pragma unroll
					for (int c = 0; c < NC; ++c) {
						if (CT || c < nch) {
							dst[i * a.io_fs + c * a.io_cs] = acc[c];
						}
					}
				}
			} else {
#pragma unroll
				for (int c = 0; c < NC; ++c) {
					if (CT || c < nch) {
						if (!last) {
							sh.xch[P][b][f][c][lane] = x[c];
						}
						sh.xch[P][b][f][NC + c][lane] = acc[c];
					}
				}
			}
		}
		if (last) {
			if (fast_out && io_ok) {
				// kChunk frames x CT channels of this thread's row, 16 bytes at a time (own column: no barrier)
				float4* row = reinterpret_cast<float4*>(dst + first * CT);
				static_assert(kChunk % 4 == 0, "output rows are written four frames at a time");
#pragma unroll
				for (int g = 0; g < kChunk * CT / 4; ++g) {
					const int e0 = 4 * g; // element index within the chunk: frame = e / CT, channel = e % CT
					__stcs(row + g, make_float4(sh.xch[P][b][(e0 + 0) / CTD][NC + (e0 + 0) % CTD][lane], sh.xch[P][b][(e0 + 1) / CTD][NC + (e0 + 1) % CTD][lane],
						sh.xch[P][b][(e0 + 2) / CTD][NC + (e0 + 2) % CTD][lane], sh.xch[P][b][(e0 + 3) / CTD][NC + (e0 + 3) % CTD][lane]));
				}
			}
		} else if (P < kMaxSlots - 1) {
			__threadfence_block();
			quartet::signal_full<HN>(b);
		}
		if (P > 0 && ci + 2 < chunks) {
			quartet::signal_empty<HP>(b); // nobody waits for the last two drains
		}
	}
	if (!kReverb) {
		cp_async_wait_group<0>();
	}
	if (SF) {
		r.end(a, P, tile, lane, nullptr, nullptr);   // effect state + this send's filter histories
		if (P == 0) {
#pragma unroll
			for (int c = 0; c < NC; ++c) {
				if (CT || c < nch) {
					store_words(dhist[SF ? c : 0], ss + c * 8 * kLanes);
				}
			}
		}
		return;
	}
	r.end_state_only(a, P, tile, lane);
	if (P == 0) {
		store_passthrough_history<CT>(ss, 0, src, a, io_ok);
	}
	store_passthrough_history<CT>(ss, 1 + a.aux_index[P], src, a, io_ok);
}

template <int CT, bool HEAVY, int P, bool SF>
__device__ __forceinline__ void dispatch(const MixArgs& a, Shared<CT>& sh, float* dyn, int tile, int lane)
{
	switch (a.relay_kind[P]) {
	case kKindModDelay: stage<CT, P, FxModDelay, SF>(a, sh, dyn, tile, lane); break;
	case kKindCompressor: stage<CT, P, FxCompressor, SF>(a, sh, dyn, tile, lane); break;
	case kKindDedicated: stage<CT, P, FxDedicated, SF>(a, sh, dyn, tile, lane); break;
	case kKindDistortion: stage<CT, P, FxDistortion, SF>(a, sh, dyn, tile, lane); break;
	case kKindEcho: stage<CT, P, FxEcho, SF>(a, sh, dyn, tile, lane); break;
	case kKindEqualizer: stage<CT, P, FxEqualizer, SF>(a, sh, dyn, tile, lane); break;
	case kKindRingMod: stage<CT, P, FxRingMod, SF>(a, sh, dyn, tile, lane); break;
	case kKindReverb:
		if (HEAVY) {
			stage<CT, P, typename std::conditional<HEAVY, FxReverb, FxDedicated>::type, SF>(a, sh, dyn, tile, lane);
		}
		break;
	default: break;
	}
}

template <int CT, bool HEAVY, bool SF = false>
__device__ __forceinline__ void relay_body(const MixArgs& a)
{
	extern __shared__ __align__(16) float dyn_base[];
	float* dyn = dyn_base;
	Shared<CT>& sh = SharedPlace<CT>::get(dyn);
	const int tile = a.tiles ? static_cast<int>(a.tiles[blockIdx.x].tile) : a.tile_first + static_cast<int>(blockIdx.x);
	const int lane = threadIdx.x % kLanes;
	// A CTA has one warp per stage (the launch gives it 32 x relay_count threads: a two-stage signature then holds half
	// the registers of a four-stage one, and twice the CTAs fit an SM).  Warp w of every CTA lands on scheduler w:
	// rotate the stages so each scheduler sees all of them.
	const int st = (static_cast<int>(threadIdx.x / kLanes) + static_cast<int>(blockIdx.x)) % a.relay_count;
	switch (st) {
	case 0: dispatch<CT, HEAVY, 0, SF>(a, sh, dyn, tile, lane); break;
	case 1: dispatch<CT, HEAVY, 1, SF>(a, sh, dyn, tile, lane); break;
	case 2: dispatch<CT, HEAVY, 2, SF>(a, sh, dyn, tile, lane); break;
	default: dispatch<CT, HEAVY, 3, SF>(a, sh, dyn, tile, lane); break;
	}
}

template <int CT, bool HEAVY>
__global__ void __launch_bounds__(kThreads, HEAVY ? 2 : 4) relay_kernel(const __grid_constant__ MixArgs a)
{
	relay_body<CT, HEAVY>(a);
}

// The same with the sends' shelf filters compiled in.
template <int CT, bool HEAVY>
__global__ void __launch_bounds__(kThreads, HEAVY ? 2 : 4) relay_sf_kernel(const __grid_constant__ MixArgs a)
{
	relay_body<CT, HEAVY, true>(a);
}

// One parameter class PER TILE (as duo_multi_kernel, duo.cuh): the tile's coefficient blocks and pending bits come
// from the class table in HBM, assembled with the launch-wide arguments in shared memory.
template <int CT, bool HEAVY>
__global__ void __launch_bounds__(kThreads, HEAVY ? 2 : 4) relay_multi_kernel(const __grid_constant__ MixArgs a)
{
	__shared__ __align__(16) MixArgs sa;
	const int tile = a.tiles ? static_cast<int>(a.tiles[blockIdx.x].tile) : a.tile_first + static_cast<int>(blockIdx.x);
	const MixClassEntry* entry = a.class_table + a.tile_class[tile];
	const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
	uint32_t* ps = reinterpret_cast<uint32_t*>(&sa);
	constexpr int kWords = static_cast<int>(sizeof(MixArgs) / 4), kCoef0 = static_cast<int>(kMixCoefOffset / 4),
		kCoefWords = static_cast<int>(kMixCoefBytes / 4);
	for (int i = threadIdx.x; i < kWords; i += static_cast<int>(blockDim.x)) {
		if (i < kCoef0 || i >= kCoef0 + kCoefWords) {
			ps[i] = pa[i];
		}
	}
	for (int i = threadIdx.x; i < kCoefWords; i += static_cast<int>(blockDim.x)) {
		ps[kCoef0 + i] = entry->coefs[i];
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		sa.update_mask = a.update_mask & entry->pending;
	}
	__syncthreads();
	relay_body<CT, HEAVY>(sa);
}

} // namespace relay
} // namespace oalsfx

#endif // __CUDACC__
#endif
