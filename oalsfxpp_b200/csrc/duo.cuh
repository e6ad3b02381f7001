// duo.cuh -- the fused 4-slot kernel as TWO specialised warps per 32-stream tile.
//
// One thread still owns one stream (coefficients stay warp-uniform constant-bank operands, every
// ring access of a warp is one full 128-byte line), but the slots are split over two warps so that
// each keeps only its own recurrent state in registers:
//
//   front warp : source encode + dry mix + slots 0..2 + the reverb's input stage
//                                                            (cfg4: equalizer, chorus, echo; B->A + shelves + main-line feed)
//   back  warp : the rest of slot 3 + output rows            (cfg4: EAX reverb early/late halves, batched 16-byte cp.async window)
//
// The front warp hands the input frame and the partially summed bus to the back warp through a
// double-buffered shared-memory exchange (kDuoChunk frames per hand-off, named barriers), so the
// bus never touches HBM and the per-sample summation order stays dry, slot 0, 1, 2, 3 -- exactly
// the reference's (oalsfxpp.cpp:2984-3037).  Per-thread registers drop from ~240 to 168,
// which doubles the warps in flight, and the two halves of a stream's work overlap in time.
// Measured operating point (B200, profiles/r01_history.md): 25 KB of shared memory per CTA, 5 CTAs per SM
// with the 135 KB carve-out and the remaining ~93 KB as L1 for the stream-major input rows: 3.07 ms per
// 1024-frame block of 65 536 streams.
//
// Host-checked requirements (as for the quad kernels): no send shelf filter active, frames >= 2,
// every tile of the launch takes part with all of its lanes.
#ifndef OALSFX_DUO_CUH
#define OALSFX_DUO_CUH

#if defined(__CUDACC__)

#include "mix.cuh"

namespace oalsfx {
namespace duo {

#ifndef OALSFX_DUO_MIN_CTAS
#define OALSFX_DUO_MIN_CTAS 6          // resident CTAs per SM the register allocation is sized for
#endif

#ifndef OALSFX_DUO_UNROLL
#define OALSFX_DUO_UNROLL 1            // sample-loop unroll factor (2 lets ptxas rename the filter histories instead of moving them)
#endif

constexpr int kDuoUnroll = OALSFX_DUO_UNROLL;
#ifndef OALSFX_DUO_UNROLL_FRONT
#define OALSFX_DUO_UNROLL_FRONT OALSFX_DUO_UNROLL
#endif
#ifndef OALSFX_DUO_UNROLL_BACK
#define OALSFX_DUO_UNROLL_BACK OALSFX_DUO_UNROLL
#endif
#ifndef OALSFX_DUO_CHUNK
#define OALSFX_DUO_CHUNK 4
#endif
constexpr int kDuoUnrollFront = OALSFX_DUO_UNROLL_FRONT, kDuoUnrollBack = OALSFX_DUO_UNROLL_BACK;
constexpr int kDuoChunk = OALSFX_DUO_CHUNK; // frames per hand-off (a multiple of the 4-frame output row batch)
constexpr int kBarFull = 0;            // named barriers 0,1: buffer b filled by the front warp
constexpr int kBarEmpty = 2;           // named barriers 2,3: buffer b drained by the back warp

// The barrier id must be an immediate: with a register operand ptxas reserves all 16 hardware
// barriers for the CTA, which caps residency at 4 CTAs per SM (ncu: launch__occupancy_limit_barriers).
template <int ID> __device__ __forceinline__ void bar_sync_imm() { asm volatile("bar.sync %0, 64;" ::"n"(ID) : "memory"); }
template <int ID> __device__ __forceinline__ void bar_arrive_imm() { asm volatile("bar.arrive %0, 64;" ::"n"(ID) : "memory"); }
template <int BASE> __device__ __forceinline__ void bar_sync(int b)
{
	if (b) {
		bar_sync_imm<BASE + 1>();
	} else {
		bar_sync_imm<BASE>();
	}
}
template <int BASE> __device__ __forceinline__ void bar_arrive(int b)
{
	if (b) {
		bar_arrive_imm<BASE + 1>();
	} else {
		bar_arrive_imm<BASE>();
	}
}

template <int CT>
__device__ __forceinline__ void store_passthrough_history(uint32_t* ss, int send, const float* src, const MixArgs& a, bool io_ok)
{
	// No shelf filter active: a processed send's filter histories are the last two input samples
	// (oalsfxpp.cpp:1038-1056).
#pragma unroll
	for (int c = 0; c < CT; ++c) {
		const float last1 = io_ok ? src[(a.frames - 1) * a.io_fs + c * a.io_cs] : 0.0F;
		const float last2 = io_ok ? src[(a.frames - 2) * a.io_fs + c * a.io_cs] : 0.0F;
		SendHist h;
		h.lp.x0 = h.lp.y0 = h.hp.x0 = h.hp.y0 = last1;
		h.lp.x1 = h.lp.y1 = h.hp.x1 = h.hp.y1 = last2;
		store_words(h, ss + (send * kMaxChannels + c) * 8 * kLanes);
	}
}

template <int CT, class F0, class F1, class F2, class F3>
__device__ __forceinline__ void duo_body(const MixArgs& a)
{
	// A reverb in slot 3 is split: its input stage (B->A conversion, shelf filters, main-line feed,
	// ~110 instructions per sample) runs in the front warp, which balances the two warps (500 / 413).
	// The front warp runs at most two hand-offs (8 frames) ahead; the main ring keeps 256 spare frames
	// beyond its longest tap (oalsfxpp.cpp:6573), so feeding it early cannot overwrite a pending read.
	constexpr bool back_has_window = std::is_same<F3, FxReverb>::value;
	constexpr bool split_reverb = back_has_window;
	using Front3 = typename std::conditional<split_reverb, FxReverbInput, FxNull>::type;
	using Back3 = typename std::conditional<split_reverb, FxReverbTail, F3>::type;
	__shared__ __align__(16) float window[back_has_window ? kPfWarpFloats : 1];
	constexpr int kBusAt = split_reverb ? 0 : CT;       // the input frame is only carried when the back warp needs it
	__shared__ float xch[2][kDuoChunk][kBusAt + CT][kLanes]; // [buffer][frame][(x_0..x_C-1,) bus_0..bus_C-1][lane]
	__shared__ float fwin[kFwWarpFloats];               // front warp: chorus/echo taps and input frames in flight
	static_assert(4 + CT <= kFwTaps, "front window too small for this channel count");

	const int tile = a.tiles ? static_cast<int>(a.tiles[blockIdx.x].tile) : a.tile_first + static_cast<int>(blockIdx.x);
	const int lane = threadIdx.x % kLanes;
	const bool front = threadIdx.x < kLanes;
	const bool io_ok = tile * kLanes + lane < a.num_streams;
	const float* src = a.src + tile * a.io_ts + lane * a.io_ls;
	float* dst = a.dst + tile * a.io_ts + lane * a.io_ls;
	uint32_t* ss = a.send_state + (static_cast<long long>(tile) * kSendStateWords) * kLanes + lane;
	const int chunks = (a.frames + kDuoChunk - 1) / kDuoChunk;
	// Row-batched I/O: with interleaved frames (the reference's buffer layout, one row per stream) a
	// thread's kDuoChunk = 4 frames x 2 channels are 32 contiguous bytes of its row -- one full sector,
	// moved with two 16-byte accesses per hand-off instead of eight scattered 4-byte ones (which lean on
	// L1 to merge 16 frames per line and reach L2 as partial-sector writes).
// Measured (gpurun_out/exp15): batching the OUTPUT rows helps (5 CTAs/SM: 3.38 -> 3.19 ms); batching the
// INPUT rows hurts (3.17 -> 3.7 ms): 4-byte requests let L1 fetch each 128-byte row segment from DRAM once
// and serve 16 frames from it, 16-byte requests fetch it as four separate sectors 4 frames apart.
#ifndef OALSFX_DUO_FAST_IN
#define OALSFX_DUO_FAST_IN 0
#endif
#ifndef OALSFX_DUO_FAST_OUT
#define OALSFX_DUO_FAST_OUT 1
#endif
	static_assert(kDuoChunk % 4 == 0 && (kDuoChunk == 4 || !OALSFX_DUO_FAST_IN), "output rows are written four frames at a time");
	__shared__ float xin[OALSFX_DUO_FAST_IN ? kDuoChunk : 1][CT][kLanes]; // front warp: the current hand-off's input frames
	const bool fast_rows = CT == 2 && a.io_cs == 1 && a.io_fs == CT && (a.frames % kDuoChunk) == 0 && (a.io_ls % 4) == 0 &&
		(a.io_ts % 4) == 0 && ((reinterpret_cast<unsigned long long>(a.src) | reinterpret_cast<unsigned long long>(a.dst)) & 15ULL) == 0;
	const bool fast_io = fast_rows && OALSFX_DUO_FAST_IN;   // input side
	const bool fast_out = fast_rows && OALSFX_DUO_FAST_OUT; // output side

	if (front) {
		SlotRunner<CT, false, F0> r0;
		SlotRunner<CT, false, F1> r1;
		SlotRunner<CT, false, F2> r2;
		SlotRunner<CT, false, Front3> r3in;
		r3in.begin(a, 3, tile, lane, nullptr);
		// Window taps: 0,1 = the first chorus/flanger, 2,3 = the first echo, 4.. = input channels.
		constexpr bool m0 = std::is_same<F0, FxModDelay>::value, m1 = std::is_same<F1, FxModDelay>::value && !m0,
			m2 = std::is_same<F2, FxModDelay>::value && !m0 && !m1;
		constexpr bool e0 = std::is_same<F0, FxEcho>::value, e1 = std::is_same<F1, FxEcho>::value && !e0,
			e2 = std::is_same<F2, FxEcho>::value && !e0 && !e1;
		float* col = fwin + lane;
		r0.begin(a, 0, tile, lane, m0 ? col : e0 ? col + 2 * kLanes : nullptr);
		r1.begin(a, 1, tile, lane, m1 ? col : e1 ? col + 2 * kLanes : nullptr);
		r2.begin(a, 2, tile, lane, m2 ? col : e2 ? col + 2 * kLanes : nullptr);
		// Everything long-latency of sample i + kFwDepth is requested while sample i is computed.
		const unsigned col_s = smem_addr(col);
		const float* in = src;                // input frame being requested (advances with `frame`)
		auto issue_input = [&](int frame) {
			if (!fast_io && io_ok && frame < a.frames) {
				const unsigned slot = col_s + static_cast<unsigned>(((frame & (kFwSlots - 1)) * kFwSlotFloats + 4 * kLanes) * 4);
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					cp_async_f32_s(slot + c * kLanes * 4, in + c * a.io_cs);
				}
			}
			in += a.io_fs;
		};
		float4 nx0 = make_float4(0.0F, 0.0F, 0.0F, 0.0F), nx1 = nx0; // the next hand-off's input frames (fast_io)
		if (fast_io && io_ok) {
			const float4* row = reinterpret_cast<const float4*>(src);
			nx0 = __ldcs(row);
			nx1 = __ldcs(row + 1);
		}
		for (int k = 0; k < kFwDepth; ++k) { // prime: samples 0 .. kFwDepth-1
			r0.fx.prefetch_issue(a.slot[0], k);
			r1.fx.prefetch_issue(a.slot[1], k);
			r2.fx.prefetch_issue(a.slot[2], k);
			issue_input(k);
			cp_async_commit_group();
		}
		auto issue = [&](int frame) {        // steady state: sample `frame` = current + kFwDepth
			r0.fx.prefetch_next(a.slot[0]);
			r1.fx.prefetch_next(a.slot[1]);
			r2.fx.prefetch_next(a.slot[2]);
			issue_input(frame);
			cp_async_commit_group();
		};
		for (int ci = 0; ci < chunks; ++ci) {
			const int b = ci & 1;
			const int first = ci * kDuoChunk;
			const int count = min(kDuoChunk, a.frames - first);
			if (fast_io) {
				// this hand-off's frames were requested one hand-off ago; request the next one's
				const float4 c0 = nx0, c1 = nx1;
				if (io_ok && ci + 1 < chunks) {
					const float4* row = reinterpret_cast<const float4*>(src + (first + kDuoChunk) * CT);
					nx0 = __ldcs(row);
					nx1 = __ldcs(row + 1);
				}
				constexpr int kIn = OALSFX_DUO_FAST_IN ? 1 : 0; // (row index folded to 0 when the path is compiled out)
				xin[0][0][lane] = c0.x;
				xin[0][1 % CT][lane] = c0.y;
				xin[1 * kIn][0][lane] = c0.z;
				xin[1 * kIn][1 % CT][lane] = c0.w;
				xin[2 * kIn][0][lane] = c1.x;
				xin[2 * kIn][1 % CT][lane] = c1.y;
				xin[3 * kIn][0][lane] = c1.z;
				xin[3 * kIn][1 % CT][lane] = c1.w;
			}
			bar_sync<kBarEmpty>(b);
#pragma unroll (kDuoUnrollFront)
			for (int f = 0; f < count; ++f) {
				const int i = first + f;
				float x[CT], acc[CT];
				issue(i + kFwDepth);
				cp_async_wait_group<kFwDepth>();
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					x[c] = fast_io ? xin[OALSFX_DUO_FAST_IN ? f : 0][c][lane] : io_ok ? col[(i & (kFwSlots - 1)) * kFwSlotFloats + (4 + c) * kLanes] : 0.0F;
					acc[c] = 0.0F;
				}
				// direct send (oalsfxpp.cpp:2924-2950); gains sanitized by the host
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					pan_add<CT, true>(acc, CT, a.direct.gains[c], x[c]);
				}
				r0.step(a, 0, x, acc);
				r1.step(a, 1, x, acc);
				r2.step(a, 2, x, acc);
				r3in.step(a, 3, x, acc);
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					if (!split_reverb) {
						xch[b][f][c][lane] = x[c];
					}
					xch[b][f][kBusAt + c][lane] = acc[c];
				}
			}
			__threadfence_block();
			bar_arrive<kBarFull>(b);
		}
		cp_async_wait_group<0>();
		r0.end_state_only(a, 0, tile, lane);
		r1.end_state_only(a, 1, tile, lane);
		r2.end_state_only(a, 2, tile, lane);
		r3in.end_state_only(a, 3, tile, lane);
		store_passthrough_history<CT>(ss, 0, src, a, io_ok);
		if (!F0::kIsNull) {
			store_passthrough_history<CT>(ss, 1 + a.aux_index[0], src, a, io_ok);
		}
		if (!F1::kIsNull) {
			store_passthrough_history<CT>(ss, 1 + a.aux_index[1], src, a, io_ok);
		}
		if (!F2::kIsNull) {
			store_passthrough_history<CT>(ss, 1 + a.aux_index[2], src, a, io_ok);
		}
	} else {
		SlotRunner<CT, false, Back3> r3;
		r3.begin(a, 3, tile, lane, back_has_window ? window + lane : nullptr);
		bar_arrive<kBarEmpty>(0);
		bar_arrive<kBarEmpty>(1);
		for (int ci = 0; ci < chunks; ++ci) {
			const int b = ci & 1;
			const int first = ci * kDuoChunk;
			const int count = min(kDuoChunk, a.frames - first);
			bar_sync<kBarFull>(b);
#pragma unroll (kDuoUnrollBack)
			for (int f = 0; f < count; ++f) {
				const int i = first + f;
				float x[CT], acc[CT];
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					x[c] = split_reverb ? 0.0F : xch[b][f][c][lane];
					acc[c] = xch[b][f][kBusAt + c][lane];
				}
				if (split_reverb) {
					r3.fx.template step<CT, true>(a.slot[3], x, acc, CT); // the tail does not read the wet bus
				} else {
					r3.step(a, 3, x, acc);
				}
				if (fast_out) {
#pragma unroll
					for (int c = 0; c < CT; ++c) {
						xch[b][f][kBusAt + c][lane] = acc[c]; // the finished frame replaces the partial bus (own column)
					}
				} else if (io_ok) {
#pragma unroll
					for (int c = 0; c < CT; ++c) {
						dst[i * a.io_fs + c * a.io_cs] = acc[c];
					}
				}
			}
			if (fast_out && io_ok) {
				float4* row = reinterpret_cast<float4*>(dst + first * CT);
#pragma unroll
				for (int g = 0; g < kDuoChunk / 4; ++g) {
					const int f0 = 4 * g;
					__stcs(row + 2 * g, make_float4(xch[b][f0][kBusAt][lane], xch[b][f0][kBusAt + 1 % CT][lane],
						xch[b][f0 + 1][kBusAt][lane], xch[b][f0 + 1][kBusAt + 1 % CT][lane]));
					__stcs(row + 2 * g + 1, make_float4(xch[b][f0 + 2][kBusAt][lane], xch[b][f0 + 2][kBusAt + 1 % CT][lane],
						xch[b][f0 + 3][kBusAt][lane], xch[b][f0 + 3][kBusAt + 1 % CT][lane]));
				}
			}
			if (ci + 2 < chunks) {
				bar_arrive<kBarEmpty>(b); // nobody waits for the last two drains
			}
		}
		r3.end_state_only(a, 3, tile, lane);
		if (!F3::kIsNull) {
			store_passthrough_history<CT>(ss, 1 + a.aux_index[3], src, a, io_ok);
		}
	}
}

// One parameter class for the whole launch: the coefficient blocks are kernel arguments (constant bank).
template <int CT, class F0, class F1, class F2, class F3>
#ifdef OALSFX_DUO_MAXNREG
__global__ void __maxnreg__(OALSFX_DUO_MAXNREG) duo_kernel(
#else
__global__ void __launch_bounds__(64, OALSFX_DUO_MIN_CTAS) duo_kernel(
#endif
	const __grid_constant__ MixArgs a)
{
	duo_body<CT, F0, F1, F2, F3>(a);
}

// One parameter class PER TILE (every stream block of 32 its own presets): the launch-wide part of the arguments
// comes from the kernel arguments, the tile's coefficient blocks (direct / aux sends, four slots) and its pending-
// update bits from a class table in HBM (MixArgs::class_table, indexed by tile_class[tile]); both are assembled
// in shared memory and the same body runs from there.  Coefficients then cost a shared-memory load instead of
// being constant-bank operands -- but the tile still runs the fused, prefetching pipeline in one launch for all
// classes, where the alternatives are one launch per class or per-lane coefficient fetches (table mode).
template <int CT, class F0, class F1, class F2, class F3>
__global__ void __launch_bounds__(64, OALSFX_DUO_MIN_CTAS) duo_multi_kernel(const __grid_constant__ MixArgs a)
{
	__shared__ __align__(16) MixArgs sa;
	const int tile = a.tiles ? static_cast<int>(a.tiles[blockIdx.x].tile) : a.tile_first + static_cast<int>(blockIdx.x);
	const MixClassEntry* entry = a.class_table + a.tile_class[tile];
	const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
	uint32_t* ps = reinterpret_cast<uint32_t*>(&sa);
	constexpr int kWords = static_cast<int>(sizeof(MixArgs) / 4), kCoef0 = static_cast<int>(kMixCoefOffset / 4),
		kCoefWords = static_cast<int>(kMixCoefBytes / 4);
	for (int i = threadIdx.x; i < kWords; i += 64) {
		if (i < kCoef0 || i >= kCoef0 + kCoefWords) {
			ps[i] = pa[i];
		}
	}
	for (int i = threadIdx.x; i < kCoefWords; i += 64) {
		ps[kCoef0 + i] = entry->coefs[i];
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		sa.update_mask = a.update_mask & entry->pending; // a.update_mask: all ones on the first block of a mix call
	}
	__syncthreads();
	duo_body<CT, F0, F1, F2, F3>(sa);
}

} // namespace duo
} // namespace oalsfx

#endif // __CUDACC__
#endif
