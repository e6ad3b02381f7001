// quartet.cuh -- the fused 4-slot kernel as a FOUR-stage warp pipeline per 32-stream tile.
//
// Why: the per-sample work of a stream is one long dependent instruction sequence with little
// instruction-level parallelism, and its recurrent state has to stay in registers, so a streaming
// multiprocessor only holds a handful of such warps (profiles/: ~0.3 instructions per cycle per warp,
// 8-12 warps per SM with the two-stage duo kernel).  Cutting the sequence into four stages, one warp
// each, doubles the warps that work on a tile at any time without adding registers per SM:
//
//   A : input frame, dry mix (direct send), slot 0 + the reverb's input stage  (cfg4: equalizer; B->A + shelves + main-line feed)
//   B : slots 1 and 2                                                          (cfg4: chorus, echo)
//   C : reverb, early half  (early taps, early all-pass, early lines, late feed, pan of lines 0..3)
//   D : reverb, late half   (late taps, T60 filters, late all-pass, late lines, pan of lines 4..7) + output rows
//
// A stage hands the running output bus (and, A -> B, the input frame) to the next one through a
// double-buffered shared-memory exchange of kQuartetChunk frames guarded by named barriers, so the
// per-sample summation order stays dry, slot 0, 1, 2, 3(early lines, late lines) -- the reference's
// (oalsfxpp.cpp:2984-3037, 6142-6166).  Stages that share a delay line only ever pass data forward in
// time (B feeds the main line that C and D read, C feeds the part of it that D reads), the later stage
// runs behind the earlier one, and every hand-off is a CTA-scope fence + barrier; the main ring keeps
// 256 spare frames beyond its longest tap (oalsfxpp.cpp:6573), far more than the <= 8 frames per
// hand-off the stages may be apart.
//
// Host-checked requirements (as for the duo kernels): no send shelf filter active, frames >= 2, every
// tile of the launch takes part with all of its lanes, sanitized static gains.
#ifndef OALSFX_QUARTET_CUH
#define OALSFX_QUARTET_CUH

#if defined(__CUDACC__)

#include "duo.cuh"

namespace oalsfx {
namespace quartet {

#ifndef OALSFX_QUARTET_MIN_CTAS
#define OALSFX_QUARTET_MIN_CTAS 4      // resident CTAs per SM the register allocation is sized for (128 regs)
#endif

#ifndef OALSFX_QUARTET_CHUNK
#define OALSFX_QUARTET_CHUNK 4
#endif
constexpr int kQuartetChunk = OALSFX_QUARTET_CHUNK; // frames per hand-off (a multiple of the 4-frame output row batch)
static_assert(kQuartetChunk % 4 == 0, "output rows are written four frames at a time");
constexpr int kThreads = 4 * kLanes;

// Named barriers of hand-off H (0: A->B, 1: B->C, 2: C->D): full = 4H + buffer, empty = 4H + 2 + buffer.
// Each is used by exactly two warps (64 threads): one arrives, the other waits.
template <int H> __device__ __forceinline__ void wait_full(int b) { duo::bar_sync<4 * H>(b); }
template <int H> __device__ __forceinline__ void signal_full(int b) { duo::bar_arrive<4 * H>(b); }
template <int H> __device__ __forceinline__ void wait_empty(int b) { duo::bar_sync<4 * H + 2>(b); }
template <int H> __device__ __forceinline__ void signal_empty(int b) { duo::bar_arrive<4 * H + 2>(b); }

template <int CT, class F0, class F1, class F2, class F3>
__global__ void __launch_bounds__(kThreads, OALSFX_QUARTET_MIN_CTAS) quartet_kernel(const __grid_constant__ MixArgs a)
{
	// RP: position of the reverb the pipeline is built around -- the LAST non-null slot, so that the stages'
	// order (slots before it in A and B, its halves in C and D) is also the reference's summation order.
	constexpr int RP = std::is_same<F3, FxReverb>::value ? 3 : std::is_same<F2, FxReverb>::value ? 2 :
		std::is_same<F1, FxReverb>::value ? 1 : std::is_same<F0, FxReverb>::value ? 0 : -1;
	static_assert(RP >= 0, "the quartet pipeline splits a reverb");
	static_assert((RP >= 3 || F3::kIsNull) && (RP >= 2 || F2::kIsNull) && (RP >= 1 || F1::kIsNull), "slots after the reverb must be empty");
	static_assert(CT == 1 || CT == 2, "output rows are batched for mono and stereo");
	using A0 = typename std::conditional<(RP > 0), F0, FxNull>::type;   // stage A: slot 0 unless it is the reverb
	using B1 = typename std::conditional<(RP > 1), F1, FxNull>::type;   // stage B: slots 1 and 2 if they precede it
	using B2 = typename std::conditional<(RP > 2), F2, FxNull>::type;
	__shared__ __align__(16) float win_c[FxReverbEarly::kWindowFloats]; // stage C: 12 early taps x 8 ring positions
	__shared__ __align__(16) float win_d[FxReverbLate::kWindowFloats];  // stage D: 12 late taps x 8 ring positions
	__shared__ float xab[2][kQuartetChunk][2 * CT][kLanes];              // A -> B: x_0..x_C-1, bus_0..bus_C-1
	__shared__ float xbc[2][kQuartetChunk][CT][kLanes];                  // B -> C: bus
	__shared__ float xcd[2][kQuartetChunk][CT][kLanes];                  // C -> D: bus; D parks the finished frames here
	__shared__ float win_a[kFwSlots][CT][kLanes];                        // stage A: input frames in flight
	__shared__ float win_b[kFwWarpFloats];                               // stage B: chorus / echo taps in flight

	const int tile = a.tiles ? static_cast<int>(a.tiles[blockIdx.x].tile) : a.tile_first + static_cast<int>(blockIdx.x);
	const int lane = threadIdx.x % kLanes;
	// Warp w of every CTA lands on scheduler w.  OALSFX_QUARTET_ROTATE = 1 rotates the stages so each scheduler
	// sees all four (even load whatever the stage sizes); 0 keeps one stage per scheduler (one hot loop per
	// instruction cache).
#ifndef OALSFX_QUARTET_ROTATE
#define OALSFX_QUARTET_ROTATE 1
#endif
	const int stage = (static_cast<int>(threadIdx.x / kLanes) + (OALSFX_QUARTET_ROTATE ? static_cast<int>(blockIdx.x) : 0)) & 3;
	const bool io_ok = tile * kLanes + lane < a.num_streams;
	const float* src = a.src + tile * a.io_ts + lane * a.io_ls;
	float* dst = a.dst + tile * a.io_ts + lane * a.io_ls;
	uint32_t* ss = a.send_state + (static_cast<long long>(tile) * kSendStateWords) * kLanes + lane;
	const int chunks = (a.frames + kQuartetChunk - 1) / kQuartetChunk;
	// Output rows: with interleaved frames a thread's 4 frames x CT channels are 16 or 32 contiguous bytes of its row.
	const bool fast_out = a.io_cs == 1 && a.io_fs == CT && (a.frames % kQuartetChunk) == 0 && (a.io_ls % 4) == 0 &&
		(a.io_ts % 4) == 0 && (reinterpret_cast<unsigned long long>(a.dst) & 15ULL) == 0;

	if (stage == 0) {
		// ---- A: input, dry, slot 0 ----
		SlotRunner<CT, false, A0> r0;
		SlotRunner<CT, false, FxReverbInput> r3in; // adds nothing to the bus, so it may run ahead of slots 1 and 2
		r0.begin(a, 0, tile, lane, nullptr);
		r3in.begin(a, RP, tile, lane, nullptr);
		const unsigned win_s = smem_addr(&win_a[0][0][lane]);
		const float* in = src;
		auto issue_input = [&](int frame) {
			if (io_ok && frame < a.frames) {
				const unsigned slot = win_s + static_cast<unsigned>((frame & (kFwSlots - 1)) * CT * kLanes * 4);
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					cp_async_f32_s(slot + c * kLanes * 4, in + c * a.io_cs);
				}
			}
			in += a.io_fs;
			cp_async_commit_group();
		};
		for (int k = 0; k < kFwDepth; ++k) {
			issue_input(k);
		}
		for (int ci = 0; ci < chunks; ++ci) {
			const int b = ci & 1;
			const int first = ci * kQuartetChunk;
			const int count = min(kQuartetChunk, a.frames - first);
			wait_empty<0>(b);
			for (int f = 0; f < count; ++f) {
				const int i = first + f;
				float x[CT], acc[CT];
				issue_input(i + kFwDepth);
				cp_async_wait_group<kFwDepth>();
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					x[c] = io_ok ? win_a[i & (kFwSlots - 1)][c][lane] : 0.0F;
					acc[c] = 0.0F;
				}
				// direct send (oalsfxpp.cpp:2924-2950); gains sanitized by the host
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					pan_add<CT, true>(acc, CT, a.direct.gains[c], x[c]);
				}
				r0.step(a, 0, x, acc);
				r3in.step(a, RP, x, acc);
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					xab[b][f][c][lane] = x[c];
					xab[b][f][CT + c][lane] = acc[c];
				}
			}
			__threadfence_block();
			signal_full<0>(b);
		}
		cp_async_wait_group<0>();
		r0.end_state_only(a, 0, tile, lane);
		r3in.end_state_only(a, RP, tile, lane);
		duo::store_passthrough_history<CT>(ss, 0, src, a, io_ok);
		if (!A0::kIsNull) {
			duo::store_passthrough_history<CT>(ss, 1 + a.aux_index[0], src, a, io_ok);
		}
		duo::store_passthrough_history<CT>(ss, 1 + a.aux_index[RP], src, a, io_ok);
	} else if (stage == 1) {
		// ---- B: slots 1, 2 and the reverb's input stage ----
		SlotRunner<CT, false, B1> r1;
		SlotRunner<CT, false, B2> r2;
		// Window taps: 0,1 = the first chorus/flanger, 2,3 = the first echo.
		constexpr bool m1 = std::is_same<B1, FxModDelay>::value, m2 = std::is_same<B2, FxModDelay>::value && !m1;
		constexpr bool e1 = std::is_same<B1, FxEcho>::value, e2 = std::is_same<B2, FxEcho>::value && !e1;
		float* col = win_b + lane;
		r1.begin(a, 1, tile, lane, m1 ? col : e1 ? col + 2 * kLanes : nullptr);
		r2.begin(a, 2, tile, lane, m2 ? col : e2 ? col + 2 * kLanes : nullptr);
		for (int k = 0; k < kFwDepth; ++k) {
			r1.fx.prefetch_issue(a.slot[1], k);
			r2.fx.prefetch_issue(a.slot[2], k);
			cp_async_commit_group();
		}
		signal_empty<0>(0);
		signal_empty<0>(1);
		for (int ci = 0; ci < chunks; ++ci) {
			const int b = ci & 1;
			const int first = ci * kQuartetChunk;
			const int count = min(kQuartetChunk, a.frames - first);
			wait_full<0>(b);
			wait_empty<1>(b);
			for (int f = 0; f < count; ++f) {
				float x[CT], acc[CT];
				r1.fx.prefetch_next(a.slot[1]);
				r2.fx.prefetch_next(a.slot[2]);
				cp_async_commit_group();
				cp_async_wait_group<kFwDepth>();
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					x[c] = xab[b][f][c][lane];
					acc[c] = xab[b][f][CT + c][lane];
				}
				r1.step(a, 1, x, acc);
				r2.step(a, 2, x, acc);
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					xbc[b][f][c][lane] = acc[c];
				}
			}
			__threadfence_block();
			signal_full<1>(b);
			if (ci + 2 < chunks) {
				signal_empty<0>(b); // nobody waits for the last two drains
			}
		}
		cp_async_wait_group<0>();
		r1.end_state_only(a, 1, tile, lane);
		r2.end_state_only(a, 2, tile, lane);
		if (!B1::kIsNull) {
			duo::store_passthrough_history<CT>(ss, 1 + a.aux_index[1], src, a, io_ok);
		}
		if (!B2::kIsNull) {
			duo::store_passthrough_history<CT>(ss, 1 + a.aux_index[2], src, a, io_ok);
		}
	} else if (stage == 2) {
		// ---- C: reverb, early half ----
		SlotRunner<CT, false, FxReverbEarly> r3;
		r3.begin(a, RP, tile, lane, win_c + lane);
		signal_empty<1>(0);
		signal_empty<1>(1);
		const float none[kWetChannels] = {0.0F, 0.0F, 0.0F, 0.0F};
		for (int ci = 0; ci < chunks; ++ci) {
			const int b = ci & 1;
			const int first = ci * kQuartetChunk;
			const int count = min(kQuartetChunk, a.frames - first);
			wait_full<1>(b);
			wait_empty<2>(b);
			for (int f = 0; f < count; ++f) {
				float acc[CT];
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					acc[c] = xbc[b][f][c][lane];
				}
				r3.fx.template step<CT, true>(a.slot[RP], none, acc, CT); // the halves do not read the wet bus
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					xcd[b][f][c][lane] = acc[c];
				}
			}
			__threadfence_block();
			signal_full<2>(b);
			if (ci + 2 < chunks) {
				signal_empty<1>(b);
			}
		}
		r3.end_state_only(a, RP, tile, lane);
	} else {
		// ---- D: reverb, late half, output ----
		SlotRunner<CT, false, FxReverbLate> r3;
		r3.begin(a, RP, tile, lane, win_d + lane);
		signal_empty<2>(0);
		signal_empty<2>(1);
		const float none[kWetChannels] = {0.0F, 0.0F, 0.0F, 0.0F};
		for (int ci = 0; ci < chunks; ++ci) {
			const int b = ci & 1;
			const int first = ci * kQuartetChunk;
			const int count = min(kQuartetChunk, a.frames - first);
			wait_full<2>(b);
			for (int f = 0; f < count; ++f) {
				const int i = first + f;
				float acc[CT];
#pragma unroll
				for (int c = 0; c < CT; ++c) {
					acc[c] = xcd[b][f][c][lane];
				}
				r3.fx.template step<CT, true>(a.slot[RP], none, acc, CT);
				if (fast_out) {
#pragma unroll
					for (int c = 0; c < CT; ++c) {
						xcd[b][f][c][lane] = acc[c]; // the finished frame replaces the partial bus (own column)
					}
				} else if (io_ok) {
#pragma unroll
					for (int c = 0; c < CT; ++c) {
						dst[i * a.io_fs + c * a.io_cs] = acc[c];
					}
				}
			}
			if (fast_out && io_ok) {
				// kQuartetChunk frames x CT channels of this thread's row, 16 bytes at a time
				float4* row = reinterpret_cast<float4*>(dst + first * CT);
#pragma unroll
				for (int g = 0; g < kQuartetChunk * CT / 4; ++g) {
					const int e0 = 4 * g; // element index within the chunk: frame = e / CT, channel = e % CT
					__stcs(row + g, make_float4(xcd[b][(e0 + 0) / CT][(e0 + 0) % CT][lane], xcd[b][(e0 + 1) / CT][(e0 + 1) % CT][lane],
						xcd[b][(e0 + 2) / CT][(e0 + 2) % CT][lane], xcd[b][(e0 + 3) / CT][(e0 + 3) % CT][lane]));
				}
			}
			if (ci + 2 < chunks) {
				signal_empty<2>(b);
			}
		}
		r3.end_state_only(a, RP, tile, lane);
	}
}

} // namespace quartet
} // namespace oalsfx

#endif // __CUDACC__
#endif
