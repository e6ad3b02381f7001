// scan.cuh -- biquad cascades as a CHUNKED LINEAR-RECURRENCE SCAN over time (few streams, long blocks).
//
// A stream's equalizer is four cascaded biquads on three wet channels (oalsfxpp.cpp:5161-5213, FilterState::process
// :984-1036): per sample a chain of dependent operations, so with a thread per stream a block of F frames takes F
// times that chain however few streams there are.  A biquad is a LINEAR recurrence,
//
//     s[t] = A s[t-1] + (ff[t], 0),   s = (y[t], y[t-1]),   A = [[-a1, -a2], [1, 0]],   ff = b0 x[t] + b1 x[t-1] + b2 x[t-2]
//
// so a warp can run one (stream, wet channel) with its 32 lanes spread over TIME: lane i owns the chunk of L = F / 32
// consecutive frames [iL, (i+1)L) and
//
//   1. runs the recurrence over its chunk from the zero state (fp64): end state z_i
//   2. the chunks are stitched by a Kogge-Stone scan over the lanes -- five shuffle steps s_i += (A^L)^(2^k) s_(i - 2^k),
//      A^L from L steps of the homogeneous recurrence, in fp64 -- which gives every lane the true state at the start of
//      its chunk
//   3. runs the recurrence over its chunk again from that state, in fp32 and in the reference's expression order: the
//      outputs, which are the next band's inputs
//
// i.e. two passes of L steps instead of one of F: 16x less dependent work per block.  The scan RE-ASSOCIATES: every chunk
// starts from a state that is the exactly rounded true state instead of the reference's own (noisy) one, so the result
// is as close to the exact-arithmetic filter as the reference's is, and the two differ by the reference's rounding
// noise -- ~1e-5 for the default bands (a 200 Hz shelf at 48 kHz amplifies every rounding by ~100), more with higher Q.
// Hence opt-in (OALSFX_SCAN=1); the bit-exact kernels stay the default.
//
// CTA = 3 warps = one stream: wet channels 0, 1, 3 (channel 2 is dead, FxEqualizer::kDeadWet); then all 96 threads pan the
// three cascade outputs onto the bus in the reference's order and write the block.
#ifndef OALSFX_SCAN_CUH
#define OALSFX_SCAN_CUH

#if defined(__CUDACC__)

#include "mix.cuh"

namespace oalsfx {
namespace scan {

constexpr int kWarps = 3;
constexpr int kThreads = kWarps * kLanes;
constexpr int kMaxChunk = kMaxBlockFrames / kLanes;   // frames per lane (64)
constexpr int kMinFrames = 2 * kLanes;                // every active lane but the last owns at least two frames

struct M2 { double a, b, c, d; };                      // [[a, b], [c, d]]
__device__ __forceinline__ M2 mul(const M2& x, const M2& y)
{
	return M2{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}

template <int CT>
__global__ void __launch_bounds__(kThreads) scan_equalizer_kernel(const __grid_constant__ MixArgs a)
{
	__shared__ float ybuf[kWarps][kMaxBlockFrames];
	const int stream = static_cast<int>(blockIdx.x);
	const int tile = stream / kLanes, slane = stream % kLanes;
	const int lane = threadIdx.x % kLanes, w = threadIdx.x / kLanes;
	const int wet_channel = w == 2 ? 3 : w;
	const EqualizerCoef& c = a.slot[0].u.equalizer;
	const SendCoef& send = a.aux[0];
	const float* src = a.src + tile * a.io_ts + slane * a.io_ls;
	float* dst = a.dst + tile * a.io_ts + slane * a.io_ls;
	uint32_t* st = a.slot_state[0] + (static_cast<long long>(tile) * kSlotStateWords) * kLanes + slane;
	const int F = a.frames;
	const int L = (F + kLanes - 1) / kLanes;             // frames per lane
	const int first = lane * L;
	const int n = min(L, F - first);                     // this lane's frames (<= 0: none)
	const int last_lane = (F - 1) / L;

	// the wet channel's input over this lane's chunk (SlotRunner::step without shelf filters: from 0, channel by channel)
	float v[kMaxChunk];
#pragma unroll
	for (int t = 0; t < kMaxChunk; ++t) {
		float wet = 0.0F;
		if (t < n) {
#pragma unroll
			for (int ch = 0; ch < CT; ++ch) {
				wet += src[(first + t) * a.io_fs + ch * a.io_cs] * send.gains[ch][wet_channel];
			}
		}
		v[t] = wet;
	}

#pragma unroll 1
	for (int b = 0; b < 4; ++b) {
		const Biquad q = c.band[b];
		BiquadHist h;
		load_words(h, st + ((b * 4 + wet_channel) * 4) * kLanes);
		// input history at the start of the chunk: the previous lane's last two inputs (every lane before the last owns
		// L >= 2 frames), the stored history for lane 0
		float xm1 = __shfl_up_sync(0xFFFFFFFFU, v[0], 1), xm2 = xm1;
		{
			float last1 = 0.0F, last2 = 0.0F;
#pragma unroll
			for (int t = 0; t < kMaxChunk; ++t) {
				if (t == L - 1) {
					last1 = v[t];
				}
				if (t == L - 2) {
					last2 = v[t];
				}
			}
			xm1 = __shfl_up_sync(0xFFFFFFFFU, last1, 1);
			xm2 = __shfl_up_sync(0xFFFFFFFFU, last2, 1);
			if (lane == 0) {
				xm1 = h.x0;
				xm2 = h.x1;
			}
		}
		// 1. zero-state response of the chunk, in fp64: with poles near the unit circle the zero-state response is a large
		//    transient that the homogeneous part (step 2) cancels again -- in fp32 that cancellation costs ten times the
		//    reference's own rounding noise (measured on a 200 Hz shelf: 1.3e-4 against 1.1e-5 from the exact filter)
		double zy0 = 0.0, zy1 = 0.0;
		{
			double x0 = xm1, x1 = xm2;
			const double b0 = q.b0, b1 = q.b1, b2 = q.b2, a1 = q.a1, a2 = q.a2;
#pragma unroll
			for (int t = 0; t < kMaxChunk; ++t) {
				if (t < n) {
					const double xv = v[t];
					const double y = (b0 * xv) + (b1 * x0) + (b2 * x1) - (a1 * zy0) - (a2 * zy1);
					x1 = x0;
					x0 = xv;
					zy1 = zy0;
					zy0 = y;
				}
			}
		}
		// 2. stitch the chunks: s_end[i] = A^L s_end[i-1] + z[i] for full chunks, the block's incoming state before lane 0
		M2 p;                                            // A^L
		{
			double e0 = 1.0, e1 = 0.0, g0 = 0.0, g1 = 1.0;  // columns: the recurrence run on the basis vectors
			for (int t = 0; t < L; ++t) {
				const double ne = -static_cast<double>(q.a1) * e0 - static_cast<double>(q.a2) * e1;
				e1 = e0;
				e0 = ne;
				const double ng = -static_cast<double>(q.a1) * g0 - static_cast<double>(q.a2) * g1;
				g1 = g0;
				g0 = ng;
			}
			p = M2{e0, g0, e1, g1};
		}
		double s0 = zy0, s1 = zy1;
		if (lane == 0) {
			s0 += p.a * h.y0 + p.b * h.y1;
			s1 += p.c * h.y0 + p.d * h.y1;
		}
#pragma unroll
		for (int k = 1; k < kLanes; k <<= 1) {
			const double u0 = __shfl_up_sync(0xFFFFFFFFU, s0, k), u1 = __shfl_up_sync(0xFFFFFFFFU, s1, k);
			if (lane >= k) {
				s0 += p.a * u0 + p.b * u1;
				s1 += p.c * u0 + p.d * u1;
			}
			p = mul(p, p);
		}
		float y0 = static_cast<float>(__shfl_up_sync(0xFFFFFFFFU, s0, 1)), y1 = static_cast<float>(__shfl_up_sync(0xFFFFFFFFU, s1, 1));
		if (lane == 0) {
			y0 = h.y0;
			y1 = h.y1;
		}
		// 3. the chunk again from its true state: outputs replace inputs
		{
			float x0 = xm1, x1 = xm2;
#pragma unroll
			for (int t = 0; t < kMaxChunk; ++t) {
				if (t < n) {
					const float y = (q.b0 * v[t]) + (q.b1 * x0) + (q.b2 * x1) - (q.a1 * y0) - (q.a2 * y1);
					x1 = x0;
					x0 = v[t];
					y1 = y0;
					y0 = y;
					v[t] = y;
				}
			}
			if (lane == last_lane) {                     // the band's history after the block
				const BiquadHist out = {x0, x1, y0, y1};
				store_words(out, st + ((b * 4 + wet_channel) * 4) * kLanes);
			}
		}
	}
#pragma unroll
	for (int t = 0; t < kMaxChunk; ++t) {
		if (t < n) {
			ybuf[w][first + t] = v[t];
		}
	}
	__syncthreads();

	// dry mix + the three cascade outputs, added in the reference's order (mix_stream / FxEqualizer::step)
	for (int t = threadIdx.x; t < F; t += kThreads) {
		float x[CT], acc[CT];
#pragma unroll
		for (int ch = 0; ch < CT; ++ch) {
			x[ch] = src[t * a.io_fs + ch * a.io_cs];
			acc[ch] = 0.0F;
		}
#pragma unroll
		for (int ch = 0; ch < CT; ++ch) {
#pragma unroll
			for (int k = 0; k < CT; ++k) {
				acc[k] += x[ch] * a.direct.gains[ch][k];
			}
		}
#pragma unroll
		for (int k = 0; k < CT; ++k) {
			acc[k] += ybuf[0][t] * c.gains[0][k];
		}
#pragma unroll
		for (int k = 0; k < CT; ++k) {
			acc[k] += ybuf[1][t] * c.gains[1][k];
		}
#pragma unroll
		for (int k = 0; k < CT; ++k) {
			acc[k] += ybuf[2][t] * c.gains[3][k];
		}
#pragma unroll
		for (int ch = 0; ch < CT; ++ch) {
			dst[t * a.io_fs + ch * a.io_cs] = acc[ch];
		}
	}
	// send filter histories without shelf filters: the last two input samples (oalsfxpp.cpp:1038-1056)
	if (threadIdx.x < 2) {
		uint32_t* ss = a.send_state + (static_cast<long long>(tile) * kSendStateWords) * kLanes + slane;
		const int sendi = threadIdx.x == 0 ? 0 : 1 + a.aux_index[0];
#pragma unroll
		for (int ch = 0; ch < CT; ++ch) {
			const float l1 = src[(F - 1) * a.io_fs + ch * a.io_cs], l2 = src[(F - 2) * a.io_fs + ch * a.io_cs];
			SendHist hh;
			hh.lp.x0 = hh.lp.y0 = hh.hp.x0 = hh.hp.y0 = l1;
			hh.lp.x1 = hh.lp.y1 = hh.hp.x1 = hh.hp.y1 = l2;
			store_words(hh, ss + (sendi * kMaxChannels + ch) * 8 * kLanes);
		}
	}
}

} // namespace scan
} // namespace oalsfx

#endif // __CUDACC__
#endif
