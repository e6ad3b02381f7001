// stream_pattern.cu -- what HBM bandwidth does the delay-line access pattern allow?
// Each warp owns one "tile" with NR read rings and NW write rings (each ring = LINES x 128 B, lane-interleaved
// like the engine's arenas).  Per visit the warp reads G consecutive 128-byte lines from every read ring
// and writes G consecutive lines to every write ring with ONE vector access per lane (4*G bytes: the warp
// request is G*128 contiguous bytes).  G = 1 is the sample-by-sample pattern of the mix kernels; larger G is
// what staging G samples per ring visit produces.  mode: 0 = read+write, 1 = read only, 2 = write only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_pattern stream_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int NR = 24, NW = 24;
template <int G> struct Vec { float v[G]; };
template <int G> struct __align__(4 * G) AVec { float v[G]; };

template <int G, int U>  // U = independent visits in flight per warp (unroll)
__global__ void __launch_bounds__(64) pattern(const float* __restrict__ rd, float* __restrict__ wr, long long ring_floats, int lines,
	int visits, int start, int mode, float* sink)
{
	using V = AVec<G>;
	const int lane = threadIdx.x & 31;
	const long long tile = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
	const V* r = reinterpret_cast<const V*>(rd + tile * NR * ring_floats) + lane;
	V* w = reinterpret_cast<V*>(wr + tile * NW * ring_floats) + lane;
	const long long ring_v = ring_floats / G;
	const int vlines = lines / G;
	float acc = 0.f;
	for (int v = 0; v < visits; v += U) {
		V val[U][NR];
		if (mode != 2) {
#pragma unroll
			for (int u = 0; u < U; ++u)
#pragma unroll
				for (int k = 0; k < NR; ++k) {
					const int pos = (start + v + u + k * 977) & (vlines - 1);
					val[u][k] = r[k * ring_v + (long long)pos * 32];
				}
#pragma unroll
			for (int u = 0; u < U; ++u)
#pragma unroll
				for (int k = 0; k < NR; ++k)
#pragma unroll
					for (int g = 0; g < G; ++g) acc += val[u][k].v[g];
		}
		if (mode != 1) {
#pragma unroll
			for (int u = 0; u < U; ++u)
#pragma unroll
				for (int k = 0; k < NW; ++k) {
					const int pos = (start + v + u + k * 1409) & (vlines - 1);
					V o;
#pragma unroll
					for (int g = 0; g < G; ++g) o.v[g] = acc + k + g;
					w[k * ring_v + (long long)pos * 32] = o;
				}
		}
	}
	if (acc == 12345.678f) *sink = acc;
}

// Mixed granularity: every "sample group" of 4 positions reads NR rings with GR-line accesses and writes NW rings
// with GW-line accesses (GR, GW in {1, 4}): the shape of "batched reads, per-sample writes" vs "both batched".
template <int GR, int GW>
__global__ void __launch_bounds__(64) mixed(const float* __restrict__ rd, float* __restrict__ wr, long long ring_floats, int lines,
	int groups, int start, float* sink)
{
	const int lane = threadIdx.x & 31;
	const long long tile = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
	const float* r = rd + tile * NR * ring_floats;
	float* w = wr + tile * NW * ring_floats;
	float acc = 0.f;
	for (int g = 0; g < groups; ++g) {
		const int p4 = ((start + g) * 4);
		if (GR == 4) {
			float4 v[NR];
#pragma unroll
			for (int k = 0; k < NR; ++k) {
				const int pos = (p4 + k * 3908) & (lines - 1);
				v[k] = *reinterpret_cast<const float4*>(r + k * ring_floats + (long long)pos * 32 + lane * 4);
			}
#pragma unroll
			for (int k = 0; k < NR; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
		} else {
#pragma unroll
			for (int s = 0; s < 4; ++s) {
				float v[NR];
#pragma unroll
				for (int k = 0; k < NR; ++k) {
					const int pos = (p4 + s + k * 3908) & (lines - 1);
					v[k] = r[k * ring_floats + (long long)pos * 32 + lane];
				}
#pragma unroll
				for (int k = 0; k < NR; ++k) acc += v[k];
			}
		}
		if (GW == 4) {
#pragma unroll
			for (int k = 0; k < NW; ++k) {
				const int pos = (p4 + k * 5636) & (lines - 1);
				*reinterpret_cast<float4*>(w + k * ring_floats + (long long)pos * 32 + lane * 4) = make_float4(acc, acc + k, acc, acc);
			}
		} else {
#pragma unroll
			for (int s = 0; s < 4; ++s)
#pragma unroll
				for (int k = 0; k < NW; ++k) {
					const int pos = (p4 + s + k * 5636) & (lines - 1);
					w[k * ring_floats + (long long)pos * 32 + lane] = acc + k + s;
				}
		}
	}
	if (acc == 12345.678f) *sink = acc;
}

template <int GR, int GW> void run_mixed(const float* rd, float* wr, long long ring_floats, int lines, int tiles, float* sink)
{
	const int samples = 2048, groups = samples / 4;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e9f;
	for (int rep = 0; rep < 4; ++rep) {
		cudaEventRecord(e0);
		mixed<GR, GW><<<tiles / 2, 64>>>(rd, wr, ring_floats, lines, groups, rep * groups, sink);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (rep && ms < best) best = ms;
	}
	const double bytes = double(tiles) * samples * (NR + NW) * 128.0;
	printf("  reads %d B/visit, writes %d B/visit: %.3f ms, %5.0f GB/s [%s]\n", GR * 128, GW * 128, best, bytes / best * 1e-6,
		cudaGetErrorString(cudaGetLastError()));
}

template <int G, int U> void run(const float* rd, float* wr, long long ring_floats, int lines, int tiles, int mode, float* sink)
{
	const int samples = 2048, visits = samples / G;
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e9f;
	for (int rep = 0; rep < 4; ++rep) {
		cudaEventRecord(e0);
		pattern<G, U><<<tiles / 2, 64>>>(rd, wr, ring_floats, lines, visits, rep * visits, mode, sink);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (rep && ms < best) best = ms;
	}
	const double bytes = double(tiles) * samples * ((mode != 2 ? NR : 0) + (mode != 1 ? NW : 0)) * 128.0;
	printf("  G=%2d (%4d B/visit, %d visits in flight): %.3f ms, %5.0f GB/s [%s]\n", G, G * 128, U, best, bytes / best * 1e-6,
		cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv)
{
	const int tiles = argc > 1 ? atoi(argv[1]) : 2048;
	const int lines = 4096;                       // 512 KiB per ring
	const long long ring_floats = (long long)lines * 32;
	float *rd, *wr, *sink;
	cudaMalloc(&rd, sizeof(float) * ring_floats * NR * tiles);
	cudaMalloc(&wr, sizeof(float) * ring_floats * NW * tiles);
	cudaMalloc(&sink, 4);
	cudaMemset(rd, 0, sizeof(float) * ring_floats * NR * tiles);
	cudaMemset(wr, 0, sizeof(float) * ring_floats * NW * tiles);
	printf("tiles %d, footprint %.1f GB\n", tiles, sizeof(float) * ring_floats * (NR + NW) * tiles * 1e-9);
	printf("mixed granularity (read+write)\n");
	run_mixed<1, 1>(rd, wr, ring_floats, lines, tiles, sink);
	run_mixed<4, 1>(rd, wr, ring_floats, lines, tiles, sink);
	run_mixed<1, 4>(rd, wr, ring_floats, lines, tiles, sink);
	run_mixed<4, 4>(rd, wr, ring_floats, lines, tiles, sink);
	const char* names[3] = {"read+write", "read only", "write only"};
	for (int mode = 0; mode < 3; ++mode) {
		printf("%s\n", names[mode]);
		run<1, 4>(rd, wr, ring_floats, lines, tiles, mode, sink);
		run<1, 2>(rd, wr, ring_floats, lines, tiles, mode, sink);
		run<2, 2>(rd, wr, ring_floats, lines, tiles, mode, sink);
		run<4, 1>(rd, wr, ring_floats, lines, tiles, mode, sink);
		run<4, 2>(rd, wr, ring_floats, lines, tiles, mode, sink);
	}
	return 0;
}
