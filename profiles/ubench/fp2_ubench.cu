// fp2_ubench.cu -- issue rate and dependent-issue latency of scalar vs packed (f32x2) FP32 ops on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp2_ubench fp2_ubench.cu
// Output: one line per op: warp-instructions per clock per SM (throughput test, 16 warps/SM-quadrant... see below)
// and cycles per dependent instruction (latency test, one warp).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

enum Op { FADD, FMUL, FFMA, FADD2, FMUL2, FFMA2, MIX12 };

template <int OP> __device__ __forceinline__ void op1(float& a, float b, float c)
{
	if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b));
	if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b));
	if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
}
template <int OP> __device__ __forceinline__ void op2(u64& a, u64 b, u64 c)
{
	if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
	if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
	if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c));
}

template <int OP, int CHAINS> __global__ void bench(float* out, int iters, float b, float c, long long* cycles)
{
	float s[CHAINS];
	u64 p[CHAINS];
	u64 pb, pc;
	asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
	asm("mov.b64 %0, {%1, %1};" : "=l"(pc) : "f"(c));
#pragma unroll
	for (int k = 0; k < CHAINS; ++k) {
		s[k] = threadIdx.x * 0.001f + k;
		asm("mov.b64 %0, {%1, %1};" : "=l"(p[k]) : "f"(s[k]));
	}
	long long t0 = clock64();
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int r = 0; r < 8; ++r) {
#pragma unroll
			for (int k = 0; k < CHAINS; ++k) {
				if (OP < FADD2) op1<OP>(s[k], b, c);
				else if (OP < MIX12) op2<OP>(p[k], pb, pc);
				else { op2<FFMA2>(p[k], pb, pc); op1<FADD>(s[k], b, c); }
			}
		}
	}
	long long t1 = clock64();
	float acc = 0;
#pragma unroll
	for (int k = 0; k < CHAINS; ++k) {
		float lo, hi;
		asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[k]));
		acc += s[k] + lo + hi;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP> void run(const char* name, int sms)
{
	float* out; long long* cyc; long long h;
	cudaMalloc(&out, sizeof(float) * 4096 * 1024);
	cudaMalloc(&cyc, 8);
	const int iters = 4096;
	// latency: one warp, one chain
	bench<OP, 1><<<1, 32>>>(out, iters, 1.0001f, 0.5f, cyc);
	cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
	double lat = double(h) / (iters * 8.0) / (OP == MIX12 ? 2 : 1);
	// throughput: 8 chains x 16 warps/SM (4 per scheduler) on every SM; clock64 of block 0
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	bench<OP, 8><<<sms, 512>>>(out, iters, 1.0001f, 0.5f, cyc);
	cudaEventRecord(e0);
	bench<OP, 8><<<sms, 512>>>(out, iters, 1.0001f, 0.5f, cyc);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	double instr_per_sm = double(iters) * 8 * 8 * 16 * (OP == MIX12 ? 2 : 1);
	printf("%-6s dependent-issue latency %.2f clk | %.3f warp-instr/clk/SM (%.3f per scheduler), kernel %.3f ms\n", name, lat,
		instr_per_sm / double(h), instr_per_sm / double(h) / 4, ms);
	cudaFree(out); cudaFree(cyc);
}

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
	run<FADD>("FADD", p.multiProcessorCount);
	run<FMUL>("FMUL", p.multiProcessorCount);
	run<FFMA>("FFMA", p.multiProcessorCount);
	run<FADD2>("FADD2", p.multiProcessorCount);
	run<FMUL2>("FMUL2", p.multiProcessorCount);
	run<FFMA2>("FFMA2", p.multiProcessorCount);
	run<MIX12>("FFMA2+FADD", p.multiProcessorCount);
	return 0;
}
