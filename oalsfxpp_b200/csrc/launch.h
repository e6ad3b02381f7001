// launch.h -- the kernel families, one translation unit each (k_*.cu), so that nvcc compiles them in parallel.
//
// Every launcher answers whether `kernel_id` belongs to its family; if it does, the kernel has been launched on
// `stream` and the caller checks cudaGetLastError().
#ifndef OALSFX_LAUNCH_H
#define OALSFX_LAUNCH_H

#include <cuda_runtime.h>

#include <cstdlib>

#include "kernel_table.h"

namespace oalsfx {

bool launch_mix_family(int kernel_id, const MixArgs& args, cudaStream_t stream);      // k_mix.cu: thread per stream (Gen*, fused, Tab*)
bool launch_duo_family(int kernel_id, const MixArgs& args, cudaStream_t stream);      // k_duo.cu: duo_kernel, duo_multi_kernel
bool launch_quartet_family(int kernel_id, const MixArgs& args, cudaStream_t stream);  // k_quartet.cu
bool launch_relay_family(int kernel_id, const MixArgs& args, cudaStream_t stream);    // k_relay.cu: relay_kernel, relay_multi_kernel
bool launch_span_family(int kernel_id, const MixArgs& args, cudaStream_t stream);     // k_span.cu: block-parallel in time
bool launch_scan_family(int kernel_id, const MixArgs& args, cudaStream_t stream);     // k_scan.cu: linear-recurrence scan over time

// Shared-memory carve-out of a kernel, set once.  Tuning knobs (experiments only): OALSFX_TUNE_CARVEOUT = percent or
// -1 (driver default), OALSFX_TUNE_DYN_SMEM = bytes of unused dynamic shared memory per CTA (caps residency).
template <class K>
inline size_t prefer_shared(bool& done, K kernel, int default_carveout)
{
	static size_t tune_dyn_smem = 0;
	if (!done) {
		int carveout = default_carveout;
		if (const char* e = getenv("OALSFX_TUNE_CARVEOUT")) {
			carveout = atoi(e);
		}
		if (const char* e = getenv("OALSFX_TUNE_DYN_SMEM")) {
			tune_dyn_smem = static_cast<size_t>(atoi(e));
			cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(tune_dyn_smem));
		}
		if (carveout >= 0) {
			cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
		}
		done = true;
	}
	return tune_dyn_smem;
}

} // namespace oalsfx

#endif
