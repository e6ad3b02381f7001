// k_span.cu -- kernels that are block-parallel in time (span.cuh).
#include "launch.h"
#include "span.cuh"

namespace oalsfx {

bool launch_span_family(int kernel_id, const MixArgs& args, cudaStream_t st)
{
	static bool done[kKernelEnd] = {};
	switch (kernel_id) {
#define OALSFX_SX(id, CT, SL) \
	case id: \
		if (!done[id]) { \
			cudaFuncSetAttribute(span::span_reverb_kernel<CT, SL>, cudaFuncAttributeMaxDynamicSharedMemorySize, span::shared_floats(CT, SL) * static_cast<int>(sizeof(float))); \
			done[id] = true; \
		} \
		span::span_reverb_kernel<CT, SL><<<static_cast<unsigned>(args.tile_count) * (kLanes / SL), span::kThreads, \
			static_cast<size_t>(span::shared_floats(CT, SL)) * sizeof(float), st>>>(args); \
		return true;
		OALSFX_SPAN_TABLE(OALSFX_SX)
#undef OALSFX_SX
	default: return false;
	}
}

} // namespace oalsfx
