// k_span.cu -- kernels that are block-parallel in time (span.cuh).
#include "launch.h"
#include "span.cuh"

namespace oalsfx {

bool launch_span_family(int kernel_id, const MixArgs& args, cudaStream_t st)
{
	static bool done[kKernelEnd] = {};
	switch (kernel_id) {
#define OALSFX_SX(id, CT, SL, CHAIN) \
	case id: { \
		constexpr size_t bytes = static_cast<size_t>(span::shared_floats(CHAIN, SL)) * sizeof(float); \
		if (!done[id]) { \
			cudaFuncSetAttribute(span::span_kernel<CT, SL, CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)); \
			done[id] = true; \
		} \
		span::span_kernel<CT, SL, CHAIN><<<static_cast<unsigned>(args.tile_count) * (kLanes / SL), span::threads(CHAIN), bytes, st>>>(args); \
		return true; }
		OALSFX_SPAN_TABLE(OALSFX_SX)
#undef OALSFX_SX
#define OALSFX_BX(id, CT, CHAIN) \
	case id: { \
		constexpr size_t bytes = static_cast<size_t>(span::bulk_shared_floats(CHAIN)) * sizeof(float); \
		if (!done[id]) { \
			cudaFuncSetAttribute(span::span_bulk_kernel<CT, CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)); \
			done[id] = true; \
		} \
		span::span_bulk_kernel<CT, CHAIN><<<static_cast<unsigned>(args.tile_count), span::bulk_threads(CHAIN), bytes, st>>>(args); \
		return true; }
		OALSFX_SPAN_BULK_TABLE(OALSFX_BX)
#undef OALSFX_BX
	default: return false;
	}
}

} // namespace oalsfx
