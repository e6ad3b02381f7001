// k_quartet.cu -- the four-stage warp pipelines (quartet.cuh).
#include "launch.h"
#include "quartet.cuh"

namespace oalsfx {

bool launch_quartet_family(int kernel_id, const MixArgs& args, cudaStream_t st)
{
	static bool done[kKernelEnd] = {};
	switch (kernel_id) {
#define OALSFX_TX(id, CT, F0, F1, F2, F3, twin) \
	case id: { \
		const size_t dyn = prefer_shared(done[id], quartet::quartet_kernel<CT, F0, F1, F2, F3>, 70); \
		quartet::quartet_kernel<CT, F0, F1, F2, F3><<<static_cast<unsigned>(args.tile_count), quartet::kThreads, dyn, st>>>(args); \
		return true; }
		OALSFX_QUARTET_TABLE(OALSFX_TX)
#undef OALSFX_TX
	default: return false;
	}
}

} // namespace oalsfx
