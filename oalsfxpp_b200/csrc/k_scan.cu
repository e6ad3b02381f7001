// k_scan.cu -- biquad cascades as a linear-recurrence scan over time (scan.cuh).
#include "launch.h"
#include "scan.cuh"

namespace oalsfx {

bool launch_scan_family(int kernel_id, const MixArgs& args, cudaStream_t st)
{
	switch (kernel_id) {
#define OALSFX_SCX(id, CT) \
	case id: \
		scan::scan_equalizer_kernel<CT><<<static_cast<unsigned>(args.num_streams), scan::kThreads, 0, st>>>(args); \
		return true;
		OALSFX_SCAN_TABLE(OALSFX_SCX)
#undef OALSFX_SCX
	default: return false;
	}
}

} // namespace oalsfx
