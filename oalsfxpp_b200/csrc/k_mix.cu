// k_mix.cu -- thread-per-stream kernels: the exact single-effect passes (Gen*), the fused twins and the table-mode passes.
#include "launch.h"

namespace oalsfx {

bool launch_mix_family(int kernel_id, const MixArgs& args, cudaStream_t st)
{
	constexpr int threads = 64;
	const unsigned blocks = static_cast<unsigned>((static_cast<long long>(args.tile_count) * kLanes + threads - 1) / threads);
	switch (kernel_id) {
#define OALSFX_X(id, CT, SF, F0, F1, F2, F3) \
	case id: mix_kernel<CT, SF, F0, F1, F2, F3><<<blocks, threads, 0, st>>>(args); return true;
		OALSFX_KERNEL_TABLE(OALSFX_X)
#undef OALSFX_X
#define OALSFX_TBX(id, Fx, kind) \
	case id: mix_kernel<0, true, Fx, FxNull, FxNull, FxNull, true><<<blocks, threads, 0, st>>>(args); return true;
		OALSFX_TABMODE_TABLE(OALSFX_TBX)
#undef OALSFX_TBX
	default: return false;
	}
}

} // namespace oalsfx
