// k_duo.cu -- the two-warp fused kernels (duo.cuh): one parameter class per launch, and one per tile.
//
// They stream through per-warp shared-memory windows (~25 KB per CTA) AND lean on L1: the stream-major input rows
// are read 4 bytes per lane per frame, so each 128-byte line serves 16 consecutive frames from L1.  Measured on B200
// (65 536 streams): the optimum is 4 CTAs per SM with the rest of the 228 KB as L1 (carve-out 46-54 %: 3.07 ms);
// 5 CTAs (62-80 %) 3.23 ms; 6 CTAs / ~28 KB of L1 (100 %) 3.8 ms; 3 CTAs 3.2 ms.
// The class-per-tile kernel holds its tile's coefficient block in shared memory on top (29 KB per CTA): carve-out 75 %
// (65 536 streams, 113 classes: 40 % 6.23 ms, 50 % 5.34, 62 % 4.97, 75 % 4.72, 100 % 4.97; 16 384 streams: 1.96-1.98 throughout).
#include "launch.h"
#include "duo.cuh"

namespace oalsfx {

bool launch_duo_family(int kernel_id, const MixArgs& args, cudaStream_t st)
{
	static bool done[kKernelEnd] = {};
	switch (kernel_id) {
#define OALSFX_DX(id, CT, F0, F1, F2, F3, twin) \
	case id: { \
		const size_t dyn = prefer_shared(done[id], duo::duo_kernel<CT, F0, F1, F2, F3>, 50); \
		duo::duo_kernel<CT, F0, F1, F2, F3><<<static_cast<unsigned>(args.tile_count), 64, dyn, st>>>(args); \
		return true; }
		OALSFX_DUO_TABLE(OALSFX_DX)
#undef OALSFX_DX
#define OALSFX_MX(id, CT, F0, F1, F2, F3, duo_id) \
	case id: { \
		const size_t dyn = prefer_shared(done[id], duo::duo_multi_kernel<CT, F0, F1, F2, F3>, 75); \
		duo::duo_multi_kernel<CT, F0, F1, F2, F3><<<static_cast<unsigned>(args.tile_count), 64, dyn, st>>>(args); \
		return true; }
		OALSFX_MULTI_TABLE(OALSFX_MX)
#undef OALSFX_MX
	default: return false;
	}
}

} // namespace oalsfx
