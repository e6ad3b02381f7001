// k_relay.cu -- the relay pipelines (relay.cuh): any slot signature, one warp per non-null slot.
#include "launch.h"
#include "relay.cuh"

namespace oalsfx {

namespace {
// Up to four reverb windows of dynamic shared memory (96 KB) behind the wide kernels' exchange, carve-out as large as needed.
template <int CT>
constexpr size_t exchange_bytes() { return CT ? 0 : sizeof(relay::Shared<0>); }

template <class K>
void relay_attributes(bool& done, K kernel, size_t exchange)
{
	if (!done) {
		cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSlots * kPfWarpFloats * static_cast<int>(sizeof(float)) + static_cast<int>(exchange));
		cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
		done = true;
	}
}
} // namespace

bool launch_relay_family(int kernel_id, const MixArgs& args, cudaStream_t st)
{
	static bool done[kKernelEnd] = {};
	const size_t dyn = static_cast<size_t>(args.relay_smem_floats) * sizeof(float);
	switch (kernel_id) {
#define OALSFX_RX(id, CT, HEAVY) \
	case id: \
		relay_attributes(done[id], relay::relay_kernel<CT, HEAVY>, exchange_bytes<CT>()); \
		relay::relay_kernel<CT, HEAVY><<<static_cast<unsigned>(args.tile_count), static_cast<unsigned>(kLanes * args.relay_count), dyn + exchange_bytes<CT>(), st>>>(args); \
		return true;
		OALSFX_RELAY_TABLE(OALSFX_RX)
#undef OALSFX_RX
#define OALSFX_RX(id, CT, HEAVY) \
	case id: \
		relay_attributes(done[id], relay::relay_multi_kernel<CT, HEAVY>, exchange_bytes<CT>()); \
		relay::relay_multi_kernel<CT, HEAVY><<<static_cast<unsigned>(args.tile_count), static_cast<unsigned>(kLanes * args.relay_count), dyn + exchange_bytes<CT>(), st>>>(args); \
		return true;
		OALSFX_RELAY_MULTI_TABLE(OALSFX_RX)
#undef OALSFX_RX
#define OALSFX_RX(id, CT, HEAVY) \
	case id: \
		relay_attributes(done[id], relay::relay_sf_kernel<CT, HEAVY>, exchange_bytes<CT>()); \
		relay::relay_sf_kernel<CT, HEAVY><<<static_cast<unsigned>(args.tile_count), static_cast<unsigned>(kLanes * args.relay_count), dyn + exchange_bytes<CT>(), st>>>(args); \
		return true;
		OALSFX_RELAY_SF_TABLE(OALSFX_RX)
#undef OALSFX_RX
	default: return false;
	}
}

} // namespace oalsfx
