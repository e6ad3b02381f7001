// derive.h -- host-side parameter -> coefficient derivation.
// O(1) per parameter change, heavy in libm calls, therefore kept on the host so that every
// coefficient is bit-identical to what the CPU reference computes with the same glibc.
#ifndef OALSFX_DERIVE_H
#define OALSFX_DERIVE_H

#include <cstdint>
#include <vector>

#include "coefs.h"
#include "oalsfxpp.h"

namespace oalsfx {

// Output-device description (reference: struct Device, oalsfxpp.cpp:2378-2625, and the source
// channel maps, oalsfxpp.cpp:3048-3098).
struct DeviceLayout {
	int channel_format = 0;
	int channels = 0;                  // device / interleave channel count
	int dry_coeff_count = 0;           // ambisonic coefficients used by the dry decoder
	float dry[kMaxChannels][16] = {};  // decoder rows per output channel (LFE rows are zero)
	float foa[kMaxChannels][4] = {};   // first-order subset
	int source_channels = 0;           // entries in the source channel map (0 for 5.1-rear: reference quirk)
	float source_angle[kMaxChannels] = {};
	bool source_is_lfe[kMaxChannels] = {};
};

bool make_device_layout(int channel_format, DeviceLayout& out);

struct SendSettings { float gain, gain_hf, gain_lf; };

// Direct + aux send coefficients (reference: calc_non_attn_source_params / calc_panning_and_filters,
// oalsfxpp.cpp:3348-3395, 3172-3346).
void derive_sends(
	const DeviceLayout& dev, int sampling_rate, int effect_count,
	const SendSettings& direct, const SendSettings* aux,
	SendCoef& direct_out, SendCoef* aux_out);

// Host-evaluated lookup tables that accompany a SlotCoef (uploaded by the engine, which then
// patches the device pointers into the block).
struct SlotTables {
	std::vector<int32_t> sin_delays;   // chorus/flanger sinusoid LFO delays, [lfo_range]
	std::vector<float> mod_sinus;      // reverb modulator sinus, [mod_range]
};

// Effect coefficients for one slot (reference: each EffectState::do_update_device + do_update).
// `props` must already be normalized.  Table pointers inside `out` are left null.
void derive_slot(
	const DeviceLayout& dev, int sampling_rate, int fx_type, const oalsfxpp::EffectProps& props,
	SlotCoef& out, SlotTables& tables);

} // namespace oalsfx

#endif
