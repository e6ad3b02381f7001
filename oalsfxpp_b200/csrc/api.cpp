// api.cpp -- oalsfxpp::Api, the reference's public class (reference: src/oalsfxpp.cpp:3468-3903),
// re-implemented as a thin host shell over the C ABI (include/oalsfx_engine.h) with one engine
// stream per Api instance.  The shell owns what the reference's Api/Impl own on the control side:
// argument checks, deferred -> active property commit, the "changed" flags, error strings.  All
// sample processing happens in the CUDA engine; there is no CPU path.
//
// Reference quirks kept on purpose, each marked QUIRK below; the two that crash the
// reference (null pimpl in get_error_message) are made safe instead.
#include "oalsfxpp.h"

#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "oalsfx_engine.h"

namespace oalsfxpp {
namespace {

constexpr int kMaxEffects = 4;
const char* const kNoError = "";
const char* const kAllocateImpl = "Failed to allocate implementaion class."; // sic, oalsfxpp.cpp:3452
const char* const kNotInitialized = "Not initialized.";
const char* const kEffectIndex = "Effect index is out of range.";
const char* const kNoSrc = "No source samples.";
const char* const kNoDst = "No destination samples.";

int channel_count_of(const ChannelFormat f)
{
	switch (f) {
	case ChannelFormat::mono: return 1;
	case ChannelFormat::stereo: return 2;
	case ChannelFormat::quad: return 4;
	case ChannelFormat::five_point_one:
	case ChannelFormat::five_point_one_rear: return 6;
	case ChannelFormat::six_point_one: return 7;
	case ChannelFormat::seven_point_one: return 8;
	default: return 0;
	}
}

} // namespace

class Api::Impl {
public:
	struct Slot {
		Effect deferred;   // EffectContext::deferred_effect_
		Effect active;     // EffectSlot::effect_
		bool changed;      // EffectSlot::is_props_changed_
	};
	struct Send {
		SendProps props;
		SendProps deferred;
	};

	oalsfx_engine* engine = nullptr;
	ChannelFormat format = ChannelFormat::none;
	int rate = 0;
	int channels = 0;
	int effect_count = 0;
	Slot slots[kMaxEffects];
	Send direct;
	Send aux[kMaxEffects];
	bool source_changed = true; // Source::are_props_changed_
	std::string error_text;
	const char* error_message = kNoError;

	~Impl() { oalsfx_engine_destroy(engine); }

	bool engine_failed()
	{
		error_text = oalsfx_last_error(engine);
		error_message = error_text.c_str();
		return false;
	}

	// The top of the reference's first chunk: update_context_sources (oalsfxpp.cpp:3397-3412).
	bool push_changes()
	{
		bool updated = false;
		for (int i = 0; i < effect_count; ++i) {
			if (!slots[i].changed) {
				continue;
			}
			slots[i].changed = false;
			updated = true;
			if (oalsfx_engine_set_effect(engine, 0, 1, i, static_cast<int>(slots[i].active.type_),
					&slots[i].active.props_, sizeof(EffectProps)) != OALSFX_OK) {
				return engine_failed();
			}
		}
		if (source_changed) {
			source_changed = false;
			updated = true;
		}
		if (updated) {
			// calc_non_attn_source_params reads direct.props_ and each aux.props_ as they are now.
			const float d[3] = {direct.props.gain_, direct.props.gain_hf_, direct.props.gain_lf_};
			float a[3 * kMaxEffects] = {};
			for (int i = 0; i < effect_count; ++i) {
				a[3 * i] = aux[i].props.gain_;
				a[3 * i + 1] = aux[i].props.gain_hf_;
				a[3 * i + 2] = aux[i].props.gain_lf_;
			}
			if (oalsfx_engine_set_sends(engine, 0, 1, d, a) != OALSFX_OK) {
				return engine_failed();
			}
		}
		return true;
	}
};

Api::Api() : pimpl_{}, error_message_{kNoError} {}

Api::~Api() { uninitialize(); }

bool Api::initialize(const ChannelFormat channel_format, const int sampling_rate, const int effect_count)
{
	uninitialize();
	pimpl_.reset(new (std::nothrow) Impl{});
	if (!pimpl_) {
		error_message_ = kAllocateImpl;
		return false;
	}
	oalsfx_engine_desc desc;
	const char* dev_env = std::getenv("OALSFX_DEVICE");
	desc.device = dev_env ? std::atoi(dev_env) : 0;
	desc.num_streams = 1;
	desc.channel_format = static_cast<int>(channel_format);
	desc.sampling_rate = sampling_rate;
	desc.effect_count = effect_count;
	const int rc = oalsfx_engine_create(&desc, &pimpl_->engine);
	if (rc != OALSFX_OK) {
		// Same three messages as the reference for its three checks (oalsfxpp.cpp:2807-2810); a
		// missing CUDA device reports the engine's own text.
		static std::string last_create_error;
		last_create_error = oalsfx_last_error(nullptr);
		error_message_ = last_create_error.c_str();
		uninitialize();
		return false;
	}
	Impl& p = *pimpl_;
	p.format = channel_format;
	p.rate = sampling_rate;
	p.channels = channel_count_of(channel_format);
	p.effect_count = effect_count;
	for (int i = 0; i < kMaxEffects; ++i) {
		std::memset(&p.slots[i], 0, sizeof(p.slots[i]));
		p.slots[i].deferred.set_type_and_defaults(EffectType::null);
		p.slots[i].active.type_ = EffectType::null;
		p.slots[i].changed = false; // a null slot's update is a no-op
		p.aux[i].props.set_defaults();
		p.aux[i].deferred.set_defaults();
	}
	p.direct.props.set_defaults();
	p.direct.deferred.set_defaults();
	p.source_changed = true;
	return true;
}

bool Api::is_initialized() const { return pimpl_ != nullptr; }

int Api::get_sampling_rate() const
{
	if (!is_initialized()) {
		error_message_ = kNotInitialized;
		return 0;
	}
	return pimpl_->rate;
}

ChannelFormat Api::get_channel_format() const
{
	if (!is_initialized()) {
		error_message_ = kNotInitialized;
		return ChannelFormat::none;
	}
	return pimpl_->format;
}

int Api::get_channel_count() const
{
	if (!is_initialized()) {
		error_message_ = kNotInitialized;
		return 0;
	}
	return pimpl_->channels;
}

int Api::get_effect_count() const
{
	if (!is_initialized()) {
		error_message_ = kNotInitialized;
		return 0;
	}
	return pimpl_->effect_count;
}

// Shared precondition of the per-slot calls (oalsfxpp.cpp:3557-3568 and siblings).
#define REQUIRE_SLOT(index, allow_negative) \
	if (!is_initialized()) { error_message_ = kNotInitialized; return false; } \
	if ((!(allow_negative) && (index) < 0) || (index) >= pimpl_->effect_count) { error_message_ = kEffectIndex; return false; }

bool Api::get_effect(const int effect_index, Effect& effect) const
{
	REQUIRE_SLOT(effect_index, false)
	effect = pimpl_->slots[effect_index].active;
	return true;
}

bool Api::get_deferred_effect(const int effect_index, Effect& effect) const
{
	REQUIRE_SLOT(effect_index, false)
	effect = pimpl_->slots[effect_index].deferred;
	return true;
}

bool Api::set_effect_type(const int effect_index, const EffectType effect_type)
{
	REQUIRE_SLOT(effect_index, false)
	pimpl_->slots[effect_index].deferred.set_type_and_defaults(effect_type);
	return true;
}

bool Api::set_effect_props(const int effect_index, const EffectProps& effect_props)
{
	REQUIRE_SLOT(effect_index, false)
	pimpl_->slots[effect_index].deferred.props_ = effect_props;
	return true;
}

bool Api::set_effect(const int effect_index, const Effect& effect)
{
	REQUIRE_SLOT(effect_index, false)
	pimpl_->slots[effect_index].deferred = effect;
	return false; // QUIRK: the reference returns false on success (oalsfxpp.cpp:3657)
}

bool Api::get_send_props(const int effect_index, SendProps& send_props) const
{
	REQUIRE_SLOT(effect_index, true)
	send_props = (effect_index < 0 ? pimpl_->direct.props : pimpl_->aux[effect_index].props);
	return true;
}

bool Api::get_deferred_send_props(const int effect_index, SendProps& send_props) const
{
	REQUIRE_SLOT(effect_index, true)
	send_props = (effect_index < 0 ? pimpl_->direct.deferred : pimpl_->aux[effect_index].deferred);
	return true;
}

bool Api::set_send_props(const int effect_index, const SendProps& send_props)
{
	REQUIRE_SLOT(effect_index, true)
	if (effect_index < 0) {
		pimpl_->direct.deferred = send_props;
	} else {
		// QUIRK: aux sends bypass deferral and normalization (oalsfxpp.cpp:3728-3731); the values are
		// picked up by the next source-parameter refresh.
		pimpl_->aux[effect_index].props = send_props;
	}
	return true;
}

bool Api::apply_changes()
{
	if (!is_initialized()) {
		error_message_ = kNotInitialized;
		return false;
	}
	Impl& p = *pimpl_;
	for (int i = 0; i < p.effect_count; ++i) {
		Impl::Slot& s = p.slots[i];
		s.deferred.normalize();
		if (!Effect::are_equal(s.deferred, s.active)) {
			// EffectSlot::set_effect (oalsfxpp.cpp:2688-2709); the engine resets the slot state on a
			// type change when the change is pushed at the next mix.
			s.active = s.deferred;
			s.changed = true;
		}
	}
	p.direct.deferred.normalize();
	if (!SendProps::are_equal(p.direct.deferred, p.direct.props)) {
		p.source_changed = true;
		p.direct.props = p.direct.deferred;
	}
	for (int i = 0; i < p.effect_count; ++i) {
		p.aux[i].deferred.normalize();
		if (!SendProps::are_equal(p.aux[i].props, p.aux[i].deferred)) {
			p.source_changed = true; // QUIRK: flags a change but never copies (oalsfxpp.cpp:3772-3780)
		}
	}
	return true;
}

bool Api::mix(const int sample_count, const float* src_samples, float* dst_samples)
{
	if (!is_initialized()) {
		error_message_ = kNotInitialized;
		return false;
	}
	if (sample_count == 0) {
		return true;
	}
	if (!src_samples) {
		error_message_ = kNoSrc;
		return false;
	}
	if (!dst_samples) {
		error_message_ = kNoDst;
		return false;
	}
	if (sample_count < 0) {
		return true; // QUIRK: a negative count is a no-op (the reference's loop never runs, oalsfxpp.cpp:3818)
	}
	Impl& p = *pimpl_;
	if (!p.push_changes()) {
		error_message_ = p.error_message;
		return false;
	}
	if (oalsfx_engine_mix(p.engine, sample_count, src_samples, dst_samples, OALSFX_LAYOUT_STREAM_MAJOR,
			OALSFX_SPACE_HOST, nullptr) != OALSFX_OK) {
		p.engine_failed();
		error_message_ = p.error_message;
		return false;
	}
	return true;
}

void Api::uninitialize() { pimpl_ = nullptr; }

const char* Api::get_error_message() const
{
	// The reference dereferences a null pimpl here when uninitialised (oalsfxpp.cpp:3836-3839); this
	// shell answers with the Api-level message instead of crashing.
	if (!pimpl_) {
		return error_message_;
	}
	return (error_message_ && error_message_[0] != '\0') ? error_message_ : pimpl_->error_message;
}

int Api::get_min_channels() { return 1; }
int Api::get_max_channels() { return 8; }
int Api::get_min_sampling_rate() { return 8000; }
int Api::get_max_sampling_rate() { return 8000000; }
int Api::get_min_effects() { return 1; }
int Api::get_max_effects() { return kMaxEffects; }

ChannelFormat Api::channel_count_to_channel_format(const int channel_count)
{
	switch (channel_count) {
	case 1: return ChannelFormat::mono;
	case 2: return ChannelFormat::stereo;
	case 4: return ChannelFormat::quad;
	case 6: return ChannelFormat::five_point_one;
	case 7: return ChannelFormat::six_point_one;
	case 8: return ChannelFormat::seven_point_one;
	default: return ChannelFormat::none;
	}
}

int Api::channel_format_to_channel_count(const ChannelFormat channel_format) { return channel_count_of(channel_format); }

} // namespace oalsfxpp
