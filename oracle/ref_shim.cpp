// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product (oalsfxpp_b200/).
//
// C-ABI shim around the UNMODIFIED reference `oalsfxpp::Api`
// (/root/reference/src/oalsfxpp.h:760-922).  oracle/Makefile compiles this file together with
// /root/reference/src/oalsfxpp.cpp *where it lies* into oracle/_ref/liboalsfx_ref.so
// (git-ignored, travels to the GPU box as a binary).  No reference source is copied.
//
// The same `orc_*` ABI is exported by the hand-written CPU restatement
// (oracle/oalsfx_oracle.c -> oracle/_build/liboalsfx_oracle.so) so tests drive both alike.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may load this library.

#include "oalsfxpp.h"

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

using oalsfxpp::Api;

namespace {

// Synthetic white noise shared by oracle, tests and bench: murmur3 fmix32
// of (seed, stream, channel, frame) -> uniform [-0.5, 0.5), exact in fp32.
inline uint32_t fmix32(uint32_t h)
{
	h ^= h >> 16; h *= 0x85EBCA6BU; h ^= h >> 13; h *= 0xC2B2AE35U; h ^= h >> 16;
	return h;
}
inline float noise_sample(uint32_t seed, uint32_t stream, uint32_t chan, uint32_t frame)
{
	const uint32_t h = fmix32(seed ^ (stream * 0x9E3779B9U) ^ (chan * 0x85EBCA6BU) ^ (frame * 0xC2B2AE35U));
	return (static_cast<float>(h >> 8) * (1.0F / 8388608.0F) - 1.0F) * 0.5F;
}

} // namespace

extern "C" {

const char* orc_kind() { return "reference"; }

void* orc_create(int channel_format, int sampling_rate, int effect_count)
{
	auto* api = new Api{};
	if (!api->initialize(static_cast<oalsfxpp::ChannelFormat>(channel_format), sampling_rate, effect_count))
	{
		delete api;
		return nullptr;
	}
	return api;
}

void orc_destroy(void* h) { delete static_cast<Api*>(h); }

int orc_channel_count(void* h) { return static_cast<Api*>(h)->get_channel_count(); }

int orc_set_effect_type(void* h, int slot, int type)
{
	return static_cast<Api*>(h)->set_effect_type(slot, static_cast<oalsfxpp::EffectType>(type)) ? 1 : 0;
}

// props: raw bytes of the reference's `union EffectProps` (108 bytes).
int orc_set_effect_props(void* h, int slot, const void* props)
{
	oalsfxpp::EffectProps p;
	std::memcpy(&p, props, sizeof(p));
	return static_cast<Api*>(h)->set_effect_props(slot, p) ? 1 : 0;
}

// Reads back the ACTIVE effect (type + props bytes) of a slot.
int orc_get_effect(void* h, int slot, int* type, void* props)
{
	oalsfxpp::Effect e;
	if (!static_cast<Api*>(h)->get_effect(slot, e))
	{
		return 0;
	}
	*type = static_cast<int>(e.type_);
	std::memcpy(props, &e.props_, sizeof(e.props_));
	return 1;
}

int orc_get_deferred_effect(void* h, int slot, int* type, void* props)
{
	oalsfxpp::Effect e;
	if (!static_cast<Api*>(h)->get_deferred_effect(slot, e))
	{
		return 0;
	}
	*type = static_cast<int>(e.type_);
	std::memcpy(props, &e.props_, sizeof(e.props_));
	return 1;
}

// send_index < 0 addresses the direct send.  gains = {gain, gain_hf, gain_lf}.
int orc_set_send_props(void* h, int send_index, const float* gains)
{
	oalsfxpp::SendProps p;
	p.gain_ = gains[0];
	p.gain_hf_ = gains[1];
	p.gain_lf_ = gains[2];
	return static_cast<Api*>(h)->set_send_props(send_index, p) ? 1 : 0;
}

int orc_apply(void* h) { return static_cast<Api*>(h)->apply_changes() ? 1 : 0; }

int orc_mix(void* h, int frames, const float* src, float* dst)
{
	return static_cast<Api*>(h)->mix(frames, src, dst) ? 1 : 0;
}

int orc_sizeof_effect_props() { return static_cast<int>(sizeof(oalsfxpp::EffectProps)); }
int orc_sizeof_effect() { return static_cast<int>(sizeof(oalsfxpp::Effect)); }

void orc_noise(uint32_t seed, uint32_t stream, int channels, int first_frame, int frames, float* dst)
{
	for (int n = 0; n < frames; ++n)
	{
		for (int c = 0; c < channels; ++c)
		{
			dst[(n * channels) + c] = noise_sample(seed, stream, static_cast<uint32_t>(c), static_cast<uint32_t>(first_frame + n));
		}
	}
}

// Multi-threaded CPU baseline driver (SURVEY.md 8d "CPU baseline timing").
//
// `n_streams` independent Api instances, each with `n_slots` slots of the given effect types at
// their defaults, are distributed over `n_threads` worker threads (one instance live per thread at
// a time).  Every instance mixes `n_blocks` blocks of `block_frames` frames of the synthetic white
// noise above (each worker pre-generates 8 distinct blocks once and cycles through them, so
// input generation stays out of the measurement).  Returns wall seconds for the whole job
// (instance construction included; it is < 1 % of the work at >= 32 blocks); checksum_out
// receives a sum over outputs so the work cannot be optimised away.
double orc_bench(
	int n_threads, int n_streams, int channel_format, int sampling_rate,
	const int* slot_types, int n_slots, int block_frames, int n_blocks, uint32_t seed,
	double* checksum_out)
{
	std::atomic<int> next{0};
	std::vector<double> sums(static_cast<size_t>(n_threads), 0.0);
	const auto t0 = std::chrono::steady_clock::now();
	auto worker = [&](int tid)
	{
		constexpr int unique_blocks = 8;
		std::vector<float> src(static_cast<size_t>(unique_blocks) * block_frames * 8);
		std::vector<float> dst(static_cast<size_t>(block_frames) * 8);
		int src_ch = 0;
		for (;;)
		{
			const int s = next.fetch_add(1);
			if (s >= n_streams)
			{
				break;
			}
			Api api;
			if (!api.initialize(static_cast<oalsfxpp::ChannelFormat>(channel_format), sampling_rate, n_slots))
			{
				break;
			}
			for (int i = 0; i < n_slots; ++i)
			{
				api.set_effect_type(i, static_cast<oalsfxpp::EffectType>(slot_types[i]));
			}
			api.apply_changes();
			const int ch = api.get_channel_count();
			if (src_ch != ch)
			{
				orc_noise(seed, static_cast<uint32_t>(tid), ch, 0, unique_blocks * block_frames, src.data());
				src_ch = ch;
			}
			double acc = 0.0;
			for (int b = 0; b < n_blocks; ++b)
			{
				const float* in = src.data() + static_cast<size_t>(b % unique_blocks) * block_frames * ch;
				api.mix(block_frames, in, dst.data());
				acc += dst[static_cast<size_t>(block_frames) * ch - 1];
			}
			sums[static_cast<size_t>(tid)] += acc;
		}
	};
	std::vector<std::thread> threads;
	for (int t = 0; t < n_threads; ++t)
	{
		threads.emplace_back(worker, t);
	}
	for (auto& t : threads)
	{
		t.join();
	}
	const auto t1 = std::chrono::steady_clock::now();
	double total = 0.0;
	for (double v : sums)
	{
		total += v;
	}
	if (checksum_out)
	{
		*checksum_out = total;
	}
	return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
