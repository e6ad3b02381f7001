// TEST INFRASTRUCTURE ONLY -- a Backend (oalsfxpp_b200/csrc/backend.h) that runs the engine's
// __host__ __device__ kernel bodies on the CPU, one (tile, lane) at a time.
//
// Purpose: let `pytest -m "not gpu"` exercise the engine's host logic (classes, grouping, arenas,
// block chunking, multi-pass chaining, the Api shell) and the per-sample arithmetic against the
// oracle in a container without a GPU.  It is built into tests/_build/liboalsfx_emu.so by
// tests/emu/Makefile, is never part of the package and is not a fallback: liboalsfx_b200.so links
// cuda_backend.cu only and refuses to create an engine without a CUDA device.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

#include "backend.h"
#include "kernel_table.h"
#include "span.cuh"

namespace oalsfx {
namespace {

long long g_span_streams = 0; // streams x launches that took the span schedule (read by the tests)
long long g_span_bulk_streams = 0; // ... of those, with the bulk-copy timing

class HostBackend final : public Backend {
public:
	const char* name() const override { return "host-emu"; }
	void* alloc(size_t bytes) override { return std::calloc(bytes ? bytes : 1, 1); }
	void release(void* p) override { std::free(p); }
	bool zero(void* p, size_t bytes, void*) override { std::memset(p, 0, bytes); return true; }
	bool upload(void* dst, const void* src, size_t bytes, void*) override { std::memcpy(dst, src, bytes); return true; }
	bool download(void* dst, const void* src, size_t bytes, void*) override { std::memcpy(dst, src, bytes); return true; }
	bool copy_2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width, size_t rows, void*) override
	{
		for (size_t r = 0; r < rows; ++r) {
			std::memcpy(static_cast<char*>(dst) + r * dst_pitch, static_cast<const char*>(src) + r * src_pitch, width);
		}
		return true;
	}
	bool zero_lanes(uint32_t* base, long long tile_stride, int words, const TileRef* tiles, int n_tiles, void*) override
	{
		for (int i = 0; i < n_tiles; ++i) {
			for (int lane = 0; lane < kLanes; ++lane) {
				if (!((tiles[i].mask >> lane) & 1U)) {
					continue;
				}
				uint32_t* p = base + static_cast<long long>(tiles[i].tile) * tile_stride + lane;
				for (int w = 0; w < words; ++w) {
					p[static_cast<long long>(w) * kLanes] = 0U;
				}
			}
		}
		return true;
	}
	// The device-only kernel families are emulated with the thread-per-stream bodies, so that the engine's selection
	// logic, argument assembly and class tables are exercised on the CPU as well:
	//   span   -> span::emulate_stream: the device kernel's own phase bodies, its three-stage software pipeline executed
	//             serially in an adversarial order (any legal schedule must give the reference's values); streams off the
	//             steady state run the exact thread-per-stream body, as on the device
	//   relay  -> one exact single-effect pass per stage, in stage order, the running bus parked in dst
	//             (the same additions in the same order; sanitized zero gains are simply skipped)
	//   *multi -> tile by tile, the tile's class copied into the arguments as the kernels do in shared memory
	bool has_relay() const override { return true; }
	bool launch_mix(int kernel_id, const MixArgs& a, void* stream) override
	{
		int multi_base = -1;
		if (kernel_id == kMultiChainStereo) {
			multi_base = kChainStereo;
		}
#define OALSFX_RX(id, CT, HEAVY) if (kernel_id == id) { multi_base = (CT == 1 ? (HEAVY ? kRelayMonoHeavy : kRelayMono) : (HEAVY ? kRelayStereoHeavy : kRelayStereo)); }
		OALSFX_RELAY_MULTI_TABLE(OALSFX_RX)
#undef OALSFX_RX
		if (multi_base >= 0) {
			for (int w = 0; w < a.tile_count; ++w) {
				const int tile = a.tiles ? static_cast<int>(a.tiles[w].tile) : a.tile_first + w;
				const MixClassEntry& entry = a.class_table[a.tile_class[tile]];
				MixArgs t = a;
				std::memcpy(reinterpret_cast<char*>(&t) + kMixCoefOffset, entry.coefs, kMixCoefBytes);
				t.update_mask = a.update_mask & entry.pending;
				t.tiles = nullptr;
				t.tile_first = tile;
				t.tile_count = 1;
				t.class_table = nullptr;
				t.tile_class = nullptr;
				if (!launch_mix(multi_base, t, stream)) {
					return false;
				}
			}
			return true;
		}
		bool is_relay = false;
#define OALSFX_RX(id, CT, HEAVY) is_relay = is_relay || kernel_id == id;
		OALSFX_RELAY_TABLE(OALSFX_RX)
		OALSFX_RELAY_SF_TABLE(OALSFX_RX)
#undef OALSFX_RX
		if (is_relay) {
			static const int gen_for_kind[] = {kGenDry, kGenModDelay, kGenCompressor, kGenDedicated, kGenDistortion,
				kGenEcho, kGenEqualizer, kGenRingMod, kGenReverb};
			for (int p = 0; p < a.relay_count; ++p) {
				MixArgs t = a;
				for (int q = 0; q < kMaxSlots; ++q) {
					t.ring[q] = nullptr;
					t.slot_state[q] = nullptr;
					std::memset(&t.slot[q], 0, sizeof(t.slot[q]));
					std::memset(&t.aux[q], 0, sizeof(t.aux[q]));
				}
				t.with_dry = p == 0 ? 1 : 0;
				t.accumulate = p == 0 ? 0 : 1;
				t.update_mask = (a.update_mask >> p) & 1U;
				t.aux_index[0] = a.aux_index[p];
				t.ring[0] = a.ring[p];
				t.ring_tile_stride[0] = a.ring_tile_stride[p];
				t.slot_state[0] = a.slot_state[p];
				t.slot[0] = a.slot[p];
				t.aux[0] = a.aux[p];
				t.relay_count = 0;
				if (!launch_mix(gen_for_kind[a.relay_kind[p]], t, stream)) {
					return false;
				}
			}
			return true;
		}
		// the scan kernels re-associate; the CPU build runs the exact single-effect pass in their place
#define OALSFX_SCX(id, CT) if (kernel_id == id) { kernel_id = kGenEqualizer; }
		OALSFX_SCAN_TABLE(OALSFX_SCX)
#undef OALSFX_SCX
		int span_twin = -1;
#define OALSFX_SX(id, CT, SL, CHAIN) if (kernel_id == id) { span_twin = (CHAIN ? kChainStereo : CT == 1 ? kReverbMono : kReverbStereo); }
		OALSFX_SPAN_TABLE(OALSFX_SX)
#undef OALSFX_SX
		bool bulk = false;
#define OALSFX_BX(id, CT, CHAIN) if (kernel_id == id) { span_twin = (CHAIN ? kChainStereo : CT == 1 ? kReverbMono : kReverbStereo); bulk = true; }
		OALSFX_SPAN_BULK_TABLE(OALSFX_BX)
#undef OALSFX_BX
		if (span_twin >= 0) {
			static int launch_parity = 0;
			const bool late_stores = bulk && (++launch_parity & 1);
			for (int w = 0; w < a.tile_count; ++w) {
				const int tile = a.tiles ? static_cast<int>(a.tiles[w].tile) : a.tile_first + w;
				// the device decides per CTA (= per tile, or per share of a tile); per tile here
				bool steady = true;
				int32_t off0[2] = {0, 0};
				for (int lane = 0; lane < kLanes && steady; ++lane) {
					if (tile * kLanes + lane < a.num_streams) {
						int32_t off[2] = {0, 0};
						auto probe = [&](auto cx) {
							const bool ok = cx.setup(a, tile, lane);
							off[0] = cx.rev_off;
							off[1] = span_twin == kChainStereo ? cx.echo_off : 0;
							return ok;
						};
						steady = span_twin == kChainStereo ? probe(span::Context<2, true>()) :
							span_twin == kReverbMono ? probe(span::Context<1, false>()) : probe(span::Context<2, false>());
						if (lane == 0) {
							off0[0] = off[0];
							off0[1] = off[1];
						}
						// the bulk kernel moves whole rows: every stream of the tile at the same ring positions
						steady = steady && (!bulk || (off[0] == off0[0] && off[1] == off0[1]));
					}
				}
				for (int lane = 0; lane < kLanes; ++lane) {
					if (tile * kLanes + lane >= a.num_streams) {
						continue;
					}
					if (!steady) {
						switch (span_twin) {
						case kChainStereo: mix_stream<2, false, FxEqualizer, FxModDelay, FxEcho, FxReverb>(a, tile, lane); break;
						case kReverbMono: mix_stream<1, false, FxReverb, FxNull, FxNull, FxNull>(a, tile, lane); break;
						default: mix_stream<2, false, FxReverb, FxNull, FxNull, FxNull>(a, tile, lane); break;
						}
					} else {
						++g_span_streams;
						if (bulk) {
							++g_span_bulk_streams;
							switch (span_twin) {
							case kChainStereo: span::emulate_stream_bulk<2, true>(a, tile, lane, late_stores); break;
							case kReverbMono: span::emulate_stream_bulk<1, false>(a, tile, lane, late_stores); break;
							default: span::emulate_stream_bulk<2, false>(a, tile, lane, late_stores); break;
							}
							continue;
						}
						switch (span_twin) {
						case kChainStereo: span::emulate_stream<2, true>(a, tile, lane); break;
						case kReverbMono: span::emulate_stream<1, false>(a, tile, lane); break;
						default: span::emulate_stream<2, false>(a, tile, lane); break;
						}
					}
				}
			}
			return true;
		}
		if (kernel_id >= kKernelCount && kernel_id < kTabDry) { // a duo / quartet kernel: the CPU build runs its thread-per-stream twin
			kernel_id = twin_of_fused(kernel_id);
		}
		for (int w = 0; w < a.tile_count; ++w) {
			int tile = a.tile_first + w;
			uint32_t mask = 0xFFFFFFFFU;
			if (a.tiles) {
				tile = static_cast<int>(a.tiles[w].tile);
				mask = a.tiles[w].mask;
			}
			for (int lane = 0; lane < kLanes; ++lane) {
				if (!((mask >> lane) & 1U) || tile * kLanes + lane >= a.num_streams) {
					continue;
				}
				switch (kernel_id) {
#define OALSFX_X(id, CT, SF, F0, F1, F2, F3) \
				case id: mix_stream<CT, SF, F0, F1, F2, F3>(a, tile, lane); break;
					OALSFX_KERNEL_TABLE(OALSFX_X)
#undef OALSFX_X
#define OALSFX_TBX(id, Fx, kind) \
				case id: mix_stream<0, true, Fx, FxNull, FxNull, FxNull, true>(a, tile, lane); break;
					OALSFX_TABMODE_TABLE(OALSFX_TBX)
#undef OALSFX_TBX
				default:
					error_ = "unknown kernel id";
					return false;
				}
			}
		}
		return true;
	}
	bool reduce_bus(const float* data, long long ts, long long ls, long long fs, long long cs,
		int num_streams, int frames, int channels, float* bus, void*) override
	{
		for (int f = 0; f < frames; ++f) {
			for (int c = 0; c < channels; ++c) {
				double sum = 0.0;
				for (int s = 0; s < num_streams; ++s) {
					sum += data[(s / kLanes) * ts + (s % kLanes) * ls + f * fs + c * cs];
				}
				bus[f * channels + c] = static_cast<float>(sum);
			}
		}
		return true;
	}
	bool pcm_to_float(const void* src, int bits, float* dst, long long count, void*) override
	{
		if (bits != 8 && bits != 16) {
			error_ = "Invalid bit depth.";
			return false;
		}
		for (long long i = 0; i < count; ++i) {
			dst[i] = bits == 16 ? static_cast<float>(static_cast<const int16_t*>(src)[i]) / 32768.0F :
				static_cast<float>(static_cast<int>(static_cast<const uint8_t*>(src)[i]) - 128) / 128.0F;
		}
		return true;
	}
	bool debug_waveshaper(const float* samples, float edge_coeff, float* out, long long count, void*) override
	{
		for (long long i = 0; i < count; i += 4) {
			float smp[4];
			for (int k = 0; k < 4; ++k) {
				smp[k] = samples[i + k < count ? i + k : count - 1];
			}
			oalsfx::FxDistortion::shape(smp, edge_coeff);
			for (int k = 0; k < 4 && i + k < count; ++k) {
				out[i + k] = smp[k];
			}
		}
		return true;
	}
	bool float_to_s16(const float* src, int16_t* dst, int rows, long long row_len, float* row_scale, void*) override
	{
		for (int r = 0; r < rows; ++r) {
			const float* row = src + static_cast<long long>(r) * row_len;
			float lo = -1.0F, hi = 1.0F;
			for (long long i = 0; i < row_len; ++i) {
				lo = row[i] < lo ? row[i] : lo;
				hi = row[i] > hi ? row[i] : hi;
			}
			const float scale = 1.0F / (hi > -lo ? hi : -lo);
			if (row_scale) {
				row_scale[r] = scale;
			}
			for (long long i = 0; i < row_len; ++i) {
				dst[static_cast<long long>(r) * row_len + i] = static_cast<int16_t>(scale * row[i] * 32767.0F);
			}
		}
		return true;
	}
	bool sync(void*) override { return true; }
	void* stream_create() override { return this; } // everything is synchronous here
	void stream_destroy(void*) override {}
	bool stream_wait(void*, void*) override { return true; }
	const std::string& error() const override { return error_; }

private:
	std::string error_;
};

} // namespace

Backend* make_backend(int, std::string&) { return new HostBackend; }
extern "C" long long oalsfx_emu_span_streams() { return g_span_streams; }
extern "C" long long oalsfx_emu_span_bulk_streams() { return g_span_bulk_streams; }
const char* backend_build_info() { return "oalsfx host-emu (tests only)"; }

} // namespace oalsfx
