// engine.cpp -- host logic of the batched engine behind include/oalsfx_engine.h.
//
// What it does (replacing the per-instance bookkeeping of the reference's Api::Impl,
// oalsfxpp.cpp:2820-3432, for thousands of streams at once):
//   * keeps per stream and slot an effect "class" (effect type + normalized properties) and per
//     stream a send class; classes are de-duplicated, their coefficient blocks derived once on the
//     host (derive.cpp);
//   * owns the HBM arenas: per slot a lane-interleaved ring region sized by the largest effect type
//     in that slot, per slot a state region, one send-filter state region;
//   * at mix time groups streams by (slot classes, send class, pending-update bits) and launches one
//     fused kernel per group and <= 2048-frame block, with the group's coefficient blocks as kernel
//     arguments; signatures without a fused kernel run as a chain of single-effect passes that
//     accumulate onto the bus in slot order (same summation order as the reference).
//
// No CUDA types here: device services come through backend.h.
#include "oalsfx_engine.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "backend.h"
#include "derive.h"
#include "kernel_table.h"
#include "span.cuh"

using namespace oalsfx;

namespace {

std::string g_create_error;

struct FxClass {
	int type = kFxNull;
	oalsfxpp::EffectProps props;
	SlotCoef coef;          // device table pointers already patched in
	bool used = false;      // scratch for compaction
};

struct SendClass {
	SendSettings direct;
	SendSettings aux[kMaxSlots];
	SendCoef direct_coef;
	SendCoef aux_coef[kMaxSlots];
};

struct GroupKey {
	int fx[kMaxSlots];
	int send;
	uint32_t pending;
	bool operator<(const GroupKey& o) const
	{
		return std::memcmp(this, &o, sizeof(GroupKey)) < 0;
	}
	bool operator==(const GroupKey& o) const { return std::memcmp(this, &o, sizeof(GroupKey)) == 0; }
};

struct Group {
	GroupKey key;
	bool identity = false;     // covers tiles 0..T-1 with all lanes
	bool full_tiles = false;   // every listed tile takes part with all of its (existing) lanes
	bool table = false;        // table mode: key.fx[] holds effect KINDS, streams differ in parameters
	bool multi = false;        // class-per-tile launch: every tile has ONE parameter class, coefficient blocks per tile from HBM
	size_t tile_begin = 0;     // into the concatenated tile list
	int tile_count = 0;
};

} // namespace

struct oalsfx_engine {
	oalsfx_engine_desc desc{};
	DeviceLayout dev;
	Backend* be = nullptr;
	int streams = 0, tiles = 0, channels = 0, slots = 0;

	// per-stream configuration (host)
	std::vector<int> fx_class[kMaxSlots];
	std::vector<int> send_class;
	std::vector<uint8_t> pending;       // bit s: slot s has an un-consumed `update`
	// Frames mixed since a launch last carried an update (or the engine was created / restored): a hint for the
	// kernels that only run streams in the steady state (span.cuh) -- a reverb's tap cross-fade takes 128 frames
	// from its update on.  The device checks the real per-stream state; this only avoids launches that would fall back.
	long long frames_since_update = 0;

	std::vector<FxClass> classes;
	std::unordered_map<std::string, int> class_index;
	std::vector<SendClass> send_classes;
	struct Table { void* p; size_t bytes; };
	std::map<std::string, Table> tables; // device lookup tables by content key (released when no class refers to them)
	size_t table_bytes = 0;
	// The stream the caller last handed to a mix / bus / PCM call: work may still be running there (a non-blocking
	// stream does not order against the legacy NULL stream), so everything that mutates, reallocates, uploads or
	// reads back engine state waits for it first (quiesce).
	void* last_stream = nullptr;
	bool stream_busy = false;
	void note_stream(void* stream)
	{
		last_stream = stream;
		stream_busy = true;
	}
	bool quiesce()
	{
		bool ok = true;
		if (stream_busy) {
			ok = be->sync(last_stream);
			stream_busy = false;
		}
		return be->sync(nullptr) && ok;
	}

	// device arenas
	float* ring[kMaxSlots] = {};
	int ring_cap[kMaxSlots] = {};        // words per lane
	uint32_t* slot_state[kMaxSlots] = {};
	uint32_t* send_state = nullptr;
	TileRef* tile_list = nullptr;
	size_t tile_list_cap = 0;
	float* stage_in = nullptr;
	float* stage_out = nullptr;
	size_t stage_cap = 0;                // floats
	long long device_bytes = 0;

	// Table mode (many parameter sets per engine): coefficient blocks of every class and every stream's
	// class indices in HBM, read by the kTab* kernels.  Rebuilt with the groups.
	SlotCoef* slot_table_dev = nullptr;
	size_t slot_table_cap = 0;
	SendCoef* send_table_dev = nullptr;
	size_t send_table_cap = 0;
	int32_t* lane_class_dev[kMaxSlots] = {};
	int32_t* lane_send_dev = nullptr;

	std::vector<Group> groups;
	// Class-per-tile launch (duo_multi_kernel): the class table and every tile's class index in HBM; `groups` is
	// then the single multi group and `fallback_groups` holds the per-class / table-mode groups for the blocks the
	// fused kernel cannot take (frames < 2).
	std::vector<Group> fallback_groups;
	MixClassEntry* class_table_dev = nullptr;
	size_t class_table_cap = 0;
	int32_t* tile_class_dev = nullptr;
	bool groups_dirty = true;
	long long launches = 0;
	int last_kernel = -1;               // id of the most recent mix kernel launched (oalsfx_engine_last_kernel)
	// Host buffers the engine has page-locked in place (oalsfx_engine_pin_host, or on first sight with OALSFX_PIN_HOST=1)
	struct PinnedHost { void* p; size_t bytes; bool automatic; };
	std::vector<PinnedHost> pinned_host;
	bool auto_pin_host = false;
	bool pin_host(void* p, size_t bytes, bool automatic)
	{
		for (PinnedHost& h : pinned_host) {
			if (h.p == p && h.bytes >= bytes) {
				return true;
			}
		}
		// a new size or place: automatic registrations that overlap it are stale (the caller's buffer moved)
		for (size_t i = 0; i < pinned_host.size();) {
			const char* a0 = static_cast<const char*>(pinned_host[i].p);
			const char* b0 = static_cast<const char*>(p);
			if (pinned_host[i].automatic && a0 < b0 + bytes && b0 < a0 + pinned_host[i].bytes) {
				be->host_unregister(pinned_host[i].p);
				pinned_host.erase(pinned_host.begin() + static_cast<long>(i));
			} else {
				++i;
			}
		}
		if (automatic && pinned_host.size() >= 16) { // keep the table small: forget the oldest automatic entry
			for (size_t i = 0; i < pinned_host.size(); ++i) {
				if (pinned_host[i].automatic) {
					be->host_unregister(pinned_host[i].p);
					pinned_host.erase(pinned_host.begin() + static_cast<long>(i));
					break;
				}
			}
		}
		if (!be->host_register(p, bytes)) {
			return false;
		}
		pinned_host.push_back(PinnedHost{p, bytes, automatic});
		return true;
	}

	bool mix_launch(int id, const MixArgs& a, void* stream)
	{
		last_kernel = id;
		return be->launch_mix(id, a, stream);
	}
	// Which fused kernel family serves whole-tile groups: 2 = automatic (default): the two-stage duo kernel,
	// or the four-stage quartet pipeline when the group has too few tiles to fill the GPU (and for
	// signatures that only have a quartet entry); 3 = quartet wherever it exists; 4 = duo wherever it exists;
	// 0 = the plain thread-per-stream twin, 5 = the relay pipeline wherever eligible.
	// 6 = as automatic (names the span kernel's tests).  OALSFX_KERNEL=auto|quartet|duo|single|relay|span overrides
	// (A/B measurements and the parity tests of every family).
	int family = 2;
	bool scan_enabled = false;          // OALSFX_SCAN=1
	int family_span_bulk = 1;           // OALSFX_SPAN_BULK=0: whole-tile spans stay on the load / store kernel; 2: whole tiles (and the bulk kernel) however few (A/B, tests)
	// Host-buffer mix: tile slice the next launches are restricted to (0 = all tiles), and the
	// engine-owned streams that overlap H2D, kernels and D2H of consecutive slices.
	int slice_first = 0, slice_count = 0;
	void* pipe_in = nullptr;
	void* pipe_run = nullptr;
	void* pipe_out = nullptr;
	std::string error;

	~oalsfx_engine()
	{
		if (!be) {
			return;
		}
		be->sync(nullptr);
		for (int s = 0; s < kMaxSlots; ++s) {
			be->release(ring[s]);
			be->release(slot_state[s]);
		}
		be->release(send_state);
		be->release(tile_list);
		be->release(class_table_dev);
		be->release(tile_class_dev);
		be->release(slot_table_dev);
		be->release(send_table_dev);
		be->release(lane_send_dev);
		for (int s = 0; s < kMaxSlots; ++s) {
			be->release(lane_class_dev[s]);
		}
		for (const PinnedHost& h : pinned_host) {
			be->host_unregister(h.p);
		}
		be->release(stage_in);
		be->release(stage_out);
		be->stream_destroy(pipe_in);
		be->stream_destroy(pipe_run);
		be->stream_destroy(pipe_out);
		for (auto& kv : tables) {
			be->release(kv.second.p);
		}
		delete be;
	}

	int fail(int code, const std::string& msg)
	{
		error = msg;
		return code;
	}

	void* dev_alloc(size_t bytes)
	{
		void* p = be->alloc(bytes);
		if (p) {
			device_bytes += static_cast<long long>(bytes);
		}
		return p;
	}

	// ---- classes --------------------------------------------------------------------------------
	void* table_for(const std::string& key, const void* data, size_t bytes)
	{
		auto it = tables.find(key);
		if (it != tables.end()) {
			return it->second.p;
		}
		void* p = dev_alloc(bytes);
		if (!p || !be->upload(p, data, bytes, nullptr) || !be->sync(nullptr)) {
			return nullptr;
		}
		tables.emplace(key, Table{p, bytes});
		table_bytes += bytes;
		return p;
	}

	int class_for(int type, const oalsfxpp::EffectProps& props_in)
	{
		oalsfxpp::Effect fx;
		std::memset(&fx, 0, sizeof(fx));
		fx.type_ = static_cast<oalsfxpp::EffectType>(type);
		if (type != kFxNull) {
			fx.props_ = props_in;
		}
		fx.normalize();
		// Key = type + the bytes of the property block that type uses.
		size_t used = 0;
		switch (type) {
		case kFxChorus: used = sizeof(fx.props_.chorus_); break;
		case kFxCompressor: used = sizeof(fx.props_.compressor_); break;
		case kFxDedicatedDialog: case kFxDedicatedLfe: used = sizeof(fx.props_.dedicated_); break;
		case kFxDistortion: used = sizeof(fx.props_.distortion_); break;
		case kFxEcho: used = sizeof(fx.props_.echo_); break;
		case kFxEqualizer: used = sizeof(fx.props_.equalizer_); break;
		case kFxFlanger: used = sizeof(fx.props_.flanger_); break;
		case kFxRingModulator: used = sizeof(fx.props_.ring_modulator_); break;
		case kFxReverb: case kFxEaxReverb: used = offsetof(oalsfxpp::EffectProps::Reverb, decay_hf_limit_) + 1; break;
		default: break;
		}
		std::string key(reinterpret_cast<const char*>(&type), sizeof(type));
		key.append(reinterpret_cast<const char*>(&fx.props_), used);
		auto it = class_index.find(key);
		if (it != class_index.end()) {
			return it->second;
		}
		FxClass c;
		c.type = type;
		c.props = fx.props_;
		SlotTables t;
		derive_slot(dev, desc.sampling_rate, type, fx.props_, c.coef, t);
		if (!t.sin_delays.empty()) {
			const ModDelayCoef& m = c.coef.u.mod_delay;
			std::string tk = "sd";
			tk.append(reinterpret_cast<const char*>(&m.lfo_range), 4);
			tk.append(reinterpret_cast<const char*>(&m.lfo_scale), 4);
			tk.append(reinterpret_cast<const char*>(&m.depth), 4);
			tk.append(reinterpret_cast<const char*>(&m.delay), 4);
			void* p = table_for(tk, t.sin_delays.data(), t.sin_delays.size() * sizeof(int32_t));
			if (!p) {
				return -1;
			}
			c.coef.u.mod_delay.sin_delays = static_cast<const int32_t*>(p);
		}
		if (!t.mod_sinus.empty()) {
			std::string tk = "ms";
			tk.append(reinterpret_cast<const char*>(&c.coef.u.reverb.mod_range), 4);
			void* p = table_for(tk, t.mod_sinus.data(), t.mod_sinus.size() * sizeof(float));
			if (!p) {
				return -1;
			}
			c.coef.u.reverb.mod_sinus = static_cast<const float*>(p);
		}
		classes.push_back(c);
		const int id = static_cast<int>(classes.size()) - 1;
		class_index.emplace(key, id);
		return id;
	}

	// Drop classes no stream refers to any more (cfg2-style per-block parameter changes would
	// otherwise grow the table without bound).  Class 0 (null) always stays.
	void compact_classes()
	{
		for (auto& c : classes) {
			c.used = false;
		}
		classes[0].used = true;
		for (int s = 0; s < kMaxSlots; ++s) {
			for (int id : fx_class[s]) {
				classes[static_cast<size_t>(id)].used = true;
			}
		}
		std::vector<int> remap(classes.size(), -1);
		std::vector<FxClass> kept;
		for (size_t i = 0; i < classes.size(); ++i) {
			if (classes[i].used) {
				remap[i] = static_cast<int>(kept.size());
				kept.push_back(classes[i]);
			}
		}
		for (int s = 0; s < kMaxSlots; ++s) {
			for (int& id : fx_class[s]) {
				id = remap[static_cast<size_t>(id)];
			}
		}
		for (auto it = class_index.begin(); it != class_index.end();) {
			const int to = remap[static_cast<size_t>(it->second)];
			if (to < 0) {
				it = class_index.erase(it);
			} else {
				it->second = to;
				++it;
			}
		}
		classes.swap(kept);
		// lookup tables no surviving class points at: nothing may still be reading them
		std::vector<const void*> live;
		for (const FxClass& c : classes) {
			const int kind = kind_of_type(c.type);
			if (kind == kKindModDelay && c.coef.u.mod_delay.sin_delays) {
				live.push_back(c.coef.u.mod_delay.sin_delays);
			} else if (kind == kKindReverb && c.coef.u.reverb.mod_sinus) {
				live.push_back(c.coef.u.reverb.mod_sinus);
			}
		}
		bool quiet = false;
		for (auto it = tables.begin(); it != tables.end();) {
			if (std::find(live.begin(), live.end(), it->second.p) == live.end()) {
				if (!quiet) {
					quiesce();
					quiet = true;
				}
				be->release(it->second.p);
				device_bytes -= static_cast<long long>(it->second.bytes);
				table_bytes -= it->second.bytes;
				it = tables.erase(it);
			} else {
				++it;
			}
		}
		// send classes no stream uses any more
		std::vector<int> send_remap(send_classes.size(), -1);
		std::vector<SendClass> send_kept;
		for (int id : send_class) {
			if (send_remap[static_cast<size_t>(id)] < 0) {
				send_remap[static_cast<size_t>(id)] = static_cast<int>(send_kept.size());
				send_kept.push_back(send_classes[static_cast<size_t>(id)]);
			}
		}
		for (int& id : send_class) {
			id = send_remap[static_cast<size_t>(id)];
		}
		send_classes.swap(send_kept);
		compacted_at_bytes = table_bytes;
	}
	size_t compacted_at_bytes = 0;

	int send_class_for(const SendSettings& direct, const SendSettings* aux)
	{
		for (size_t i = 0; i < send_classes.size(); ++i) {
			const SendClass& sc = send_classes[i];
			if (std::memcmp(&sc.direct, &direct, sizeof(direct)) == 0 &&
				std::memcmp(sc.aux, aux, sizeof(SendSettings) * static_cast<size_t>(slots)) == 0) {
				return static_cast<int>(i);
			}
		}
		SendClass sc;
		std::memset(&sc, 0, sizeof(sc));
		sc.direct = direct;
		for (int i = 0; i < slots; ++i) {
			sc.aux[i] = aux[i];
		}
		derive_sends(dev, desc.sampling_rate, slots, sc.direct, sc.aux, sc.direct_coef, sc.aux_coef);
		send_classes.push_back(sc);
		return static_cast<int>(send_classes.size()) - 1;
	}

	// ---- arenas -------------------------------------------------------------------------------------
	bool ensure_ring(int slot, int words)
	{
		if (words <= ring_cap[slot]) {
			return true;
		}
		const size_t new_tile_bytes = static_cast<size_t>(words) * kLanes * sizeof(float);
		float* fresh = static_cast<float*>(dev_alloc(new_tile_bytes * static_cast<size_t>(tiles)));
		if (!fresh) {
			return false;
		}
		if (ring[slot]) {
			const size_t old_tile_bytes = static_cast<size_t>(ring_cap[slot]) * kLanes * sizeof(float);
			if (!be->copy_2d(fresh, new_tile_bytes, ring[slot], old_tile_bytes, old_tile_bytes,
					static_cast<size_t>(tiles), nullptr) || !be->sync(nullptr)) {
				return false;
			}
			device_bytes -= static_cast<long long>(old_tile_bytes) * tiles;
			be->release(ring[slot]);
		}
		ring[slot] = fresh;
		ring_cap[slot] = words;
		return true;
	}

	// (tile, mask) list of a stream range
	static void tiles_of_range(int first, int n, std::vector<TileRef>& out)
	{
		out.clear();
		const int last = first + n - 1;
		for (int t = first / kLanes; t <= last / kLanes; ++t) {
			const int lo = std::max(first, t * kLanes) - t * kLanes;
			const int hi = std::min(last, t * kLanes + kLanes - 1) - t * kLanes;
			const uint32_t mask = (hi - lo == 31 ? 0xFFFFFFFFU : ((1U << (hi - lo + 1)) - 1U) << lo);
			out.push_back(TileRef{static_cast<uint32_t>(t), mask});
		}
	}

	// ---- grouping -----------------------------------------------------------------------------------
	GroupKey key_of(int s) const
	{
		GroupKey k;
		std::memset(&k, 0, sizeof(k));
		for (int i = 0; i < kMaxSlots; ++i) {
			k.fx[i] = fx_class[i][static_cast<size_t>(s)];
		}
		k.send = send_class[static_cast<size_t>(s)];
		k.pending = pending[static_cast<size_t>(s)];
		return k;
	}

	template <class T>
	bool grow(T*& p, size_t& cap, size_t need)
	{
		if (need <= cap) {
			return true;
		}
		be->sync(nullptr);
		if (p) {
			device_bytes -= static_cast<long long>(cap * sizeof(T));
			be->release(p);
		}
		cap = need * 2;
		p = static_cast<T*>(dev_alloc(cap * sizeof(T)));
		return p != nullptr;
	}

	// Table mode: every class's coefficient block, every send class's five send blocks and every stream's
	// class indices go to HBM (~0.7 KB per class, 20 bytes per stream).
	bool upload_tables()
	{
		std::vector<SlotCoef> coefs(classes.size());
		for (size_t i = 0; i < classes.size(); ++i) {
			coefs[i] = classes[i].coef;
		}
		std::vector<SendCoef> sends(send_classes.size() * kSendCount);
		for (size_t i = 0; i < send_classes.size(); ++i) {
			sends[i * kSendCount] = send_classes[i].direct_coef;
			for (int k = 0; k < kMaxSlots; ++k) {
				sends[i * kSendCount + 1 + static_cast<size_t>(k)] = send_classes[i].aux_coef[k];
			}
		}
		const size_t padded = static_cast<size_t>(tiles) * kLanes;
		if (!grow(slot_table_dev, slot_table_cap, coefs.size()) || !grow(send_table_dev, send_table_cap, sends.size())) {
			return false;
		}
		if (!lane_send_dev && !(lane_send_dev = static_cast<int32_t*>(dev_alloc(padded * sizeof(int32_t))))) {
			return false;
		}
		for (int s = 0; s < kMaxSlots; ++s) {
			if (!lane_class_dev[s] && !(lane_class_dev[s] = static_cast<int32_t*>(dev_alloc(padded * sizeof(int32_t))))) {
				return false;
			}
		}
		bool ok = be->upload(slot_table_dev, coefs.data(), coefs.size() * sizeof(SlotCoef), nullptr) &&
			be->upload(send_table_dev, sends.data(), sends.size() * sizeof(SendCoef), nullptr);
		std::vector<int32_t> idx(padded, 0);
		for (int s = 0; s < streams; ++s) {
			idx[static_cast<size_t>(s)] = send_class[static_cast<size_t>(s)];
		}
		ok = ok && be->upload(lane_send_dev, idx.data(), padded * sizeof(int32_t), nullptr) && be->sync(nullptr);
		for (int k = 0; k < kMaxSlots && ok; ++k) {
			std::fill(idx.begin(), idx.end(), 0);
			for (int s = 0; s < streams; ++s) {
				idx[static_cast<size_t>(s)] = fx_class[k][static_cast<size_t>(s)];
			}
			ok = be->upload(lane_class_dev[k], idx.data(), padded * sizeof(int32_t), nullptr) && be->sync(nullptr);
		}
		return ok;
	}

	bool rebuild_groups()
	{
		// parameter automation leaves unused classes, send classes and lookup tables behind: compact by count and by bytes
		if (classes.size() > 256 || send_classes.size() > 64 || table_bytes > compacted_at_bytes + (size_t(32) << 20)) {
			compact_classes();
		}
		quiesce();  // the group tables below are uploaded in place
		groups.clear();
		const GroupKey k0 = key_of(0);
		bool uniform = true;
		for (int s = 1; s < streams && uniform; ++s) {
			uniform = key_of(s) == k0;
		}
		if (uniform) {
			Group g;
			g.key = k0;
			g.identity = true;
			g.tile_count = tiles;
			groups.push_back(g);
			fallback_groups.clear();
			groups_dirty = false;
			return true;
		}
		std::map<GroupKey, std::vector<TileRef>> by_key;
		for (int s = 0; s < streams; ++s) {
			std::vector<TileRef>& v = by_key[key_of(s)];
			const uint32_t tile = static_cast<uint32_t>(s / kLanes);
			if (v.empty() || v.back().tile != tile) {
				v.push_back(TileRef{tile, 0});
			}
			v.back().mask |= 1U << (s % kLanes);
		}
		const std::map<GroupKey, std::vector<TileRef>> full_keys = by_key;
		// One launch per parameter class is fine for a handful of classes (each runs the fused kernels with
		// constant-bank coefficients).  Beyond that the launches multiply -- every class touches most tiles --
		// so the streams are grouped by effect KINDS only and the kTab* kernels fetch each stream's own
		// coefficient block from HBM.
		constexpr size_t kTableModeMinGroups = 8;
		const bool table_mode = by_key.size() > kTableModeMinGroups;
		if (table_mode) {
			by_key.clear();
			for (int s = 0; s < streams; ++s) {
				GroupKey k = key_of(s);
				for (int i = 0; i < kMaxSlots; ++i) {
					k.fx[i] = kind_of_type(classes[static_cast<size_t>(k.fx[i])].type);
				}
				k.send = 0;
				std::vector<TileRef>& v = by_key[k];
				const uint32_t tile = static_cast<uint32_t>(s / kLanes);
				if (v.empty() || v.back().tile != tile) {
					v.push_back(TileRef{tile, 0});
				}
				v.back().mask |= 1U << (s % kLanes);
			}
			if (!upload_tables()) {
				return false;
			}
		}
		std::vector<TileRef> all;
		for (auto& kv : by_key) {
			Group g;
			g.key = kv.first;
			g.table = table_mode;
			g.tile_begin = all.size();
			g.tile_count = static_cast<int>(kv.second.size());
			g.full_tiles = true;
			for (const TileRef& t : kv.second) {
				const int lanes_here = std::min(kLanes, streams - static_cast<int>(t.tile) * kLanes);
				const uint32_t want = (lanes_here == kLanes ? 0xFFFFFFFFU : ((1U << lanes_here) - 1U));
				g.full_tiles = g.full_tiles && t.mask == want;
			}
			all.insert(all.end(), kv.second.begin(), kv.second.end());
			groups.push_back(g);
		}
		if (all.size() > tile_list_cap) {
			be->sync(nullptr);
			be->release(tile_list);
			tile_list_cap = all.size() * 2;
			tile_list = static_cast<TileRef*>(dev_alloc(tile_list_cap * sizeof(TileRef)));
			if (!tile_list) {
				return false;
			}
		}
		if (!be->upload(tile_list, all.data(), all.size() * sizeof(TileRef), nullptr) || !be->sync(nullptr)) {
			return false;
		}
		if (!build_multi(full_keys)) {
			return false;
		}
		groups_dirty = false;
		return true;
	}

	// Class-per-tile launch: possible when every tile holds ONE parameter class (presets assigned in runs of 32
	// streams), all classes share one slot signature that has a *_multi kernel (the fused stereo chain, or the relay
	// pipeline for anything else with one or two channels) and nothing needs the exact kernels.  One launch then
	// covers all classes at the fused kernel's speed.
	int multi_kernel = -1;
	int multi_kinds[kMaxSlots] = {};   // per slot
	bool build_multi(const std::map<GroupKey, std::vector<TileRef>>& by_key)
	{
		fallback_groups.clear();
		multi_kernel = -1;
		constexpr size_t kMultiMinGroups = 3; // one or two big classes: a constant-bank launch each is faster
		if (by_key.size() < kMultiMinGroups || !be->has_relay() || (channels != 1 && channels != 2) || (family != 2 && family != 4)) {
			return true;
		}
		std::vector<int32_t> tile_class(static_cast<size_t>(tiles), -1);
		std::vector<MixClassEntry> table;
		table.reserve(by_key.size());
		uint32_t pending_any = 0;
		const KernelInfo* infos = kernel_infos();
		bool first = true, heavy = false, chain = false;
		int active = 0;
		for (const auto& kv : by_key) {
			const GroupKey& key = kv.first;
			const SendClass& sc = send_classes[static_cast<size_t>(key.send)];
			bool ok = sc.direct_coef.filter_type == 0;
			int kinds[kMaxSlots];
			for (int s = 0; s < kMaxSlots; ++s) {
				const FxClass& fc = classes[static_cast<size_t>(key.fx[s])];
				kinds[s] = kind_of_type(fc.type);
				ok = ok && (kinds[s] == kKindNull || sc.aux_coef[s].filter_type == 0) && !(fc.coef.flags & kCoefUnstable);
			}
			if (first) {
				std::memcpy(multi_kinds, kinds, sizeof(kinds));
				chain = channels == 2 && std::memcmp(infos[kChainStereo].kind, kinds, sizeof(kinds)) == 0;
				for (int s = 0; s < kMaxSlots; ++s) {
					active += kinds[s] != kKindNull ? 1 : 0;
					heavy = heavy || kinds[s] == kKindReverb;
				}
				first = false;
			}
			ok = ok && std::memcmp(multi_kinds, kinds, sizeof(kinds)) == 0 && active >= 1;
			for (const TileRef& t : kv.second) {
				const int lanes_here = std::min(kLanes, streams - static_cast<int>(t.tile) * kLanes);
				const uint32_t want = (lanes_here == kLanes ? 0xFFFFFFFFU : ((1U << lanes_here) - 1U));
				ok = ok && t.mask == want;
				tile_class[t.tile] = static_cast<int32_t>(table.size());
			}
			if (!ok) {
				return true;
			}
			Group g;
			g.key = key;
			MixArgs a;
			fill_common(a, g, 2, nullptr, nullptr, OALSFX_LAYOUT_STREAM_MAJOR, 2, 0);
			MixClassEntry e;
			std::memset(&e, 0, sizeof(e));
			int pos = 0;
			for (int s = 0; s < kMaxSlots; ++s) { // the chain kernel addresses slots by index, the relay by compacted position
				if (chain || kinds[s] != kKindNull) {
					fill_slot(a, g, pos, s, false);
					e.pending |= ((key.pending >> s) & 1U) << pos;
					++pos;
				}
			}
			sanitize_gains(a);
			std::memcpy(e.coefs, reinterpret_cast<const char*>(&a) + kMixCoefOffset, kMixCoefBytes);
			table.push_back(e);
			pending_any |= key.pending;
		}
		if (!grow(class_table_dev, class_table_cap, table.size())) {
			return false;
		}
		if (!tile_class_dev) {
			tile_class_dev = static_cast<int32_t*>(dev_alloc(static_cast<size_t>(tiles) * sizeof(int32_t)));
			if (!tile_class_dev) {
				return false;
			}
		}
		if (!be->upload(class_table_dev, table.data(), table.size() * sizeof(MixClassEntry), nullptr) ||
			!be->upload(tile_class_dev, tile_class.data(), tile_class.size() * sizeof(int32_t), nullptr) || !be->sync(nullptr)) {
			return false;
		}
		multi_kernel = chain ? kMultiChainStereo :
			channels == 1 ? (heavy ? kRelayMultiMonoHeavy : kRelayMultiMono) : (heavy ? kRelayMultiStereoHeavy : kRelayMultiStereo);
		fallback_groups.swap(groups);
		Group g;
		g.key = by_key.begin()->first;
		g.key.pending = pending_any;
		g.identity = true;
		g.multi = true;
		g.tile_count = tiles;
		groups.assign(1, g);
		return true;
	}

	// ---- launching ----------------------------------------------------------------------------------
	void fill_common(MixArgs& a, const Group& g, int frames, const float* src, float* dst, int layout,
		long long frames_total, long long frame0) const
	{
		std::memset(&a, 0, sizeof(a));
		a.frames = frames;
		a.channels = channels;
		a.num_streams = streams;
		if (layout == OALSFX_LAYOUT_TILED) {
			a.io_ts = frames_total * channels * kLanes;
			a.io_ls = 1;
			a.io_fs = static_cast<long long>(channels) * kLanes;
			a.io_cs = kLanes;
		} else {
			a.io_ts = frames_total * channels * kLanes;
			a.io_ls = frames_total * channels;
			a.io_fs = channels;
			a.io_cs = 1;
		}
		a.src = src + frame0 * a.io_fs;
		a.dst = dst + frame0 * a.io_fs;
		a.tiles = g.identity ? nullptr : tile_list + g.tile_begin;
		a.tile_count = g.tile_count;
		if (g.identity && slice_count > 0) { // host-buffer pipelining: only tiles [slice_first, +slice_count)
			a.tile_first = slice_first;
			a.tile_count = slice_count;
		}
		a.send_state = send_state;
		if (g.table) {
			a.slot_table = slot_table_dev;
			a.send_table = send_table_dev;
			a.lane_send = lane_send_dev;
			return;
		}
		const SendClass& sc = send_classes[static_cast<size_t>(g.key.send)];
		a.direct = sc.direct_coef;
	}

	void fill_slot(MixArgs& a, const Group& g, int pos, int slot, bool first_block) const
	{
		if (g.table) {
			a.lane_class[pos] = lane_class_dev[slot];
		} else {
			const FxClass& fc = classes[static_cast<size_t>(g.key.fx[slot])];
			const SendClass& sc = send_classes[static_cast<size_t>(g.key.send)];
			a.slot[pos] = fc.coef;
			a.aux[pos] = sc.aux_coef[slot];
		}
		a.aux_index[pos] = slot;
		a.ring[pos] = ring[slot];
		a.ring_tile_stride[pos] = static_cast<long long>(ring_cap[slot]) * kLanes;
		a.slot_state[pos] = slot_state[slot];
		if (first_block && ((g.key.pending >> slot) & 1U)) {
			a.update_mask |= 1U << pos;
		}
	}

	// Fast kernels add `sample * gain` unconditionally: make every static gain the reference would
	// skip (|g| <= 1e-5, oalsfxpp.cpp:56) an exact zero so the sum is unchanged.
	static void sanitize(float* gains, int n)
	{
		for (int i = 0; i < n; ++i) {
			if (!(std::fabs(gains[i]) > kSilenceGain)) {
				gains[i] = 0.0F;
			}
		}
	}

	static void sanitize_gains(MixArgs& a)
	{
		sanitize(&a.direct.gains[0][0], kMaxChannels * kMaxChannels);
		for (int p = 0; p < kMaxSlots; ++p) {
			sanitize(&a.aux[p].gains[0][0], kMaxChannels * kMaxChannels);
			SlotCoef& sc = a.slot[p];
			switch (kind_of_type(sc.type)) {
			case kKindModDelay: sanitize(&sc.u.mod_delay.gains[0][0], 2 * kMaxChannels); break;
			case kKindCompressor: sanitize(&sc.u.compressor.gains[0][0], kWetChannels * kMaxChannels); break;
			case kKindDedicated: sanitize(sc.u.dedicated.gains, kMaxChannels); break;
			case kKindDistortion: sanitize(sc.u.distortion.gains, kMaxChannels); break;
			case kKindEcho: sanitize(&sc.u.echo.gains[0][0], 2 * kMaxChannels); break;
			case kKindEqualizer: sanitize(&sc.u.equalizer.gains[0][0], kWetChannels * kMaxChannels); break;
			case kKindRingMod: sanitize(&sc.u.ring_mod.gains[0][0], kWetChannels * kMaxChannels); break;
			default: break; // reverb: its pan gains are running state, tested per sub-chunk on the device
			}
		}
	}

	bool launch_group(const Group& g, int frames, const float* src, float* dst, int layout,
		long long frames_total, long long frame0, bool first_block, void* stream)
	{
		if (g.multi) {
			MixArgs a;
			fill_common(a, g, frames, src, dst, layout, frames_total, frame0);
			a.with_dry = 1;
			const bool chain = multi_kernel == kMultiChainStereo;
			int pos = 0, floats = 0;
			for (int s = 0; s < kMaxSlots; ++s) {
				if (chain || multi_kinds[s] != kKindNull) {
					fill_slot(a, g, pos, s, false); // pointers and strides; the coefficient blocks come from the class table
					a.relay_kind[pos] = multi_kinds[s];
					a.relay_win[pos] = floats;
					floats += multi_kinds[s] == kKindReverb ? kPfWarpFloats :
						(multi_kinds[s] == kKindModDelay || multi_kinds[s] == kKindEcho) ? kFwWarpFloats : 0;
					++pos;
				}
			}
			a.relay_count = pos;
			a.relay_smem_floats = floats;
			a.update_mask = first_block ? 0xFFFFFFFFU : 0U;
			a.class_table = class_table_dev;
			a.tile_class = tile_class_dev;
			++launches;
			return mix_launch(multi_kernel, a, stream);
		}
		if (g.table) {
			// key.fx[] holds the kinds; one exact single-effect pass per slot, coefficients from the tables
			bool first = true;
			for (int s = 0; s < kMaxSlots; ++s) {
				if (g.key.fx[s] == kKindNull) {
					continue;
				}
				MixArgs a;
				fill_common(a, g, frames, src, dst, layout, frames_total, frame0);
				a.with_dry = first ? 1 : 0;
				a.accumulate = first ? 0 : 1;
				fill_slot(a, g, 0, s, first_block);
				++launches;
				if (!mix_launch(tab_kernel_for_kind(g.key.fx[s]), a, stream)) {
					return false;
				}
				first = false;
			}
			if (first) {
				MixArgs a;
				fill_common(a, g, frames, src, dst, layout, frames_total, frame0);
				a.with_dry = 1;
				fill_slot(a, g, 0, 0, first_block);
				++launches;
				return mix_launch(kTabDry, a, stream);
			}
			return true;
		}
		const SendClass& sc = send_classes[static_cast<size_t>(g.key.send)];
		int kinds[kMaxSlots];
		bool send_filter = sc.direct_coef.filter_type != 0;   // a send's shelf filter is active (apply_filters, oalsfxpp.cpp:3101-3143)
		bool unstable = false;
		for (int s = 0; s < kMaxSlots; ++s) {
			const int type = classes[static_cast<size_t>(g.key.fx[s])].type;
			kinds[s] = kind_of_type(type);
			if (kinds[s] != kKindNull && sc.aux_coef[s].filter_type != 0) {
				send_filter = true;
			}
			if (classes[static_cast<size_t>(g.key.fx[s])].coef.flags & kCoefUnstable) {
				unstable = true; // keeps the group on the exact (generic) kernels, see coefs.h
			}
		}
		const bool any_filter = send_filter || unstable;
		// The relay pipeline (relay.cuh): any signature of two or more effects in one launch, one warp per effect.
		const bool whole_tiles_group = g.identity || g.full_tiles;
		int active = 0;
		for (int s = 0; s < kMaxSlots; ++s) {
			active += kinds[s] != kKindNull ? 1 : 0;
		}
		// (with an active send filter the relay's filter-carrying variant is the only one-launch kernel: also for one slot)
		const bool relay_ok = be->has_relay() && !unstable && frames >= 2 && whole_tiles_group &&
			active >= (send_filter ? 1 : 2) && family != 0;
		auto launch_relay = [&]() {
			MixArgs a;
			fill_common(a, g, frames, src, dst, layout, frames_total, frame0);
			a.with_dry = 1;
			int pos = 0, floats = 0;
			bool heavy = false;
			for (int s = 0; s < kMaxSlots; ++s) {
				if (kinds[s] == kKindNull) {
					continue;
				}
				fill_slot(a, g, pos, s, first_block);
				a.relay_kind[pos] = kinds[s];
				a.relay_win[pos] = floats;
				floats += kinds[s] == kKindReverb ? kPfWarpFloats : (kinds[s] == kKindModDelay || kinds[s] == kKindEcho) ? kFwWarpFloats : 0;
				heavy = heavy || kinds[s] == kKindReverb;
				++pos;
			}
			a.relay_count = pos;
			a.relay_smem_floats = floats;
			sanitize_gains(a);
			++launches;
			// quad / 5.1 / 6.1 / 7.1: the kernels with a run-time channel count
			if (send_filter) {
				return mix_launch(channels == 1 ? (heavy ? kRelaySfMonoHeavy : kRelaySfMono) : channels == 2 ? (heavy ? kRelaySfStereoHeavy : kRelaySfStereo) :
					(heavy ? kRelaySfWideHeavy : kRelaySfWide), a, stream);
			}
			return mix_launch(channels == 1 ? (heavy ? kRelayMonoHeavy : kRelayMono) : channels == 2 ? (heavy ? kRelayStereoHeavy : kRelayStereo) :
				(heavy ? kRelayWideHeavy : kRelayWide), a, stream);
		};
		if (relay_ok && family == 5) { // OALSFX_KERNEL=relay: wherever eligible (A/B measurements, parity tests)
			return launch_relay();
		}
		// Opt-in (OALSFX_SCAN=1): a lone equalizer slot on few streams as a linear-recurrence scan over time (scan.cuh).
		// Re-associates the filter sums: within 1e-5 of the reference, not bit-exact -- hence never chosen by default.
		if (scan_enabled && be->has_relay() && active == 1 && !any_filter && g.identity && slice_count == 0 &&
			(channels == 1 || channels == 2) && frames >= 64 && streams <= 4096) {
			for (int s = 0; s < kMaxSlots; ++s) {
				if (kinds[s] == kKindEqualizer) {
					MixArgs a;
					fill_common(a, g, frames, src, dst, layout, frames_total, frame0);
					a.with_dry = 1;
					fill_slot(a, g, 0, s, first_block);
					sanitize_gains(a);
					++launches;
					return mix_launch(channels == 1 ? kScanEqualizerMono : kScanEqualizerStereo, a, stream);
				}
			}
		}
		// A fused single-pass kernel for this signature?
		const KernelInfo* infos = kernel_infos();
		for (int k = 0; k < kKernelCount; ++k) {
			const KernelInfo& ki = infos[k];
			if (ki.ct != channels || ki.sf || any_filter || frames < 2) {
				continue;
			}
			if (std::memcmp(ki.kind, kinds, sizeof(kinds)) != 0) {
				continue;
			}
			// A signature whose only fused kernel is the thread-per-stream one (cfg3's flanger + ring modulator +
			// distortion + compressor): below ~1000 tiles a stream's serial per-sample work bounds the block, and
			// the relay pipeline (a warp per effect) halves it -- B200, ms per 1024-frame block, mono 96 kHz:
			// 32 tiles 0.58 vs 1.08, 512 tiles 0.76 vs 1.40; 2048 tiles 2.34 vs 1.83 (throughput-bound: fewer
			// instructions win).
			constexpr int kRelayMaxTiles = 1024;
			if (relay_ok && family == 2 && duo_for_twin(ki.id) < 0 && quartet_for_twin(ki.id) < 0 &&
				(slice_count > 0 ? slice_count : g.tile_count) <= kRelayMaxTiles) {
				return launch_relay();
			}
			MixArgs a;
			fill_common(a, g, frames, src, dst, layout, frames_total, frame0);
			a.with_dry = 1;
			for (int s = 0; s < kMaxSlots; ++s) {
				if (kinds[s] != kKindNull) {
					fill_slot(a, g, s, s, first_block);
				}
			}
			++launches;
			sanitize_gains(a);
			const bool whole_tiles = g.identity || g.full_tiles;
			int id = ki.id;
			// Few tiles: block-parallel in time (span.cuh) for the single reverb slot and for the 4-slot chain, whenever no
			// update is pending and the delays allow a span (the device verifies the stream state and falls back).
			// A CTA fills an SM, so beyond one wave of the 148 SMs the kernel pays whole extra waves; few enough tiles are
			// shared among 2 or 4 CTAs each (16 / 8 streams per CTA).
			constexpr int kSpanMaxTiles = 296;
			const bool span_chain = ki.id == kChainStereo;
			if ((ki.id == kReverbMono || ki.id == kReverbStereo || span_chain) && whole_tiles && be->has_relay() && (family == 2 || family == 6) &&
				a.update_mask == 0 && a.tile_count <= kSpanMaxTiles && slice_count == 0 && frames_since_update >= 128) {
				const int share = family_span_bulk == 2 ? 0 : a.tile_count * 4 <= 148 ? 2 : a.tile_count * 2 <= 148 ? 1 : 0;
				// whole tiles per CTA: the ring traffic as bulk copies (span_bulk_kernel) if its timing is legal for this preset
				if (share == 0 && family_span_bulk) {
					const int tb = span::plan_bulk_frames(a, span_chain);
					if (tb > 0) {
						a.span_frames = tb;
						return mix_launch(span_chain ? kSpanBulkChainStereo : ki.id == kReverbMono ? kSpanBulkReverbMono : kSpanBulkReverbStereo, a, stream);
					}
				}
				const int t = span::plan_frames(a, span_chain, kLanes >> share);
				if (t > 0) {
					a.span_frames = t;
					return mix_launch((span_chain ? kSpanChainStereo : ki.id == kReverbMono ? kSpanReverbMono : kSpanReverbStereo) + share, a, stream);
				}
			}
			// Few tiles cannot fill the GPU with two warps each: below ~one tile per SM the four-stage pipeline
			// (twice the warps per tile) wins -- measured on B200, 4-slot stereo chain, ms per 1024-frame block:
			// 128 tiles 0.47 (quartet) vs 0.63 (duo); 256 tiles 0.64 vs 0.64; 2048 tiles 3.6 vs 3.07.
			constexpr int kPipelineMaxTiles = 160;
			// (Not for the tile slices of the host-buffer path: there the copies bound the step, and the duo
			// kernel measured 12.4 ms per step against 13.4 ms with quartet slices.)
			const bool few_tiles = a.tile_count <= kPipelineMaxTiles && slice_count == 0;
			if (whole_tiles && (family == 3 || (family == 2 && few_tiles)) && quartet_for_twin(ki.id) >= 0) {
				id = quartet_for_twin(ki.id);
			} else if (whole_tiles && family >= 2 && duo_for_twin(ki.id) >= 0) {
				id = duo_for_twin(ki.id);

			} else if (whole_tiles && family >= 2 && quartet_for_twin(ki.id) >= 0) {
				id = quartet_for_twin(ki.id); // no duo entry for this signature (single reverb slot)
			}
			return mix_launch(id, a, stream);
		}
		if (relay_ok) {
			return launch_relay();
		}
		// Chain of single-effect passes: the dry pass carries the first non-null slot.
		static const int gen_for_kind[] = {kGenDry, kGenModDelay, kGenCompressor, kGenDedicated, kGenDistortion,
			kGenEcho, kGenEqualizer, kGenRingMod, kGenReverb};
		bool first = true;
		for (int s = 0; s < kMaxSlots; ++s) {
			if (kinds[s] == kKindNull) {
				continue;
			}
			MixArgs a;
			fill_common(a, g, frames, src, dst, layout, frames_total, frame0);
			a.with_dry = first ? 1 : 0;
			a.accumulate = first ? 0 : 1;
			fill_slot(a, g, 0, s, first_block);
			++launches;
			if (!mix_launch(gen_for_kind[kinds[s]], a, stream)) {
				return false;
			}
			first = false;
		}
		if (first) { // no effect at all: dry only
			MixArgs a;
			fill_common(a, g, frames, src, dst, layout, frames_total, frame0);
			a.with_dry = 1;
			++launches;
			return mix_launch(kGenDry, a, stream);
		}
		return true;
	}
};

// =================================================================================================
extern "C" {

int oalsfx_engine_create(const oalsfx_engine_desc* desc, oalsfx_engine** out)
{
	if (!desc || !out) {
		g_create_error = "Null argument.";
		return OALSFX_ERR_ARGUMENT;
	}
	*out = nullptr;
	DeviceLayout dev;
	// Same checks, same order and same messages as Api::Impl::initialize (oalsfxpp.cpp:2853-2875).
	if (!make_device_layout(desc->channel_format, dev)) {
		g_create_error = "Invalid channel format.";
		return OALSFX_ERR_FORMAT;
	}
	if (desc->sampling_rate < 8000) {
		g_create_error = "Sampling rate is out of range.";
		return OALSFX_ERR_RATE;
	}
	if (desc->effect_count <= 0 || desc->effect_count > kMaxSlots) {
		g_create_error = "Effect count is out of range.";
		return OALSFX_ERR_EFFECTS;
	}
	if (desc->num_streams < 1) {
		g_create_error = "Stream count is out of range.";
		return OALSFX_ERR_ARGUMENT;
	}
	oalsfx_engine* e = new (std::nothrow) oalsfx_engine;
	if (!e) {
		g_create_error = "Failed to allocate the engine.";
		return OALSFX_ERR_MEMORY;
	}
	e->desc = *desc;
	e->dev = dev;
	e->streams = desc->num_streams;
	e->tiles = (desc->num_streams + kLanes - 1) / kLanes;
	e->channels = dev.channels;
	e->slots = desc->effect_count;
	if (const char* fam = std::getenv("OALSFX_KERNEL")) {
		e->family = (std::strcmp(fam, "single") == 0 ? 0 : std::strcmp(fam, "quartet") == 0 ? 3 : std::strcmp(fam, "duo") == 0 ? 4 : std::strcmp(fam, "relay") == 0 ? 5 : std::strcmp(fam, "span") == 0 ? 6 : 2);
	}
	if (const char* sc = std::getenv("OALSFX_SCAN")) {
		e->scan_enabled = std::atoi(sc) != 0;
	}
	if (const char* ph = std::getenv("OALSFX_PIN_HOST")) {
		e->auto_pin_host = std::atoi(ph) != 0;
	}
	if (const char* sb = std::getenv("OALSFX_SPAN_BULK")) {
		e->family_span_bulk = std::atoi(sb);
	}
	e->be = make_backend(desc->device, g_create_error);
	if (!e->be) {
		delete e;
		return OALSFX_ERR_DEVICE;
	}
	const size_t slot_bytes = static_cast<size_t>(e->tiles) * kSlotStateWords * kLanes * sizeof(uint32_t);
	for (int s = 0; s < e->slots; ++s) {
		e->slot_state[s] = static_cast<uint32_t*>(e->dev_alloc(slot_bytes));
		if (!e->slot_state[s]) {
			g_create_error = "Device allocation failed: " + e->be->error();
			delete e;
			return OALSFX_ERR_MEMORY;
		}
	}
	e->send_state = static_cast<uint32_t*>(
		e->dev_alloc(static_cast<size_t>(e->tiles) * kSendStateWords * kLanes * sizeof(uint32_t)));
	if (!e->send_state) {
		g_create_error = "Device allocation failed: " + e->be->error();
		delete e;
		return OALSFX_ERR_MEMORY;
	}
	// Every slot starts as the null effect, every send at its defaults (oalsfxpp.cpp:2877-2902).
	oalsfxpp::EffectProps none;
	std::memset(&none, 0, sizeof(none));
	e->class_for(kFxNull, none);
	for (int s = 0; s < kMaxSlots; ++s) {
		e->fx_class[s].assign(static_cast<size_t>(e->streams), 0);
	}
	SendSettings unit = {1.0F, 1.0F, 1.0F};
	SendSettings aux[kMaxSlots] = {unit, unit, unit, unit};
	e->send_class.assign(static_cast<size_t>(e->streams), e->send_class_for(unit, aux));
	e->pending.assign(static_cast<size_t>(e->streams), 0);
	*out = e;
	return OALSFX_OK;
}

void oalsfx_engine_destroy(oalsfx_engine* e) { delete e; }

int oalsfx_engine_set_effect(oalsfx_engine* e, int first_stream, int n_streams, int slot,
	int effect_type, const void* props, size_t props_bytes)
{
	if (!e) {
		return OALSFX_ERR_ARGUMENT;
	}
	if (first_stream < 0 || n_streams < 1 || first_stream + n_streams > e->streams) {
		return e->fail(OALSFX_ERR_ARGUMENT, "Stream range is out of range.");
	}
	if (slot < 0 || slot >= e->slots) {
		return e->fail(OALSFX_ERR_ARGUMENT, "Effect index is out of range.");
	}
	if (effect_type < 0 || effect_type >= kFxTypeCount) {
		return e->fail(OALSFX_ERR_ARGUMENT, "Unknown effect type.");
	}
	oalsfxpp::EffectProps p;
	std::memset(&p, 0, sizeof(p));
	if (effect_type != kFxNull) {
		if (!props || props_bytes != sizeof(oalsfxpp::EffectProps)) {
			return e->fail(OALSFX_ERR_ARGUMENT, "props must be the bytes of an oalsfxpp::EffectProps.");
		}
		std::memcpy(&p, props, sizeof(p));
	}
	const int id = e->class_for(effect_type, p);
	if (id < 0) {
		return e->fail(OALSFX_ERR_MEMORY, "Lookup table upload failed: " + e->be->error());
	}
	if (ring_words_for(effect_type, e->desc.sampling_rate) > e->ring_cap[slot]) {
		e->quiesce();  // the arena is about to be copied and freed: no mix may still be running on the caller's stream
	}
	if (!e->ensure_ring(slot, ring_words_for(effect_type, e->desc.sampling_rate))) {
		return e->fail(OALSFX_ERR_MEMORY, "Delay-line arena allocation failed: " + e->be->error());
	}
	// A type change constructs a fresh effect state (oalsfxpp.cpp:2692-2698): zero the slot state and
	// the rings of exactly those streams.
	std::vector<TileRef> reset;
	for (int s = first_stream; s < first_stream + n_streams; ++s) {
		int& cur = e->fx_class[slot][static_cast<size_t>(s)];
		if (e->classes[static_cast<size_t>(cur)].type != effect_type) {
			const uint32_t tile = static_cast<uint32_t>(s / kLanes);
			if (reset.empty() || reset.back().tile != tile) {
				reset.push_back(TileRef{tile, 0});
			}
			reset.back().mask |= 1U << (s % kLanes);
		}
		cur = id;
		e->pending[static_cast<size_t>(s)] |= static_cast<uint8_t>(1U << slot); // is_props_changed_ (oalsfxpp.cpp:2707)
	}
	if (!reset.empty()) {
		const int n = static_cast<int>(reset.size());
		e->quiesce();  // zeroing state a running mix may still be using
		if (!e->be->zero_lanes(e->slot_state[slot], static_cast<long long>(kSlotStateWords) * kLanes, kSlotStateWords,
				reset.data(), n, nullptr)) {
			return e->fail(OALSFX_ERR_DEVICE, e->be->error());
		}
		if (e->ring_cap[slot] > 0 && !e->be->zero_lanes(reinterpret_cast<uint32_t*>(e->ring[slot]),
				static_cast<long long>(e->ring_cap[slot]) * kLanes, e->ring_cap[slot], reset.data(), n, nullptr)) {
			return e->fail(OALSFX_ERR_DEVICE, e->be->error());
		}
	}
	e->groups_dirty = true;
	return OALSFX_OK;
}

int oalsfx_engine_set_sends(oalsfx_engine* e, int first_stream, int n_streams, const float* direct, const float* aux)
{
	if (!e) {
		return OALSFX_ERR_ARGUMENT;
	}
	if (first_stream < 0 || n_streams < 1 || first_stream + n_streams > e->streams || !direct || !aux) {
		return e->fail(OALSFX_ERR_ARGUMENT, "Bad stream range or null send properties.");
	}
	SendSettings d = {direct[0], direct[1], direct[2]};
	SendSettings a[kMaxSlots];
	std::memset(a, 0, sizeof(a));
	for (int i = 0; i < e->slots; ++i) {
		a[i] = SendSettings{aux[3 * i], aux[3 * i + 1], aux[3 * i + 2]};
	}
	const int id = e->send_class_for(d, a);
	for (int s = first_stream; s < first_stream + n_streams; ++s) {
		e->send_class[static_cast<size_t>(s)] = id;
	}
	e->groups_dirty = true;
	return OALSFX_OK;
}

int oalsfx_engine_mix(oalsfx_engine* e, int frames, const float* src, float* dst, int layout, int space, void* cuda_stream)
{
	if (!e) {
		return OALSFX_ERR_ARGUMENT;
	}
	if (frames == 0) {
		return OALSFX_OK; // reference: oalsfxpp.cpp:3796-3799
	}
	if (frames < 0) {
		return e->fail(OALSFX_ERR_ARGUMENT, "Negative frame count.");
	}
	if (!src) {
		return e->fail(OALSFX_ERR_ARGUMENT, "No source samples.");
	}
	if (!dst) {
		return e->fail(OALSFX_ERR_ARGUMENT, "No destination samples.");
	}
	if (layout != OALSFX_LAYOUT_STREAM_MAJOR && layout != OALSFX_LAYOUT_TILED) {
		return e->fail(OALSFX_ERR_ARGUMENT, "Unknown layout.");
	}
	const long long padded = (layout == OALSFX_LAYOUT_TILED ? static_cast<long long>(e->tiles) * kLanes : e->streams);
	const size_t count = static_cast<size_t>(padded) * static_cast<size_t>(frames) * static_cast<size_t>(e->channels);
	const float* dsrc = src;
	float* ddst = dst;
	e->note_stream(cuda_stream);
	if (space == OALSFX_SPACE_HOST && e->auto_pin_host) {
		// pageable buffers copy at a fraction of the pinned rate and do not overlap with the kernels: lock them in place
		// (a failure only means the copies stay slow)
		e->pin_host(const_cast<float*>(src), count * sizeof(float), true);
		e->pin_host(dst, count * sizeof(float), true);
	}
	if (space == OALSFX_SPACE_HOST) {
		if (count > e->stage_cap) {
			e->be->sync(cuda_stream);
			e->be->release(e->stage_in);
			e->be->release(e->stage_out);
			e->device_bytes -= static_cast<long long>(e->stage_cap) * 2 * sizeof(float);
			e->stage_in = static_cast<float*>(e->dev_alloc(count * sizeof(float)));
			e->stage_out = static_cast<float*>(e->dev_alloc(count * sizeof(float)));
			e->stage_cap = count;
			if (!e->stage_in || !e->stage_out) {
				e->stage_cap = 0;
				return e->fail(OALSFX_ERR_MEMORY, "Staging allocation failed: " + e->be->error());
			}
		}
		dsrc = e->stage_in;
		ddst = e->stage_out;
		// One block, one group covering every tile (the common case): slice the tiles and overlap the
		// upload of slice j+1, the kernel of slice j and the download of slice j-1 (PCIe is full duplex
		// and the copies, not the kernel, bound this path).
		if (e->groups_dirty && !e->rebuild_groups()) {
			return e->fail(OALSFX_ERR_DEVICE, "Group table upload failed: " + e->be->error());
		}
		int max_slices = 16;
		if (const char* tune = std::getenv("OALSFX_TUNE_SLICES")) { // experiments only
			max_slices = std::max(2, std::atoi(tune));
		}
		const int slices = std::min(e->tiles / 64, max_slices);
		if (frames <= kMaxBlockFrames && e->groups.size() == 1 && e->groups[0].identity && (!e->groups[0].multi || frames >= 2) && slices >= 2 &&
			layout == OALSFX_LAYOUT_STREAM_MAJOR) {
			if (!e->pipe_in) {
				e->pipe_in = e->be->stream_create();
				e->pipe_run = e->be->stream_create();
				e->pipe_out = e->be->stream_create();
			}
			if (e->pipe_in && e->pipe_run && e->pipe_out) {
				if (!e->be->sync(cuda_stream)) {
					return e->fail(OALSFX_ERR_DEVICE, e->be->error());
				}
				const Group g = e->groups[0];
				const size_t per_tile = static_cast<size_t>(kLanes) * static_cast<size_t>(frames) * static_cast<size_t>(e->channels);
				bool ok = true;
				for (int j = 0; j < slices && ok; ++j) {
					const int t0 = static_cast<int>(static_cast<long long>(e->tiles) * j / slices);
					const int t1 = static_cast<int>(static_cast<long long>(e->tiles) * (j + 1) / slices);
					const size_t first = per_tile * static_cast<size_t>(t0);
					const size_t last = std::min(per_tile * static_cast<size_t>(t1), count);
					ok = e->be->upload(e->stage_in + first, src + first, (last - first) * sizeof(float), e->pipe_in) &&
						e->be->stream_wait(e->pipe_run, e->pipe_in);
					e->slice_first = t0;
					e->slice_count = t1 - t0;
					ok = ok && e->launch_group(g, frames, dsrc, ddst, layout, frames, 0, true, e->pipe_run);
					e->slice_first = e->slice_count = 0;
					ok = ok && e->be->stream_wait(e->pipe_out, e->pipe_run) &&
						e->be->download(dst + first, e->stage_out + first, (last - first) * sizeof(float), e->pipe_out);
				}
				ok = ok && e->be->sync(e->pipe_out) && e->be->sync(e->pipe_run);
				if (!ok) {
					return e->fail(OALSFX_ERR_DEVICE, e->be->error());
				}
				if (g.key.pending != 0) {
					std::fill(e->pending.begin(), e->pending.end(), static_cast<uint8_t>(0));
					e->groups_dirty = true;
					e->frames_since_update = 0;
				}
				e->frames_since_update += frames;
				return OALSFX_OK;
			}
		}
		if (!e->be->upload(e->stage_in, src, count * sizeof(float), cuda_stream)) {
			return e->fail(OALSFX_ERR_DEVICE, e->be->error());
		}
	} else if (space != OALSFX_SPACE_DEVICE) {
		return e->fail(OALSFX_ERR_ARGUMENT, "Unknown pointer space.");
	}

	bool first_block = true;
	for (int done = 0; done < frames;) {
		const int todo = std::min(frames - done, kMaxBlockFrames); // oalsfxpp.cpp:3818-3826
		if (e->groups_dirty && !e->rebuild_groups()) {
			return e->fail(OALSFX_ERR_DEVICE, "Group table upload failed: " + e->be->error());
		}
		bool had_pending = false;
		const bool use_fallback = e->groups.size() == 1 && e->groups[0].multi && todo < 2;
		had_pending = use_fallback && e->groups[0].key.pending != 0;
		for (const Group& g : (use_fallback ? e->fallback_groups : e->groups)) {
			had_pending = had_pending || g.key.pending != 0;
			if (!e->launch_group(g, todo, dsrc, ddst, layout, frames, done, first_block, cuda_stream)) {
				return e->fail(OALSFX_ERR_DEVICE, e->be->error());
			}
		}
		if (had_pending) {
			std::fill(e->pending.begin(), e->pending.end(), static_cast<uint8_t>(0));
			e->groups_dirty = true;
			e->frames_since_update = 0;
		}
		e->frames_since_update += todo;
		first_block = false;
		done += todo;
	}

	if (space == OALSFX_SPACE_HOST) {
		if (!e->be->download(dst, e->stage_out, count * sizeof(float), cuda_stream) || !e->be->sync(cuda_stream)) {
			return e->fail(OALSFX_ERR_DEVICE, e->be->error());
		}
	}
	return OALSFX_OK;
}

int oalsfx_engine_pin_host(oalsfx_engine* e, void* buffer, size_t bytes)
{
	if (!e || !buffer || bytes == 0) {
		return e ? e->fail(OALSFX_ERR_ARGUMENT, "Bad pin_host arguments.") : OALSFX_ERR_ARGUMENT;
	}
	return e->pin_host(buffer, bytes, false) ? OALSFX_OK : e->fail(OALSFX_ERR_DEVICE, "Page-locking the buffer failed: " + e->be->error());
}

int oalsfx_engine_unpin_host(oalsfx_engine* e, void* buffer)
{
	if (!e || !buffer) {
		return e ? e->fail(OALSFX_ERR_ARGUMENT, "Bad unpin_host arguments.") : OALSFX_ERR_ARGUMENT;
	}
	e->quiesce();
	for (size_t i = 0; i < e->pinned_host.size(); ++i) {
		if (e->pinned_host[i].p == buffer) {
			e->be->host_unregister(buffer);
			e->pinned_host.erase(e->pinned_host.begin() + static_cast<long>(i));
			return OALSFX_OK;
		}
	}
	return e->fail(OALSFX_ERR_ARGUMENT, "The buffer is not pinned by this engine.");
}

int oalsfx_engine_mix_bus(oalsfx_engine* e, int frames, const float* src, float* dst, int layout, float* bus, void* cuda_stream)
{
	if (!e || !src || !dst || !bus || frames <= 0) {
		return e ? e->fail(OALSFX_ERR_ARGUMENT, "Bad mix_bus arguments.") : OALSFX_ERR_ARGUMENT;
	}
	// One pass over the block's output right behind the mix, on the same stream.  (Summing inside the fused kernel's
	// epilogue -- per-tile shuffle trees in its back warp, in one go, spread over the next hand-off, or in a third warp --
	// was measured at +0.21 / +0.35 / +0.57 ms per 65 536-stream block against 0.10 ms for this pass: that warp is the
	// kernel's critical path.  profiles/r02_history.md.)
	const int rc = oalsfx_engine_mix(e, frames, src, dst, layout, OALSFX_SPACE_DEVICE, cuda_stream);
	return rc != OALSFX_OK ? rc : oalsfx_engine_reduce_bus(e, frames, dst, layout, bus, cuda_stream);
}

int oalsfx_engine_reduce_bus(oalsfx_engine* e, int frames, const float* dst, int layout, float* bus, void* cuda_stream)
{
	if (!e || !dst || !bus || frames <= 0) {
		return e ? e->fail(OALSFX_ERR_ARGUMENT, "Bad reduce_bus arguments.") : OALSFX_ERR_ARGUMENT;
	}
	long long ts, ls, fs, cs;
	if (layout == OALSFX_LAYOUT_TILED) {
		ts = static_cast<long long>(frames) * e->channels * kLanes;
		ls = 1;
		fs = static_cast<long long>(e->channels) * kLanes;
		cs = kLanes;
	} else {
		ts = static_cast<long long>(frames) * e->channels * kLanes;
		ls = static_cast<long long>(frames) * e->channels;
		fs = e->channels;
		cs = 1;
	}
	e->note_stream(cuda_stream);
	++e->launches;
	if (!e->be->reduce_bus(dst, ts, ls, fs, cs, e->streams, frames, e->channels, bus, cuda_stream)) {
		return e->fail(OALSFX_ERR_DEVICE, e->be->error());
	}
	return OALSFX_OK;
}

int oalsfx_pcm_to_float(oalsfx_engine* e, const void* src, int bit_depth, float* dst, long long count, void* cuda_stream)
{
	if (!e || !src || !dst || count < 0) {
		return e ? e->fail(OALSFX_ERR_ARGUMENT, "Bad pcm_to_float arguments.") : OALSFX_ERR_ARGUMENT;
	}
	if (bit_depth != 8 && bit_depth != 16) {
		return e->fail(OALSFX_ERR_ARGUMENT, "Invalid bit depth."); // reference: oalsfxpp_test.cpp:738
	}
	e->note_stream(cuda_stream);
	++e->launches;
	if (!e->be->pcm_to_float(src, bit_depth, dst, count, cuda_stream)) {
		return e->fail(OALSFX_ERR_DEVICE, e->be->error());
	}
	return OALSFX_OK;
}

int oalsfx_debug_waveshaper(oalsfx_engine* e, const float* samples, float edge_coeff, float* out, long long count, void* cuda_stream)
{
	if (!e || !samples || !out || count < 0) {
		return e ? e->fail(OALSFX_ERR_ARGUMENT, "Bad debug_waveshaper arguments.") : OALSFX_ERR_ARGUMENT;
	}
	e->note_stream(cuda_stream);
	if (!e->be->debug_waveshaper(samples, edge_coeff, out, count, cuda_stream)) {
		return e->fail(OALSFX_ERR_DEVICE, e->be->error());
	}
	return OALSFX_OK;
}

int oalsfx_float_to_s16(oalsfx_engine* e, const float* src, int16_t* dst, int rows, long long row_len, float* row_scale,
	void* cuda_stream)
{
	if (!e || !src || !dst || rows < 0 || row_len < 0) {
		return e ? e->fail(OALSFX_ERR_ARGUMENT, "Bad float_to_s16 arguments.") : OALSFX_ERR_ARGUMENT;
	}
	e->note_stream(cuda_stream);
	++e->launches;
	if (!e->be->float_to_s16(src, dst, rows, row_len, row_scale, cuda_stream)) {
		return e->fail(OALSFX_ERR_DEVICE, e->be->error());
	}
	return OALSFX_OK;
}

// ---- state snapshot / restore --------------------------------------------------
// Everything a stream carries from one block to the next lives in three kinds of device arenas (delay-line
// rings and slot state per slot, send filter history) plus one host byte per stream (pending `update`
// bits).  A snapshot is those bytes behind a header that pins the geometry.
namespace {
struct SnapshotHeader {
	uint32_t magic, version;
	int32_t streams, tiles, channels, slots, rate, format;
	int32_t ring_cap[kMaxSlots];
	uint64_t total_bytes;
};
constexpr uint32_t kSnapshotMagic = 0x58464C4FU; // "OLFX"

void snapshot_header(const oalsfx_engine* e, SnapshotHeader& h)
{
	std::memset(&h, 0, sizeof(h));
	h.magic = kSnapshotMagic;
	h.version = 1;
	h.streams = e->streams;
	h.tiles = e->tiles;
	h.channels = e->channels;
	h.slots = e->slots;
	h.rate = e->desc.sampling_rate;
	h.format = e->desc.channel_format;
	uint64_t bytes = sizeof(SnapshotHeader) + static_cast<uint64_t>(e->streams);
	for (int s = 0; s < kMaxSlots; ++s) {
		h.ring_cap[s] = (s < e->slots ? e->ring_cap[s] : 0);
		if (s < e->slots) {
			bytes += static_cast<uint64_t>(e->tiles) * static_cast<uint64_t>(h.ring_cap[s]) * kLanes * sizeof(float);
			bytes += static_cast<uint64_t>(e->tiles) * kSlotStateWords * kLanes * sizeof(uint32_t);
		}
	}
	bytes += static_cast<uint64_t>(e->tiles) * kSendStateWords * kLanes * sizeof(uint32_t);
	h.total_bytes = bytes;
}
} // namespace

long long oalsfx_engine_snapshot_size(const oalsfx_engine* e)
{
	if (!e) {
		return 0;
	}
	SnapshotHeader h;
	snapshot_header(e, h);
	return static_cast<long long>(h.total_bytes);
}

int oalsfx_engine_snapshot(oalsfx_engine* e, void* dst, size_t bytes)
{
	if (!e || !dst) {
		return e ? e->fail(OALSFX_ERR_ARGUMENT, "Bad snapshot arguments.") : OALSFX_ERR_ARGUMENT;
	}
	SnapshotHeader h;
	snapshot_header(e, h);
	if (bytes < h.total_bytes) {
		return e->fail(OALSFX_ERR_ARGUMENT, "Snapshot buffer too small.");
	}
	if (!e->quiesce()) {
		return e->fail(OALSFX_ERR_DEVICE, e->be->error());
	}
	char* out = static_cast<char*>(dst);
	std::memcpy(out, &h, sizeof(h));
	out += sizeof(h);
	std::memcpy(out, e->pending.data(), static_cast<size_t>(e->streams));
	out += e->streams;
	bool ok = true;
	for (int s = 0; s < e->slots && ok; ++s) {
		const size_t ring_bytes = static_cast<size_t>(e->tiles) * static_cast<size_t>(e->ring_cap[s]) * kLanes * sizeof(float);
		const size_t state_bytes = static_cast<size_t>(e->tiles) * kSlotStateWords * kLanes * sizeof(uint32_t);
		if (ring_bytes) {
			ok = e->be->download(out, e->ring[s], ring_bytes, nullptr);
			out += ring_bytes;
		}
		ok = ok && e->be->download(out, e->slot_state[s], state_bytes, nullptr);
		out += state_bytes;
	}
	ok = ok && e->be->download(out, e->send_state, static_cast<size_t>(e->tiles) * kSendStateWords * kLanes * sizeof(uint32_t), nullptr);
	ok = ok && e->be->sync(nullptr);
	return ok ? OALSFX_OK : e->fail(OALSFX_ERR_DEVICE, e->be->error());
}

int oalsfx_engine_restore(oalsfx_engine* e, const void* src, size_t bytes)
{
	if (!e || !src || bytes < sizeof(SnapshotHeader)) {
		return e ? e->fail(OALSFX_ERR_ARGUMENT, "Bad restore arguments.") : OALSFX_ERR_ARGUMENT;
	}
	SnapshotHeader want, got;
	snapshot_header(e, want);
	std::memcpy(&got, src, sizeof(got));
	if (got.magic != kSnapshotMagic || got.version != want.version) {
		return e->fail(OALSFX_ERR_ARGUMENT, "Not an engine snapshot.");
	}
	if (std::memcmp(&got, &want, sizeof(got)) != 0 || bytes < got.total_bytes) {
		return e->fail(OALSFX_ERR_ARGUMENT, "Snapshot does not match this engine (streams, format, rate, slots or effect types differ).");
	}
	if (!e->quiesce()) {
		return e->fail(OALSFX_ERR_DEVICE, e->be->error());
	}
	const char* in = static_cast<const char*>(src) + sizeof(SnapshotHeader);
	std::memcpy(e->pending.data(), in, static_cast<size_t>(e->streams));
	in += e->streams;
	e->groups_dirty = true;
	e->frames_since_update = 128; // unknown history: let the steady-state kernels try (they check the stream state)
	bool ok = true;
	for (int s = 0; s < e->slots && ok; ++s) {
		const size_t ring_bytes = static_cast<size_t>(e->tiles) * static_cast<size_t>(e->ring_cap[s]) * kLanes * sizeof(float);
		const size_t state_bytes = static_cast<size_t>(e->tiles) * kSlotStateWords * kLanes * sizeof(uint32_t);
		if (ring_bytes) {
			ok = e->be->upload(e->ring[s], in, ring_bytes, nullptr);
			in += ring_bytes;
		}
		ok = ok && e->be->upload(e->slot_state[s], in, state_bytes, nullptr);
		in += state_bytes;
	}
	ok = ok && e->be->upload(e->send_state, in, static_cast<size_t>(e->tiles) * kSendStateWords * kLanes * sizeof(uint32_t), nullptr);
	ok = ok && e->be->sync(nullptr);
	return ok ? OALSFX_OK : e->fail(OALSFX_ERR_DEVICE, e->be->error());
}

int oalsfx_engine_debug_state(oalsfx_engine* e, int stream, int slot, int32_t out[4])
{
	if (!e || !out || stream < 0 || stream >= e->streams || slot < 0 || slot >= e->slots) {
		return e ? e->fail(OALSFX_ERR_ARGUMENT, "Bad debug_state arguments.") : OALSFX_ERR_ARGUMENT;
	}
	std::vector<uint32_t> words(static_cast<size_t>(kSlotStateWords) * kLanes);
	const uint32_t* p = e->slot_state[slot] + static_cast<size_t>(stream / kLanes) * kSlotStateWords * kLanes;
	if (!e->quiesce() || !e->be->download(words.data(), p, words.size() * sizeof(uint32_t), nullptr) ||
		!e->be->sync(nullptr)) {
		return e->fail(OALSFX_ERR_DEVICE, e->be->error());
	}
	const int lane = stream % kLanes;
	auto word = [&](size_t i) { return static_cast<int32_t>(words[i * kLanes + static_cast<size_t>(lane)]); };
	out[0] = out[1] = out[2] = out[3] = 0;
	const int type = e->classes[static_cast<size_t>(e->fx_class[slot][static_cast<size_t>(stream)])].type;
	switch (kind_of_type(type)) {
	case kKindModDelay: out[0] = word(offsetof(FxModDelay::State, offset) / 4); break;
	case kKindEcho: out[0] = word(offsetof(FxEcho::State, offset) / 4); break;
	case kKindRingMod: out[3] = word(offsetof(FxRingMod::State, index) / 4); break;
	case kKindReverb:
		out[0] = word(offsetof(FxReverb::State, offset) / 4);
		out[1] = word(offsetof(FxReverb::State, fade_count) / 4);
		out[2] = word(offsetof(FxReverb::State, mod_index) / 4);
		break;
	default: break;
	}
	return OALSFX_OK;
}

long long oalsfx_engine_launch_count(const oalsfx_engine* e) { return e ? e->launches : 0; }

const char* oalsfx_engine_last_kernel(const oalsfx_engine* e) { return (e && e->last_kernel >= 0) ? kernel_name(e->last_kernel) : ""; }

long long oalsfx_engine_device_bytes(const oalsfx_engine* e) { return e ? e->device_bytes : 0; }

const char* oalsfx_last_error(const oalsfx_engine* e) { return e ? e->error.c_str() : g_create_error.c_str(); }

const char* oalsfx_build_info(void) { return backend_build_info(); }

int oalsfx_effect_defaults(int effect_type, void* props, size_t props_bytes)
{
	if (!props || props_bytes != sizeof(oalsfxpp::EffectProps) || effect_type < 0 || effect_type >= kFxTypeCount) {
		return OALSFX_ERR_ARGUMENT;
	}
	oalsfxpp::Effect fx;
	std::memset(&fx, 0, sizeof(fx));
	fx.set_type_and_defaults(static_cast<oalsfxpp::EffectType>(effect_type));
	std::memcpy(props, &fx.props_, sizeof(fx.props_));
	return OALSFX_OK;
}

int oalsfx_effect_normalize(int effect_type, void* props, size_t props_bytes)
{
	if (!props || props_bytes != sizeof(oalsfxpp::EffectProps) || effect_type < 0 || effect_type >= kFxTypeCount) {
		return OALSFX_ERR_ARGUMENT;
	}
	oalsfxpp::Effect fx;
	fx.type_ = static_cast<oalsfxpp::EffectType>(effect_type);
	std::memcpy(&fx.props_, props, sizeof(fx.props_));
	fx.normalize();
	std::memcpy(props, &fx.props_, sizeof(fx.props_));
	return OALSFX_OK;
}

namespace {
struct PresetRow { const char* full; const oalsfxpp::EffectProps::Reverb* value; };
const PresetRow kPresets[] = {
#define OALSFX_PRESET_GROUP_BEGIN(G)
#define OALSFX_PRESET(G, N, ...) {#G "::" #N, &oalsfxpp::ReverbPresets::G::N},
#define OALSFX_PRESET_GROUP_END(G)
#include "oalsfxpp_presets.inc"
#undef OALSFX_PRESET_GROUP_BEGIN
#undef OALSFX_PRESET
#undef OALSFX_PRESET_GROUP_END
};
constexpr int kPresetCount = static_cast<int>(sizeof(kPresets) / sizeof(kPresets[0]));
} // namespace

int oalsfx_reverb_preset(const char* group, const char* name, void* props, size_t props_bytes)
{
	if (!group || !name || !props || props_bytes != sizeof(oalsfxpp::EffectProps)) {
		return OALSFX_ERR_ARGUMENT;
	}
	const std::string full = std::string(group) + "::" + name;
	for (int i = 0; i < kPresetCount; ++i) {
		if (full == kPresets[i].full) {
			oalsfxpp::EffectProps p;
			std::memset(&p, 0, sizeof(p));
			p.reverb_ = *kPresets[i].value;
			std::memcpy(props, &p, sizeof(p));
			return OALSFX_OK;
		}
	}
	return OALSFX_ERR_ARGUMENT;
}

long long oalsfx_plan_placement(const int32_t* class_of_stream, int n_streams, int32_t* engine_index_of_stream,
	int32_t* class_triples, int class_capacity, int* class_count)
{
	if (!class_of_stream || !engine_index_of_stream || n_streams < 1 || class_capacity < 0 || (class_capacity > 0 && !class_triples)) {
		return OALSFX_ERR_ARGUMENT;
	}
	// classes in order of first appearance, with their stream counts
	std::map<int32_t, int> slot_of_label;
	std::vector<int32_t> labels;
	std::vector<long long> count;
	for (int s = 0; s < n_streams; ++s) {
		auto it = slot_of_label.find(class_of_stream[s]);
		if (it == slot_of_label.end()) {
			it = slot_of_label.emplace(class_of_stream[s], static_cast<int>(labels.size())).first;
			labels.push_back(class_of_stream[s]);
			count.push_back(0);
		}
		++count[static_cast<size_t>(it->second)];
	}
	// every class starts on a tile boundary
	std::vector<long long> next(labels.size());
	long long at = 0;
	for (size_t c = 0; c < labels.size(); ++c) {
		next[c] = at;
		if (static_cast<int>(c) < class_capacity) {
			class_triples[3 * c + 0] = labels[c];
			class_triples[3 * c + 1] = static_cast<int32_t>(at);
			class_triples[3 * c + 2] = static_cast<int32_t>((count[c] + kLanes - 1) / kLanes * kLanes);
		}
		at += (count[c] + kLanes - 1) / kLanes * kLanes;
		if (at > 0x7FFFFFFFLL) {
			return OALSFX_ERR_ARGUMENT;
		}
	}
	for (int s = 0; s < n_streams; ++s) {
		engine_index_of_stream[s] = static_cast<int32_t>(next[static_cast<size_t>(slot_of_label[class_of_stream[s]])]++);
	}
	if (class_count) {
		*class_count = static_cast<int>(labels.size());
	}
	return at;
}

const char* oalsfx_reverb_preset_name(int index)
{
	return (index >= 0 && index < kPresetCount) ? kPresets[index].full : nullptr;
}

} // extern "C"
