// kernel_table.h -- the list of mix_kernel instantiations.
//
// X(id, CT, SF, F0, F1, F2, F3):
//   CT  compile-time channel count (0 = run-time, any layout)
//   SF  send shelf filters compiled in (false = host guarantees none is active and frames >= 2)
//   F*  effect processor per slot position
//
// "Gen*" entries process ONE effect in slot position 0 for any channel count; the engine chains
// them (dry pass first, then one accumulate pass per remaining slot) for slot signatures that have
// no fused entry.  The fused entries cover the BASELINE.json configurations in a single pass.
#ifndef OALSFX_KERNEL_TABLE_H
#define OALSFX_KERNEL_TABLE_H

#include "mix.cuh"

#define OALSFX_KERNEL_TABLE(X) \
	X(kGenDry, 0, true, FxNull, FxNull, FxNull, FxNull) \
	X(kGenModDelay, 0, true, FxModDelay, FxNull, FxNull, FxNull) \
	X(kGenCompressor, 0, true, FxCompressor, FxNull, FxNull, FxNull) \
	X(kGenDedicated, 0, true, FxDedicated, FxNull, FxNull, FxNull) \
	X(kGenDistortion, 0, true, FxDistortion, FxNull, FxNull, FxNull) \
	X(kGenEcho, 0, true, FxEcho, FxNull, FxNull, FxNull) \
	X(kGenEqualizer, 0, true, FxEqualizer, FxNull, FxNull, FxNull) \
	X(kGenRingMod, 0, true, FxRingMod, FxNull, FxNull, FxNull) \
	X(kGenReverb, 0, true, FxReverb, FxNull, FxNull, FxNull) \
	/* cfg2 / cfg4: equalizer + chorus + echo + (EAX) reverb, stereo */ \
	X(kChainStereo, 2, false, FxEqualizer, FxModDelay, FxEcho, FxReverb) \
	/* cfg3: flanger + ring modulator + distortion + compressor, mono */ \
	X(kChain2Mono, 1, false, FxModDelay, FxRingMod, FxDistortion, FxCompressor) \
	/* cfg1: one (EAX) reverb slot, mono -- and the same for stereo, the demo program's usual case */ \
	X(kReverbMono, 1, false, FxReverb, FxNull, FxNull, FxNull) \
	X(kReverbStereo, 2, false, FxReverb, FxNull, FxNull, FxNull) \
	/* cfg0: one echo slot, stereo */ \
	X(kEchoStereo, 2, false, FxEcho, FxNull, FxNull, FxNull)

// Duo kernels (duo.cuh: thread per stream, front warp = dry + slots 0..2, back warp = slot 3 +
// output, shared-memory hand-off).  DX(id, CT, F0, F1, F2, F3, twin), same twin rule as above.
#define OALSFX_DUO_TABLE(DX) \
	DX(kDuoChainStereo, 2, FxEqualizer, FxModDelay, FxEcho, FxReverb, kChainStereo)

// Class-per-tile duo kernels (duo_multi_kernel): MX(id, CT, F0, F1, F2, F3, duo id with the same signature).
#define OALSFX_MULTI_TABLE(MX) \
	MX(kMultiChainStereo, 2, FxEqualizer, FxModDelay, FxEcho, FxReverb, kDuoChainStereo)

// Quartet kernels (quartet.cuh: four pipeline stages per tile -- dry + slot 0 | slots 1, 2 + reverb input |
// reverb early half | reverb late half + output).  TX(id, CT, F0, F1, F2, F3, twin), same twin rule.
#define OALSFX_QUARTET_TABLE(TX) \
	TX(kQuartetChainStereo, 2, FxEqualizer, FxModDelay, FxEcho, FxReverb, kChainStereo) \
	/* cfg1: one (EAX) reverb slot, mono: few streams, latency-bound -- more warps per tile is the lever */ \
	TX(kQuartetReverbMono, 1, FxReverb, FxNull, FxNull, FxNull, kReverbMono) \
	TX(kQuartetReverbStereo, 2, FxReverb, FxNull, FxNull, FxNull, kReverbStereo)

// Table-mode kernels (mix.cuh, TABLE = true): the single-effect "Gen" passes with per-stream coefficient blocks
// read from HBM, for engines whose streams carry many different parameter sets.  TBX(id, Fx, kind).
#define OALSFX_TABMODE_TABLE(TBX) \
	TBX(kTabDry, FxNull, kKindNull) \
	TBX(kTabModDelay, FxModDelay, kKindModDelay) \
	TBX(kTabCompressor, FxCompressor, kKindCompressor) \
	TBX(kTabDedicated, FxDedicated, kKindDedicated) \
	TBX(kTabDistortion, FxDistortion, kKindDistortion) \
	TBX(kTabEcho, FxEcho, kKindEcho) \
	TBX(kTabEqualizer, FxEqualizer, kKindEqualizer) \
	TBX(kTabRingMod, FxRingMod, kKindRingMod) \
	TBX(kTabReverb, FxReverb, kKindReverb)

// Relay kernels (relay.cuh): any slot signature as a warp pipeline, one warp per non-null slot, the effect of
// each stage chosen at run time.  RX(id, CT, HEAVY): HEAVY = the reverb is compiled in.
#define OALSFX_RELAY_TABLE(RX) \
	RX(kRelayMono, 1, false) \
	RX(kRelayStereo, 2, false) \
	RX(kRelayMonoHeavy, 1, true) \
	RX(kRelayStereoHeavy, 2, true) \
	RX(kRelayWide, 0, false) \
	RX(kRelayWideHeavy, 0, true)
// (CT = 0: quad / 5.1 / 6.1 / 7.1, the channel count is a run-time value)
// the same with the sends' shelf filters compiled in (relay_sf_kernel)
#define OALSFX_RELAY_SF_TABLE(RX) \
	RX(kRelaySfMono, 1, false) \
	RX(kRelaySfStereo, 2, false) \
	RX(kRelaySfMonoHeavy, 1, true) \
	RX(kRelaySfStereoHeavy, 2, true) \
	RX(kRelaySfWide, 0, false) \
	RX(kRelaySfWideHeavy, 0, true)
// the same with one parameter class per tile (relay_multi_kernel)
#define OALSFX_RELAY_MULTI_TABLE(RX) \
	RX(kRelayMultiMono, 1, false) \
	RX(kRelayMultiStereo, 2, false) \
	RX(kRelayMultiMonoHeavy, 1, true) \
	RX(kRelayMultiStereoHeavy, 2, true)

// Span kernels (span.cuh): block-parallel in time.  SX(id, CT, SL, CHAIN): SL = streams of a tile per CTA (a tile is
// shared by 32 / SL CTAs), CHAIN = the 4-slot equalizer + chorus + echo + reverb chain, else the single reverb slot;
// the three ids of one signature are consecutive (SL = 32, 16, 8).
#define OALSFX_SPAN_TABLE(SX) \
	SX(kSpanReverbMono, 1, 32, false) \
	SX(kSpanReverbMono16, 1, 16, false) \
	SX(kSpanReverbMono8, 1, 8, false) \
	SX(kSpanReverbStereo, 2, 32, false) \
	SX(kSpanReverbStereo16, 2, 16, false) \
	SX(kSpanReverbStereo8, 2, 8, false) \
	SX(kSpanChainStereo, 2, 32, true) \
	SX(kSpanChainStereo16, 2, 16, true) \
	SX(kSpanChainStereo8, 2, 8, true)

// The same with bulk-asynchronous ring traffic, whole tiles only (span_bulk_kernel).  BX(id, CT, CHAIN).
#define OALSFX_SPAN_BULK_TABLE(BX) \
	BX(kSpanBulkReverbMono, 1, false) \
	BX(kSpanBulkReverbStereo, 2, false) \
	BX(kSpanBulkChainStereo, 2, true)

// Scan kernels (scan.cuh): a lone equalizer slot as a chunked linear-recurrence scan over time, opt-in (re-associates).
// SCX(id, CT).
#define OALSFX_SCAN_TABLE(SCX) \
	SCX(kScanEqualizerMono, 1) \
	SCX(kScanEqualizerStereo, 2)

namespace oalsfx {

enum KernelId : int {
#define OALSFX_X(id, CT, SF, F0, F1, F2, F3) id,
	OALSFX_KERNEL_TABLE(OALSFX_X)
#undef OALSFX_X
	kKernelCount,
	kFusedFirst = kKernelCount - 1,
#define OALSFX_DX(id, CT, F0, F1, F2, F3, twin) id,
	OALSFX_DUO_TABLE(OALSFX_DX)
#undef OALSFX_DX
#define OALSFX_TX(id, CT, F0, F1, F2, F3, twin) id,
	OALSFX_QUARTET_TABLE(OALSFX_TX)
#undef OALSFX_TX
#define OALSFX_TBX(id, Fx, kind) id,
	OALSFX_TABMODE_TABLE(OALSFX_TBX)
#undef OALSFX_TBX
#define OALSFX_RX(id, CT, HEAVY) id,
	OALSFX_RELAY_TABLE(OALSFX_RX)
	OALSFX_RELAY_MULTI_TABLE(OALSFX_RX)
	OALSFX_RELAY_SF_TABLE(OALSFX_RX)
#undef OALSFX_RX
#define OALSFX_SX(id, CT, SL, CHAIN) id,
	OALSFX_SPAN_TABLE(OALSFX_SX)
#undef OALSFX_SX
#define OALSFX_BX(id, CT, CHAIN) id,
	OALSFX_SPAN_BULK_TABLE(OALSFX_BX)
#undef OALSFX_BX
#define OALSFX_MX(id, CT, F0, F1, F2, F3, duo) id,
	OALSFX_MULTI_TABLE(OALSFX_MX)
#undef OALSFX_MX
#define OALSFX_SCX(id, CT) id,
	OALSFX_SCAN_TABLE(OALSFX_SCX)
#undef OALSFX_SCX
	kKernelEnd
};

// quartet kernel id for a thread-per-stream twin id, or -1
inline int quartet_for_twin(int twin_id)
{
#define OALSFX_TX(id, CT, F0, F1, F2, F3, twin) if (twin_id == twin) return id;
	OALSFX_QUARTET_TABLE(OALSFX_TX)
#undef OALSFX_TX
	return -1;
}

// duo kernel id for a thread-per-stream twin id, or -1
inline int duo_for_twin(int twin_id)
{
#define OALSFX_DX(id, CT, F0, F1, F2, F3, twin) if (twin_id == twin) return id;
	OALSFX_DUO_TABLE(OALSFX_DX)
#undef OALSFX_DX
	return -1;
}

// thread-per-stream twin of a duo / quartet kernel id, or -1
inline int twin_of_fused(int fused_id)
{
#define OALSFX_DX(id, CT, F0, F1, F2, F3, twin) if (fused_id == id) return twin;
	OALSFX_DUO_TABLE(OALSFX_DX)
#undef OALSFX_DX
#define OALSFX_TX(id, CT, F0, F1, F2, F3, twin) if (fused_id == id) return twin;
	OALSFX_QUARTET_TABLE(OALSFX_TX)
#undef OALSFX_TX
	return -1;
}

inline const char* kernel_name(int id);

// Effect "kind" = which processor handles an FxType (chorus/flanger and reverb/EAX share one).
enum FxKind : int { kKindNull, kKindModDelay, kKindCompressor, kKindDedicated, kKindDistortion, kKindEcho,
	kKindEqualizer, kKindRingMod, kKindReverb };

// table-mode kernel id for an effect kind
inline int tab_kernel_for_kind(int kind)
{
#define OALSFX_TBX(id, Fx, k) if (kind == k) return id;
	OALSFX_TABMODE_TABLE(OALSFX_TBX)
#undef OALSFX_TBX
	return -1;
}

template <class F> struct KindOf;
template <> struct KindOf<FxNull> { static constexpr int value = kKindNull; };
template <> struct KindOf<FxModDelay> { static constexpr int value = kKindModDelay; };
template <> struct KindOf<FxCompressor> { static constexpr int value = kKindCompressor; };
template <> struct KindOf<FxDedicated> { static constexpr int value = kKindDedicated; };
template <> struct KindOf<FxDistortion> { static constexpr int value = kKindDistortion; };
template <> struct KindOf<FxEcho> { static constexpr int value = kKindEcho; };
template <> struct KindOf<FxEqualizer> { static constexpr int value = kKindEqualizer; };
template <> struct KindOf<FxRingMod> { static constexpr int value = kKindRingMod; };
template <> struct KindOf<FxReverb> { static constexpr int value = kKindReverb; };

inline int kind_of_type(int fx_type)
{
	switch (fx_type) {
	case kFxChorus: case kFxFlanger: return kKindModDelay;
	case kFxCompressor: return kKindCompressor;
	case kFxDedicatedDialog: case kFxDedicatedLfe: return kKindDedicated;
	case kFxDistortion: return kKindDistortion;
	case kFxEcho: return kKindEcho;
	case kFxEqualizer: return kKindEqualizer;
	case kFxRingModulator: return kKindRingMod;
	case kFxReverb: case kFxEaxReverb: return kKindReverb;
	default: return kKindNull;
	}
}

struct KernelInfo { int id; int ct; bool sf; int kind[4]; const char* name; };

inline const KernelInfo* kernel_infos()
{
	static const KernelInfo infos[] = {
#define OALSFX_X(id, CT, SF, F0, F1, F2, F3) \
		{id, CT, SF, {KindOf<F0>::value, KindOf<F1>::value, KindOf<F2>::value, KindOf<F3>::value}, #id},
		OALSFX_KERNEL_TABLE(OALSFX_X)
#undef OALSFX_X
	};
	return infos;
}

inline const char* kernel_name(int id)
{
	if (id >= 0 && id < kKernelCount) {
		return kernel_infos()[id].name;
	}
#define OALSFX_DX(did, CT, F0, F1, F2, F3, twin) if (id == did) return #did;
	OALSFX_DUO_TABLE(OALSFX_DX)
#undef OALSFX_DX
#define OALSFX_TX(tid, CT, F0, F1, F2, F3, twin) if (id == tid) return #tid;
	OALSFX_QUARTET_TABLE(OALSFX_TX)
#undef OALSFX_TX
#define OALSFX_TBX(tid, Fx, kind) if (id == tid) return #tid;
	OALSFX_TABMODE_TABLE(OALSFX_TBX)
#undef OALSFX_TBX
#define OALSFX_RX(rid, CT, HEAVY) if (id == rid) return #rid;
	OALSFX_RELAY_TABLE(OALSFX_RX)
	OALSFX_RELAY_MULTI_TABLE(OALSFX_RX)
	OALSFX_RELAY_SF_TABLE(OALSFX_RX)
#undef OALSFX_RX
#define OALSFX_SX(sid, CT, SL, CHAIN) if (id == sid) return #sid;
	OALSFX_SPAN_TABLE(OALSFX_SX)
#undef OALSFX_SX
#define OALSFX_BX(bid, CT, CHAIN) if (id == bid) return #bid;
	OALSFX_SPAN_BULK_TABLE(OALSFX_BX)
#undef OALSFX_BX
#define OALSFX_MX(mid, CT, F0, F1, F2, F3, duo) if (id == mid) return #mid;
	OALSFX_MULTI_TABLE(OALSFX_MX)
#undef OALSFX_MX
#define OALSFX_SCX(cid, CT) if (id == cid) return #cid;
	OALSFX_SCAN_TABLE(OALSFX_SCX)
#undef OALSFX_SCX
	return "?";
}

} // namespace oalsfx

#endif
