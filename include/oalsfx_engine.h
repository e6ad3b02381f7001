/* oalsfx_engine.h -- C ABI of the batched B200 effects engine (the drop-in boundary).
 *
 * The reference has no FFI of its own; its only boundary is the C++ class oalsfxpp::Api
 * (reference: src/oalsfxpp.h:760-922).  This ABI is what a binding for that class's hot path
 * would call: one engine owns `num_streams` independent streams, each of which behaves exactly
 * like one reference `Api` instance (same channel format, sampling rate and slot count for the
 * whole engine).  include/oalsfxpp.h re-exports the reference's C++ class on top of it with
 * num_streams = 1.
 *
 * Plain pointers and sizes only; no exceptions cross this boundary.  Every call returns
 * OALSFX_OK (0) or a negative error code; oalsfx_last_error() gives the message.
 * There is no CPU implementation behind this ABI: creation fails without a CUDA device.
 */
#ifndef OALSFX_ENGINE_H
#define OALSFX_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oalsfx_engine oalsfx_engine;

enum {
	OALSFX_OK = 0,
	OALSFX_ERR_ARGUMENT = -1,   /* bad index / pointer / size */
	OALSFX_ERR_FORMAT = -2,     /* invalid channel format        (reference: oalsfxpp.cpp:2855-2861) */
	OALSFX_ERR_RATE = -3,       /* sampling rate out of range    (reference: oalsfxpp.cpp:2863-2868) */
	OALSFX_ERR_EFFECTS = -4,    /* effect count out of range     (reference: oalsfxpp.cpp:2870-2875) */
	OALSFX_ERR_DEVICE = -5,     /* CUDA failure / no device */
	OALSFX_ERR_MEMORY = -6
};

/* Sample buffer layouts for oalsfx_engine_mix.  F = frames of the call, C = channels. */
enum {
	OALSFX_LAYOUT_STREAM_MAJOR = 0, /* [stream][frame][channel]: every stream's buffer is exactly the
	                                   interleaved buffer Api::mix takes (reference: oalsfxpp.h:872-875) */
	OALSFX_LAYOUT_TILED = 1         /* [tile = stream/32][frame][channel][lane = stream%32]: engine-native,
	                                   fully coalesced; the stream count is padded to a multiple of 32 */
};

enum { OALSFX_SPACE_HOST = 0, OALSFX_SPACE_DEVICE = 1 };

typedef struct oalsfx_engine_desc {
	int32_t device;         /* CUDA device ordinal */
	int32_t num_streams;    /* >= 1 */
	int32_t channel_format; /* oalsfxpp::ChannelFormat value (1 mono .. 7 seven_point_one) */
	int32_t sampling_rate;  /* >= 8000 */
	int32_t effect_count;   /* 1..4 slots per stream */
} oalsfx_engine_desc;

/* Replaces Api::initialize (reference: oalsfxpp.cpp:3480-3504, 2846-2905). */
int oalsfx_engine_create(const oalsfx_engine_desc* desc, oalsfx_engine** out);

/* Replaces Api::uninitialize / ~Api (reference: oalsfxpp.cpp:3831-3834). */
void oalsfx_engine_destroy(oalsfx_engine* e);

/* Replaces EffectSlot::set_effect as reached from Api::apply_changes (reference:
 * oalsfxpp.cpp:2688-2709, 3748-3756) for streams [first_stream, first_stream + n_streams).
 * `props` points at the bytes of an oalsfxpp::EffectProps (108 bytes; ignored for the null
 * effect); they are normalized (clamped) like Effect::normalize does.  A change of effect TYPE
 * resets that slot's state (delay lines, filter history), a change of properties keeps it.  The
 * new coefficients take effect at the top of the next mixed block (reference: oalsfxpp.cpp:3000). */
int oalsfx_engine_set_effect(oalsfx_engine* e, int first_stream, int n_streams, int slot,
	int effect_type, const void* props, size_t props_bytes);

/* Replaces the send half of Api::apply_changes (reference: oalsfxpp.cpp:3760-3780, 3348-3395).
 * direct = {gain, gain_hf, gain_lf}; aux = effect_count such triples (slot order). */
int oalsfx_engine_set_sends(oalsfx_engine* e, int first_stream, int n_streams,
	const float* direct, const float* aux);

/* Ordering: with device buffers oalsfx_engine_mix / mix_bus / reduce_bus / the PCM conversions are ENQUEUED on
 * `cuda_stream` and return at once; consecutive calls on the same stream run in order.  Every other call that touches
 * engine state on the device -- set_effect / set_sends (state reset, arena growth, table uploads at the next mix),
 * snapshot, restore, debug_state, destroy -- first waits for the stream the caller last passed in, so it is safe to
 * call them at any time, also with cudaStreamNonBlocking streams.  Using several streams on one engine concurrently
 * is the caller's to order.  Precondition of the fused kernels: finite input samples (an Inf / NaN sample times a
 * gain the reference skips as inaudible would reach output channels the reference leaves untouched). */
/* Replaces Api::mix for all streams at once (reference: oalsfxpp.cpp:3785-3829, 2984-3037).
 * `frames` per stream; internally cut into blocks of <= 2048 frames exactly like the reference.
 * src/dst hold num_streams*frames*C floats (TILED: stream count rounded up to 32) in `space`;
 * dst is overwritten, not clipped.  Device buffers must not overlap.  With host buffers the call
 * stages through device memory and returns after the results are back; with device buffers the
 * work is enqueued on `cuda_stream` (a cudaStream_t, may be NULL) and the call returns at once. */
int oalsfx_engine_mix(oalsfx_engine* e, int frames, const float* src, float* dst,
	int layout, int space, void* cuda_stream);

/* Optional all-streams output bus: bus[frame][channel] (device, frames*C floats) = sum over this
 * engine's streams of a STREAM_MAJOR/TILED device buffer `dst` produced by oalsfx_engine_mix.
 * No reference counterpart; the cross-GPU sum is one all-reduce of `bus`. */
int oalsfx_engine_reduce_bus(oalsfx_engine* e, int frames, const float* dst, int layout,
	float* bus, void* cuda_stream);

/* Host buffers: oalsfx_engine_mix with OALSFX_SPACE_HOST copies at the pinned rate (and overlaps copies with kernels)
 * only from page-locked memory; pageable memory (malloc, new, std::vector) goes through the driver's bounce buffers at a
 * fraction of it.  oalsfx_engine_pin_host page-locks a caller's buffer IN PLACE (cudaHostRegister) until
 * oalsfx_engine_unpin_host or oalsfx_engine_destroy; the buffer must stay allocated for that long.  With the environment
 * variable OALSFX_PIN_HOST=1 the engine does this by itself for every host buffer it is handed (same lifetime rule:
 * do not free or unmap such a buffer while the engine lives; moving or resizing is detected by address overlap). */
int oalsfx_engine_pin_host(oalsfx_engine* e, void* buffer, size_t bytes);
int oalsfx_engine_unpin_host(oalsfx_engine* e, void* buffer);

/* oalsfx_engine_mix (device buffers) followed by oalsfx_engine_reduce_bus on the same stream: the block's output and
 * bus[frame * C + c] = its sum over the engine's streams (device array of frames * C floats) from one call. */
int oalsfx_engine_mix_bus(oalsfx_engine* e, int frames, const float* src, float* dst, int layout,
	float* bus, void* cuda_stream);

/* PCM formats either side of the path (device buffers).  The reference's only
 * producer and consumer of Api::mix buffers is its WAV demo, and these are its conversions:
 *   oalsfx_pcm_to_float : bit_depth 8: (u8 - 128) / 128.0f, 16: s16 / 32768.0f  (oalsfxpp_test.cpp:703-740)
 *   oalsfx_float_to_s16 : each of `rows` buffers of `row_len` samples (one per stream) is peak-normalised the
 *                         way WavFile::save does it -- scale = 1 / max(max(1, max x), -min(-1, min x)),
 *                         s16 = (int16)(scale * x * 32767.0f), truncating (oalsfxpp_test.cpp:602-651);
 *                         row_scale (device, `rows` floats, may be NULL) receives the scales.
 * They let a caller keep 2-byte samples on the host side of PCIe and in HBM between calls. */
int oalsfx_pcm_to_float(oalsfx_engine* e, const void* src, int bit_depth, float* dst, long long count,
	void* cuda_stream);
int oalsfx_float_to_s16(oalsfx_engine* e, const float* src, int16_t* dst, int rows, long long row_len,
	float* row_scale, void* cuda_stream);

/* State snapshot / restore (the reference has no counterpart: an Api's delay lines
 * and filter histories are private).  A snapshot holds every byte the streams carry from one block to the
 * next (delay-line rings, slot state, send filter histories, pending-update flags) into a HOST buffer of
 * oalsfx_engine_snapshot_size() bytes; restore puts it back into an engine of identical geometry whose
 * slots hold the same effect types and properties (the caller's configuration is not part of it).  After
 * a restore the engine continues bit for bit as the snapshotted one would have.  Both synchronize. */
long long oalsfx_engine_snapshot_size(const oalsfx_engine* e);
int oalsfx_engine_snapshot(oalsfx_engine* e, void* dst, size_t bytes);
int oalsfx_engine_restore(oalsfx_engine* e, const void* src, size_t bytes);

/* Integer state of one stream's slot, for bit-exact checks: out[0]=ring write offset,
 * out[1]=reverb fade_count, out[2]=reverb mod index, out[3]=ring-mod phase index.
 * Unused entries are 0.  Synchronizes the device. */
int oalsfx_engine_debug_state(oalsfx_engine* e, int stream, int slot, int32_t out[4]);

/* Test hook: the three waveshapers of the distortion effect (oalsfxpp.cpp:4720-4722) applied to `count` samples in a
 * DEVICE buffer with the given edge coefficient, through the very function the distortion stage runs (fx.cuh,
 * FxDistortion::shape: the divisions as one batched correctly-rounded sequence instead of the `/` operator), so that a
 * test can hold it bit for bit against IEEE division over the whole operand range.  Asynchronous on cuda_stream. */
int oalsfx_debug_waveshaper(oalsfx_engine* e, const float* samples, float edge_coeff, float* out, long long count, void* cuda_stream);

/* How many of the engine's kernels have been launched so far (bench.py's gpu_launches). */
long long oalsfx_engine_launch_count(const oalsfx_engine* e);

/* Name of the mix kernel the engine launched most recently ("" before the first mix), e.g. "kDuoChainStereo":
 * which of the kernel families served the last block (measurement records, tests of the selection logic). */
const char* oalsfx_engine_last_kernel(const oalsfx_engine* e);

/* Bytes of device memory the engine currently holds. */
long long oalsfx_engine_device_bytes(const oalsfx_engine* e);

/* Message for the last failed call on this engine (or on creation when e == NULL). */
const char* oalsfx_last_error(const oalsfx_engine* e);

/* Property helpers for bindings (reference: Effect::set_type_and_defaults / Effect::normalize,
 * oalsfxpp.cpp:1782-1833; ReverbPresets, oalsfxpp.cpp:1938-2191).  `props` = 108-byte EffectProps. */
int oalsfx_effect_defaults(int effect_type, void* props, size_t props_bytes);
int oalsfx_effect_normalize(int effect_type, void* props, size_t props_bytes);
/* group/name as in the header, e.g. ("Default", "forest"); index-based enumeration via
 * oalsfx_reverb_preset_name(i) which returns "Group::name" or NULL past the end. */
int oalsfx_reverb_preset(const char* group, const char* name, void* props, size_t props_bytes);
const char* oalsfx_reverb_preset_name(int index);

/* Stream placement for many parameter sets.  The engine runs a 32-stream tile at the fused kernels' speed
 * when all of its streams share one parameter set ("class"); a tile of mixed classes takes table mode, ~15x slower per
 * stream.  A host that owns many voices with arbitrary presets decides the ORDER of its streams in the engine -- each
 * stream is an independent row of the mix buffers, as each Api instance of the reference has its own buffer
 * (oalsfxpp.cpp:2820-2827) -- and this helper computes an order that keeps every tile class-pure: given one label per
 * caller stream (equal labels = the same settings in every slot and send), classes are laid out one after the other in
 * order of first appearance, each padded to a multiple of 32; streams of a class keep their relative order.
 * engine_index_of_stream[s] receives the engine stream index of caller stream s.  Returns the number of engine streams to
 * create (>= n_streams; indices nobody was assigned are silent streams: feed zeros, ignore their output), or a negative
 * OALSFX_ERR_* code.  class_triples / class_count are optional: with class_capacity > 0, up to class_capacity triples
 * (label, first engine index, length of the class's range INCLUDING its padding streams: a multiple of 32) are written,
 * so that one oalsfx_engine_set_effect call per class configures its whole range -- the padding streams must carry their
 * tile's settings, or the tile is a mixed one again. */
long long oalsfx_plan_placement(const int32_t* class_of_stream, int n_streams, int32_t* engine_index_of_stream,
	int32_t* class_triples, int class_capacity, int* class_count);

/* Library build identification, e.g. "oalsfx_b200 sm_100a cuda". */
const char* oalsfx_build_info(void);

#ifdef __cplusplus
}
#endif

#endif /* OALSFX_ENGINE_H */
