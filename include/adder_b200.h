/*
 * adder_b200.h — C ABI of the B200-native framed→ADΔER per-pixel transcode path.
 *
 * This is the drop-in boundary for ONE hot path of ac-freeman/adder-codec-rs:
 *   Framed::consume            adder-codec-rs/src/transcoder/source/framed.rs:127-157
 *   Video::integrate_matrix    adder-codec-rs/src/transcoder/source/video.rs:651-778
 *   integrate_for_px           adder-codec-rs/src/transcoder/source/video.rs:1317-1380
 *   PixelArena::*              adder-codec-rs/src/transcoder/event_pixel_tree.rs:68-532
 *   u8::get_frame_value        adder-codec-rs/src/framer/scale_intensity.rs:58-104
 *
 * The reference has no FFI of its own (it is all Rust); the entry points below
 * are what a Rust `extern "C"` block inside `Video<W>` would bind when
 * `event_pixel_trees: Array3<PixelArena>` (video.rs:325) is replaced by an opaque device handle.
 * INTEGRATION.md shows that binding.  Every function cites the reference item it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; nothing unwinds across the boundary; every call returns an
 *    `adder_status` (0 = ok).  `adder_b200_last_error()` gives a thread-local message.
 *  - a handle is NOT thread-safe (mirrors `&mut self` on Source::consume, video.rs:1421).
 *  - the library owns all device state and one CUDA stream per handle.  There is no CPU fallback:
 *    without a usable CUDA device `adder_b200_video_create` fails with ADDER_ERR_NO_DEVICE.
 */
#ifndef ADDER_B200_H
#define ADDER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADDER_B200_ABI_VERSION 1

/* ---- constants: adder-codec-core/src/lib.rs:181-193 ------------------------------------------ */
#define ADDER_D_MAX 127u              /* D_MAX */
#define ADDER_D_ZERO_INTEGRATION 128u /* D_ZERO_INTEGRATION */
#define ADDER_D_NO_EVENT 253u         /* D_NO_EVENT */
#define ADDER_D_EMPTY 255u            /* D_EMPTY */
#define ADDER_C_NONE 0xFFu            /* Coord.c == None (single-channel plane), lib.rs:263-274 */

/* ---- Event: adder-codec-core/src/lib.rs:371-377 ----------------------------------------------
 * The Rust `Event` is `repr(packed)` with an `Option<u8>` inside, which is not a stable ABI, so
 * the boundary uses this defined 12-byte little-endian record and the Rust shim maps it to
 * `Event { coord: Coord { x, y, c }, d, t }` (c == ADDER_C_NONE  <=>  None).                    */
typedef struct adder_event {
  uint16_t x;
  uint16_t y;
  uint8_t c; /* channel, or ADDER_C_NONE when the plane has one channel */
  uint8_t d;
  uint16_t reserved; /* always 0 */
  uint32_t t;
} adder_event_t;

typedef enum adder_status {
  ADDER_OK = 0,
  ADDER_ERR_BAD_PARAMS = 1,   /* SourceError::BadParams, video.rs:55-122 */
  ADDER_ERR_NO_DEVICE = 2,    /* no CUDA device / driver: the product path has no CPU fallback */
  ADDER_ERR_CUDA = 3,         /* a CUDA runtime call failed; see adder_b200_last_error() */
  ADDER_ERR_CAPACITY = 4,     /* caller's event buffer too small; *n_events holds the need */
  ADDER_ERR_ARENA_DEPTH = 5,  /* a pixel's node stack outgrew the allocated depth
                                 (reference: panic "Infinite loop detected", event_pixel_tree.rs:387) */
  ADDER_ERR_UNSUPPORTED = 6,  /* e.g. Mode::Continuous on the GPU path (only FramePerfect is in scope) */
  ADDER_ERR_INTERNAL = 7,     /* a state invariant the kernel relies on was violated */
  ADDER_ERR_NOMEM = 8
} adder_status;

/* Mode: adder-codec-core/src/lib.rs:196-205 */
typedef enum adder_pixel_tree_mode { ADDER_MODE_FRAME_PERFECT = 0, ADDER_MODE_CONTINUOUS = 1 } adder_pixel_tree_mode;
/* PixelMultiMode: lib.rs:207-213 (default Collapse) */
typedef enum adder_pixel_multi_mode { ADDER_MULTI_NORMAL = 0, ADDER_MULTI_COLLAPSE = 1 } adder_pixel_multi_mode;
/* TimeMode: lib.rs:72-83 (default AbsoluteT); numeric values = the header's enum index */
typedef enum adder_time_mode { ADDER_TIME_DELTA_T = 0, ADDER_TIME_ABSOLUTE_T = 1, ADDER_TIME_MIXED = 2 } adder_time_mode;
/* FramedViewMode: video.rs:140-158 */
typedef enum adder_view_mode { ADDER_VIEW_INTENSITY = 0, ADDER_VIEW_D = 1, ADDER_VIEW_DELTA_T = 2, ADDER_VIEW_SAE = 3 } adder_view_mode;

/* CrfParameters: adder-codec-core/src/codec/rate_controller.rs:40-53 */
typedef struct adder_crf_parameters {
  uint8_t c_thresh_baseline;
  uint8_t c_thresh_max;
  uint8_t c_increase_velocity;
  uint8_t reserved;
  uint16_t feature_c_radius;
  uint16_t reserved2;
} adder_crf_parameters_t;

/* Opaque: stands for Video<W>.state + Video<W>.event_pixel_trees (video.rs:322-345). */
typedef struct adder_b200_video adder_b200_video;

/* ---- library ---------------------------------------------------------------------------------- */
int adder_b200_abi_version(void);
const char* adder_b200_last_error(void);
/* Number of CUDA devices visible, or a negative adder_status. */
int adder_b200_device_count(void);
/* CRF table lookup: Crf::new, rate_controller.rs:55-70 (table :5-18). crf in 0..=9. */
int adder_b200_crf_parameters(uint8_t crf, uint16_t plane_w, uint16_t plane_h, adder_crf_parameters_t* out);

/* ---- construction: Video::new, video.rs:350-438 ----------------------------------------------
 * All W*H*C pixels start as PixelArena::new(1.0, coord) (event_pixel_tree.rs:69-87): one fresh
 * node, c_thresh 10, c_increase_counter 1, TimeMode::AbsoluteT.  Defaults as VideoState::default
 * (video.rs:226-243): chunk_rows 1, in_interval_count 1, ref_time 255, delta_t_max 7650,
 * multi-mode Collapse, CRF parameters of quality 3, view mode Intensity.
 * `max_depth` = node-stack depth to allocate per pixel (0 = derive from delta_t_max/ref_time when
 * first needed; at most 31, the reference's own iteration guard).                              */
int adder_b200_video_create(uint16_t width, uint16_t height, uint8_t channels, int pixel_tree_mode,
                            int device, uint32_t max_depth, adder_b200_video** out);
void adder_b200_video_destroy(adder_b200_video* v);

/* ---- builder / setters ------------------------------------------------------ */
/* Video::chunk_rows, video.rs:473-481 */
int adder_b200_video_chunk_rows(adder_b200_video* v, uint32_t chunk_rows);
/* Video::time_parameters, video.rs:493-537.  time_mode < 0 means None (keep).  Like the reference,
 * out-of-range values keep the old ones and still return ADDER_OK; *applied (may be NULL) says
 * which happened.  The delta_t_max % ref_time check is the CALLER's (framed.rs:100-109, :223-229). */
int adder_b200_video_time_parameters(adder_b200_video* v, uint32_t tps, uint32_t ref_time,
                                     uint32_t delta_t_max, int time_mode, int* applied);
/* Side effects of Video::write_out on the transcode state, video.rs:546-636:
 * pixel_multi_mode (<0 = None -> Collapse) and per-pixel time_mode (<0 = None -> keep). */
int adder_b200_video_write_out(adder_b200_video* v, int time_mode, int pixel_multi_mode);
/* Video::update_crf, video.rs:1241-1251: parameters from the table, every px c_thresh=baseline, counter=0 */
int adder_b200_video_update_crf(adder_b200_video* v, uint8_t crf);
/* Video::update_quality_manual, video.rs:1264-1287 */
int adder_b200_video_update_quality_manual(adder_b200_video* v, uint8_t c_thresh_baseline, uint8_t c_thresh_max,
                                           uint32_t delta_t_max_multiplier, uint8_t c_increase_velocity,
                                           float feature_c_radius);
/* Video::update_encoder_options (video.rs:1289-1291) and the `encoder_options` argument of write_out
 * (video.rs:553, :634): replaces the CRF parameter set (c_thresh_max, c_increase_velocity, ...) that
 * integrate_matrix reads (video.rs:660) WITHOUT touching any pixel's c_thresh or counter. */
int adder_b200_video_set_crf_parameters(adder_b200_video* v, const adder_crf_parameters_t* params);
/* Video::update_delta_t_max, video.rs:819-822 */
int adder_b200_video_update_delta_t_max(adder_b200_video* v, uint32_t delta_t_max);
/* Video::c_thresh_pos / update_adder_thresh_pos (deprecated), video.rs:445-455, :842-851 */
int adder_b200_video_c_thresh_pos(adder_b200_video* v, uint8_t c);
/* The per-pixel write done by handle_roi (video.rs:865-881) and the feature radius reset
 * (video.rs:1089-1104): c_thresh := value for x0..=x1, y0..=y1, all channels. */
int adder_b200_video_set_c_thresh_rect(adder_b200_video* v, uint16_t x0, uint16_t y0, uint16_t x1, uint16_t y1,
                                       uint8_t value);
/* Video.instantaneous_view_mode, video.rs:331 */
int adder_b200_video_set_view_mode(adder_b200_video* v, int view_mode);
/* VideoState.in_interval_count, video.rs:203 (adder-viz zeroes it on restart, adder.rs:155-166) */
int adder_b200_video_set_in_interval_count(adder_b200_video* v, uint32_t n);

/* Row-band sharding: this handle holds rows [row0, row0+height) of a taller frame
 * owned by several GPUs; row0 is added to the y of every event it emits, so that concatenating the
 * bands' streams in band order IS the reference's raster order (video.rs:677-734).  Default 0. */
int adder_b200_video_set_row_offset(adder_b200_video* v, uint16_t row0);
/* Measurement aid: while on, frames run through an instrumented twin of the kernel that also counts
 * [0] node loads, [1] node stores, [2] display-byte writes, [3] events (the data-dependent terms of the
 * algorithmic bytes) and the live nodes of every pixel at frame [4] entry and [5] exit (the L of
 * SURVEY.md §8(d)'s formula).  Turning it on or off zeroes the counters.  Never on in timed runs. */
int adder_b200_video_set_counting(adder_b200_video* v, int on);
int adder_b200_video_read_counters(adder_b200_video* v, uint64_t out[6]);

/* ---- getters ---------------------------------------------------------------------------------- */
typedef struct adder_b200_video_info {
  uint16_t width, height;
  uint8_t channels;
  uint8_t pixel_tree_mode, pixel_multi_mode, time_mode, view_mode;
  uint8_t state_form;        /* how the node stacks are held at the moment: 0 = one record per two levels, every level with its own
                              * integration / delta_t (the reference's PixelNode); 1 = offset form (levels below the root hold
                              * offsets against the root and are touched only when they fire; chosen by the library when
                              * PixelMultiMode::Collapse, an integral time_spanned and delta_t_max < 2^23 make it exact).
                              * Events, display bytes and adder_b200_video_read_px are the same in both. */
  uint8_t reserved[2];
  uint32_t chunk_rows, n_chunks;
  uint32_t in_interval_count, tps, ref_time, delta_t_max;
  adder_crf_parameters_t crf;
  uint32_t max_depth;        /* allocated node-stack depth */
  uint32_t device;
  uint64_t state_bytes;      /* device bytes held for per-pixel state */
  uint64_t events_capacity;  /* device event-buffer capacity, records */
} adder_b200_video_info_t;
int adder_b200_video_get_info(const adder_b200_video* v, adder_b200_video_info_t* out);

/* ---- the hot path: Video::integrate_matrix, video.rs:651-778 ---------------------------------
 * Host-buffer form (what Framed::consume calls, framed.rs:131-134).
 *   frame            H rows of W*C bytes, `row_pitch` bytes apart (row_pitch 0 = W*C)   [host]
 *   time_spanned     ticks the frame spans (Framed passes ref_time as f32)
 *   events_out       capacity `events_cap` records; receives this frame's events in the
 *                    reference's order: chunk by chunk, raster (y,x,c) inside a chunk, each
 *                    pixel's events contiguous in push order                              [host]
 *   chunk_counts     n_chunks = ceil(H/chunk_rows) entries: events per chunk, i.e. the lengths of
 *                    the reference's Vec<Vec<Event>> (driver.rs:566 needs exactly n_chunks) [host, may be NULL]
 *   n_events         total events of the frame (set even on ADDER_ERR_CAPACITY)
 * On ADDER_ERR_CAPACITY nothing is lost: the events stay on the device and
 * adder_b200_video_fetch_events() re-reads them into a larger buffer.
 * Effects kept from the reference: in_interval_count += 1 (video.rs:662); set_initial_d when
 * in_interval_count == 0 (video.rs:656-658, :780-801); running_intensities updated (:713-730). */
int adder_b200_video_integrate_matrix(adder_b200_video* v, const uint8_t* frame, size_t row_pitch,
                                      float time_spanned, adder_event_t* events_out, size_t events_cap,
                                      uint32_t* chunk_counts, uint64_t* n_events);
/* Re-read the last frame's events (after ADDER_ERR_CAPACITY). */
int adder_b200_video_fetch_events(adder_b200_video* v, adder_event_t* events_out, size_t events_cap,
                                  uint32_t* chunk_counts, uint64_t* n_events);
/* VideoState.running_intensities / Video.display_frame_features (video.rs:212, :328, :742):
 * (H,W,C) u8, copied to `out` [host]. */
int adder_b200_video_running_intensities(adder_b200_video* v, uint8_t* out);

/* Device-resident form of the same step: frames already in HBM, events left in HBM.
 *   d_frames         n_frames frames, each H*W*C bytes densely packed, `frame_stride` bytes apart [device]
 *   d_events         device buffer; frame f's events start at d_events + f*events_stride       [device]
 *   d_chunk_offsets  (n_chunks+1) u32 per frame: exclusive event offset of every chunk, last = total;
 *                    frame f's row at d_chunk_offsets + f*(n_chunks+1)                     [device, may be NULL]
 * Runs asynchronously on the handle's stream; call adder_b200_video_sync() to wait and collect status
 * (capacity / depth overflows are reported there).  n_frames consecutive frames are processed in
 * one submission, exactly as n_frames calls of integrate_matrix would. */
int adder_b200_video_integrate_frames_device(adder_b200_video* v, const uint8_t* d_frames, size_t frame_stride,
                                             uint32_t n_frames, float time_spanned, adder_event_t* d_events,
                                             size_t events_stride, uint32_t* d_chunk_offsets);
int adder_b200_video_sync(adder_b200_video* v);
/* The CUDA stream (cudaStream_t) work is queued on, for event timing by the caller. */
void* adder_b200_video_stream(adder_b200_video* v);
/* Number of kernels this handle has launched so far. */
uint64_t adder_b200_video_launch_count(const adder_b200_video* v);

/* Cumulative number of events emitted by this handle (all frames, both forms of the step). */
int adder_b200_video_events_emitted(adder_b200_video* v, uint64_t* out);
/* Device error word accumulated by the kernels since the last sync (ADDER_DEVERR_* bits OR-ed),
 * as sync() saw it; sync() maps it to ADDER_ERR_CAPACITY / _ARENA_DEPTH / _INTERNAL. */

/* Batched host-buffer form: n_frames consecutive calls of integrate_matrix in one submission, with
 * the H2D copy of frame f+1, the kernel of frame f and the D2H copy of frame f-1's events
 * overlapped on three streams (what a caller looping on Source::consume() would want, simulproc.rs:229-277).
 *   frames           n_frames frames of H*W*C bytes, `frame_stride` bytes apart (pinned memory for
 *                    full PCIe rate: adder_b200_host_alloc)                                      [host]
 *   events_out       all frames' events back to back, capacity `events_cap` records              [host]
 *   frame_counts     n_frames entries: events of each frame                          [host, may be NULL]
 *   chunk_counts     n_frames * n_chunks entries                                     [host, may be NULL]
 *   n_events         total
 * When `events_cap` is too small the call returns ADDER_ERR_CAPACITY and reports in *frames_done how many frames
 * were delivered (frame_counts / chunk_counts / events_out are valid for those).  Up to three further frames may
 * already have been integrated (the pipeline's depth); their events are kept in the handle and the frames are not
 * integrated twice: call again with frames + frames_done * frame_stride and n_frames - frames_done — the same
 * frames — and a buffer with room; that call first delivers the kept frames, then goes on integrating.  Until then
 * integrate_matrix / integrate_frames_device fail with ADDER_ERR_BAD_PARAMS; reset_state drops the kept events. */
int adder_b200_video_integrate_frames_host(adder_b200_video* v, const uint8_t* frames, size_t frame_stride,
                                           uint32_t n_frames, float time_spanned, adder_event_t* events_out,
                                           size_t events_cap, uint64_t* frame_counts, uint32_t* chunk_counts,
                                           uint64_t* n_events, uint32_t* frames_done);

/* ---- feature detection: handle_features inside integrate_matrix (video.rs:744, :883-1113) -----
 * Video::update_detect_features, video.rs:825-837 (show_features and feature_cluster only draw on the
 * GUI frame and are not part of this library).  While on, every integrate call ends with is_feature
 * (utils/cv.rs:22-212, FAST 9_16 on channel 0 of running_intensities) for the pixels that fired, the
 * update of the feature sets, and — with feature_rate_adjustment and a non-zero feature_c_radius in the
 * CRF parameters — c_thresh = min(c_thresh_baseline, 2) around every newly found feature.
 * Not available on a row band (set_row_offset != 0): the 7x7 neighbourhood would cross GPUs. */
int adder_b200_video_update_detect_features(adder_b200_video* v, int detect_features, int feature_rate_adjustment);
/* The features newly inserted by the last integrated frame (the reference's `new_features`, video.rs:919-923):
 * up to `cap` [x, y] pairs to xy_out [host], unordered; *n = how many there were. */
int adder_b200_video_new_features(adder_b200_video* v, uint16_t* xy_out, size_t cap, uint32_t* n);
/* VideoState.features as a mask: H*W bytes, 1 where (x, y) is in its chunk's feature set [host]. */
int adder_b200_video_feature_mask(adder_b200_video* v, uint8_t* out);

/* ---- colour source, gray transcode (the step before the path: framed.rs:129 handle_color) -----
 * Framed::new(.., color_input = false, ..) builds a one-channel Video but its decoder still yields
 * three-channel frames, which handle_color (utils/cv.rs:215-232) folds to gray on the CPU.  With
 * source_channels = 3 on a one-channel video every frame handed to integrate_matrix /
 * integrate_frames_host / integrate_frames_host_raw / integrate_frames_device is H*W*3 bytes and the
 * same conversion runs on the device in front of the integrate kernel.  source_channels = 0 or the
 * video's own channel count turns it off (default). */
int adder_b200_video_set_source_channels(adder_b200_video* v, uint8_t source_channels);
/* The gray frame the last integrate call worked on (Framed.input_frame, framed.rs:129, :163-169): H*W bytes to `out` [host]. */
int adder_b200_video_input_frame(adder_b200_video* v, uint8_t* out);

/* ---- raw .adder output (the step after the path: video.rs:736-740 feeding RawOutput) ---------
 * Wire format: bincode fixint big-endian (encoder.rs:64-66). */
#define ADDER_RAW_HEADER_MAX 37u
#define ADDER_RAW_EOF_BYTES 11u
/* EventStreamHeader + extensions for this video's plane and time parameters (codec/header.rs:14-85,
 * encoder.rs:170-229).  version 0..3 (LATEST_CODEC_VERSION = 3, codec/mod.rs:74) -> 25/29/33/37 bytes;
 * source_camera = SourceCamera variant index (FramedU8 = 0, lib.rs:35-47).  Host memory only. */
int adder_b200_video_raw_header(const adder_b200_video* v, uint8_t version, uint32_t source_camera, uint32_t adu_interval,
                                uint8_t* out, size_t cap, size_t* n_bytes);
/* RawOutput::into_writer's EOF event (raw/stream.rs:79-92): always the 11-byte form. */
int adder_b200_raw_eof(uint8_t* out, size_t cap, size_t* n_bytes);
/* Bytes per event on the wire for this plane: 9 (one channel) or 11 (header.rs:77-81). */
int adder_b200_video_raw_event_size(const adder_b200_video* v);
/* RawOutput::ingest_event (raw/stream.rs:100-120) for events already in HBM, queued on the handle's
 * stream: the first min(*d_n_events, n_events_max) records of d_events -> d_out (4-byte aligned,
 * event_size bytes each).  d_n_events is a device word, e.g. the last entry of a frame's chunk offsets. */
int adder_b200_video_raw_encode_device(adder_b200_video* v, const adder_event_t* d_events, const uint32_t* d_n_events,
                                       uint64_t n_events_max, uint8_t* d_out);
/* adder_b200_video_integrate_frames_host, delivering the raw stream body instead of records: what
 * Framed::consume + the raw encoder produce for n_frames frames, copies and kernels overlapped.
 * bytes_out receives event_size bytes per event, all frames back to back (no header, no EOF). */
int adder_b200_video_integrate_frames_host_raw(adder_b200_video* v, const uint8_t* frames, size_t frame_stride,
                                               uint32_t n_frames, float time_spanned, uint8_t* bytes_out, size_t bytes_cap,
                                               uint64_t* frame_counts, uint32_t* chunk_counts, uint64_t* n_bytes,
                                               uint32_t* frames_done);

/* ---- INSTANTANEOUS framer: events -> u8 frames (framer/driver.rs; what SimulProcessor runs downstream of
 * Framed::consume, utils/simulproc.rs:166-218) -------------------------------------------------------------
 * FramerBuilder::new(plane, chunk_rows).codec_version(v, time_mode).time_parameters(tps, ref, dtm, output_fps)
 * .mode(INSTANTANEOUS).view_mode(view).source(U8, source_camera).buffer_limit(limit).finish::<u8>()
 * (driver.rs:36-147, :300-399).  output_fps <= 0 means None (ticks per frame = ref_interval); buffer_limit < 0 means
 * None.  ring_frames = output frames that can be pending at once (0 = 4 * delta_t_max / tpf + 64; two bytes per pixel-channel each); an event that reaches
 * further ahead than that is reported as ADDER_ERR_CAPACITY (the reference grows its VecDeque without bound). */
typedef struct adder_b200_framer adder_b200_framer;
int adder_b200_framer_create(uint16_t width, uint16_t height, uint8_t channels, uint32_t chunk_rows, uint8_t codec_version,
                             int time_mode, uint32_t tps, uint32_t ref_interval, uint32_t delta_t_max, float output_fps,
                             int view_mode, uint32_t source_camera, int64_t buffer_limit, uint32_t ring_frames, int device,
                             adder_b200_framer** out);
void adder_b200_framer_destroy(adder_b200_framer* f);
/* Framer::ingest_events_events (driver.rs:564-626): one Vec<Event> per chunk, given as the records in chunk order
 * plus n_chunks+1 exclusive offsets — the form adder_b200_video_integrate_frames_device leaves in HBM.  Within one
 * call a pixel-channel's events must be contiguous (true of the transcoder's stream).  *frame_ready receives
 * is_frame_0_filled() (driver.rs:851-866). */
int adder_b200_framer_ingest_events_device(adder_b200_framer* f, const adder_event_t* d_events, const uint32_t* d_chunk_offsets,
                                           int* frame_ready);
/* The same without the wait: the kernels (events, front-frame status, chunk_filled_tracker) are queued on the framer's
 * stream and the call returns; the d_events / d_chunk_offsets buffers must stay untouched until a later call of this
 * library on the framer has synchronised.  adder_b200_framer_frame_ready then gives is_frame_0_filled() as of the last
 * ingest — for callers that feed several transcoded frames before they look for an output frame. */
int adder_b200_framer_ingest_events_device_async(adder_b200_framer* f, const adder_event_t* d_events, const uint32_t* d_chunk_offsets);
int adder_b200_framer_frame_ready(adder_b200_framer* f, int* frame_ready);
/* The same from host memory: `events` holds sum(chunk_counts) records, chunk after chunk. */
int adder_b200_framer_ingest_events_host(adder_b200_framer* f, const adder_event_t* events, const uint32_t* chunk_counts,
                                         int* frame_ready);
/* FrameSequence::write_multi_frame_bytes (driver.rs:971-982): pops every frame whose chunks are all filled
 * (None pixels read as 0), H*W*C bytes each, to frames_out [host]; *n_frames = how many.  With max_frames too
 * small the remaining filled frames stay queued for the next call. */
int adder_b200_framer_write_multi_frame_bytes(adder_b200_framer* f, uint8_t* frames_out, uint32_t max_frames, uint32_t* n_frames);
/* Framer::flush_frame_buffer (driver.rs:633-680). */
int adder_b200_framer_flush_frame_buffer(adder_b200_framer* f, int* frame_ready);
/* FrameSequenceState.frames_written and .tpf (driver.rs:232-238). */
int adder_b200_framer_state(const adder_b200_framer* f, int64_t* frames_written, uint32_t* tpf);

/* Back to the state of a fresh Video::new (video.rs:350-438) with the current parameters kept:
 * what adder-viz does on EOF by rebuilding the source (adder.rs:151-166). */
int adder_b200_video_reset_state(adder_b200_video* v);

/* ---- inspection (tests / debugging): one pixel's state, PixelArena fields :53-66 ------------- */
typedef struct adder_b200_px_node {
  float integration, delta_t, best_delta_t;
  uint8_t d, best_d, has_best, reserved;
} adder_b200_px_node_t;
typedef struct adder_b200_px_state {
  float last_fired_t, running_t;
  uint8_t base_val, c_thresh, c_increase_counter, length, dtm_reached, popped_dtm, time_mode, reserved;
  adder_b200_px_node_t nodes[31];
} adder_b200_px_state_t;
int adder_b200_video_read_px(adder_b200_video* v, size_t index, adder_b200_px_state_t* out);

/* ---- plumbing for callers without a CUDA runtime of their own --------------------------------- */
/* Page-locked host memory (cudaHostAlloc) so H2D/D2H run at full PCIe rate. */
int adder_b200_host_alloc(size_t bytes, void** out);
int adder_b200_host_free(void* p);
/* Device memory on the handle's device, and copies ordered on the handle's stream (synchronous to the host). */
int adder_b200_device_alloc(adder_b200_video* v, size_t bytes, void** out);
int adder_b200_device_free(adder_b200_video* v, void* p);
int adder_b200_copy_to_device(adder_b200_video* v, void* dst, const void* src, size_t bytes);
int adder_b200_copy_to_host(adder_b200_video* v, void* dst, const void* src, size_t bytes);
/* CUDA-event stopwatch on the handle's stream: start records an event, stop records another, waits
 * for it and returns the device time between the two in milliseconds. */
int adder_b200_video_timer_start(adder_b200_video* v);
int adder_b200_video_timer_stop(adder_b200_video* v, float* ms);
/* Synthetic frames generated on the device (bench input; tests/synth.py holds the same generator in
 * numpy): n_frames frames of `px` bytes, frame index starting at f0.  kind: 0 gradient (x+2y+3f)&255,
 * 1 uniform noise h(seed,f,i), 2 base +-10 jitter, 3 static base with rare changes (p = 2/256).
 * On a row band (adder_b200_video_set_row_offset) the band's rows of the undivided frame are produced. */
int adder_b200_synth_frames(adder_b200_video* v, uint8_t* d_frames, size_t frame_stride, uint32_t f0,
                            uint32_t n_frames, int kind, uint64_t seed);

/* ---- compact host form: half the bytes across PCIe ---------------------------------------------
 * The event records of a frame are in raster order, so their coordinates are redundant on the way to the host.
 * integrate_frames_host_compact delivers, per frame, back to back in bytes_out:
 *   dense  form, when 4 * E > P:   P count bytes (events of every pixel-channel, raster order), then E x {d:u8, t:u32 LE}
 *   sparse form, otherwise:        E x {index:u32 LE (flat raster index (y*W + x)*C + c inside this plane), d:u8, t:u32 LE}
 * with E = frame_counts[f] and P = W*H*C of this plane: adder_b200_compact_frame_bytes(P, E) bytes.  On uniform noise
 * that is 5.7 instead of 11.3 bytes per pixel-frame.  Everything else (pipelining, capacity / resume contract,
 * chunk_counts) is as for adder_b200_video_integrate_frames_host.  adder_b200_expand_compact turns one frame's block
 * back into the 12-byte records on n_threads host threads (a host-side helper: it needs no device); a consumer that
 * feeds an encoder can also walk the block directly. */
int adder_b200_video_integrate_frames_host_compact(adder_b200_video* v, const uint8_t* frames, size_t frame_stride,
                                                   uint32_t n_frames, float time_spanned, uint8_t* bytes_out, size_t bytes_cap,
                                                   uint64_t* frame_counts, uint32_t* chunk_counts, uint64_t* n_bytes,
                                                   uint32_t* frames_done);
uint64_t adder_b200_compact_frame_bytes(uint64_t n_px, uint64_t n_events);
int adder_b200_expand_compact(uint16_t width, uint16_t rows, uint8_t channels, uint16_t row0, const uint8_t* block, uint64_t n_events,
                              adder_event_t* events_out, uint32_t n_threads);

/* ================================================================================================
 * Event exchange between row bands.
 *
 * The path shards by rows: band g of a frame is a video of its own (adder_b200_video_set_row_offset) on its own GPU,
 * and rank-order concatenation of the bands' streams IS the reference's stream.  Nothing has to be exchanged unless ONE
 * downstream consumer needs the whole frame's events in order — which is what the reference's encoder loop does
 * (video.rs:736-740, serial ingest_event over the Vec<Vec<Event>> of the whole frame; the compressed encoder cuts ADUs
 * out of that ordered stream, adder-codec-core/src/codec/compressed/stream.rs:264-313).  For that case the consumer GPU
 * owns a ring of whole-frame buffers and every band stores its compacted records straight into its place in the
 * frame over NVLink (peer memory: CUDA IPC between processes, plain peer access inside one): the band's offset is the
 * sum of the lower bands' totals, read from a small table in the consumer's memory (an inter-GPU look-back).  No host
 * round trip, no staging copy; pushes run on the comm's own stream behind the band's integrate launch and overlap the
 * next one.  Frames are numbered by the caller (`frame_seq`, the same on all bands); frame s uses ring slot s % slots.
 *
 * One process per GPU (the usual layout):
 *   consumer rank:  comm_create -> comm_export -> (send the blob to the other ranks by any means: it is 256 bytes)
 *                   comm_attach(own band video, consumer comm) for its own band
 *   other ranks:    comm_open(band video, blob)
 *   every frame / batch of frames, every rank:   integrate_frames_device(...); comm_push_frames(...)
 *     (several bands driven by ONE process: push them in rank order — a push waits, on the device, for the totals of the
 *      lower bands, and kernels of one process can queue behind each other whatever their streams)
 *   consumer:       comm_wait_frames(seq0, n); read comm_frame(seq) on the comm's stream; comm_release_frames(seq0 + n)
 */
typedef struct adder_b200_comm adder_b200_comm;
#define ADDER_COMM_BLOB_BYTES 256u

/* Consumer side.  On `v`'s device: `slots` whole-frame buffers of `out_stride` records and total_chunks + 1 chunk
 * offsets each (total_chunks = ceil(H / chunk_rows) of the undivided frame), for `world` bands. */
int adder_b200_comm_create(adder_b200_video* v, uint32_t world, uint32_t total_chunks, uint32_t slots, size_t out_stride,
                           adder_b200_comm** out);
/* The consumer's ring as bytes another process can open (cap >= ADDER_COMM_BLOB_BYTES). */
int adder_b200_comm_export(adder_b200_comm* c, uint8_t* blob, size_t cap);
/* Producer side in another process: maps the consumer's ring into the device of the band video `v`. */
int adder_b200_comm_open(adder_b200_video* v, const uint8_t* blob, size_t blob_bytes, adder_b200_comm** out);
/* Producer side in the consumer's own process (its own band, or bands on other GPUs of the same process). */
int adder_b200_comm_attach(adder_b200_video* v, adder_b200_comm* consumer, adder_b200_comm** out);
void adder_b200_comm_destroy(adder_b200_comm* c);
/* Deliver n_frames frames of this band — the buffers an adder_b200_video_integrate_frames_device call on the band's
 * video has just been given (same d_events / events_stride / d_chunk_offsets, which is required here) — as frames
 * frame_seq0 .. frame_seq0 + n_frames - 1 of the consumer's ring.  band = this band's rank, chunk0 = index of its
 * first chunk in the undivided frame.  Asynchronous: queued on the comm's stream behind the band's stream.
 * Two pushes may be outstanding: work given to the band's stream after call k+1 waits for push k, so a band that
 * alternates two sets of buffers overlaps push k+1 with the integrate launch k+2 without further synchronisation
 * (with one set of buffers, adder_b200_comm_sync before reusing it). */
int adder_b200_comm_push_frames(adder_b200_comm* c, uint32_t band, uint32_t chunk0, const adder_event_t* d_events, size_t events_stride,
                                const uint32_t* d_chunk_offsets, uint32_t n_frames, uint64_t frame_seq0);
/* Consumer: queue a wait on the comm's stream until all bands have delivered frames frame_seq0 .. +n_frames-1. */
int adder_b200_comm_wait_frames(adder_b200_comm* c, uint64_t frame_seq0, uint32_t n_frames);
/* Consumer: where frame frame_seq lies: its events (raster order of the whole frame) and its total_chunks + 1
 * exclusive chunk offsets (the last one is the frame's event count) [device pointers]. */
int adder_b200_comm_frame(adder_b200_comm* c, uint64_t frame_seq, adder_event_t** d_events, uint32_t** d_chunk_offsets);
/* Consumer: frames below upto_seq have been read (by work queued on the comm's stream): their slots may be reused. */
int adder_b200_comm_release_frames(adder_b200_comm* c, uint64_t upto_seq);
/* Wait for everything queued on the comm's stream; reports a frame that did not fit its slot (ADDER_ERR_CAPACITY) or a
 * peer that did not show up within 20 s (ADDER_ERR_INTERNAL). */
int adder_b200_comm_sync(adder_b200_comm* c);
/* The comm's CUDA stream (cudaStream_t), for the consumer's own kernels that read the frames. */
void* adder_b200_comm_stream(adder_b200_comm* c);

#ifdef __cplusplus
}
#endif
#endif /* ADDER_B200_H */
