/*
 * adder_oracle.h — CPU restatement of the reference's framed→ADΔER per-pixel path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (adder_codec_rs_b200/, include/) may call,
 * link or import this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, as the checker / the timed CPU baseline.
 *
 * The reference is Rust and cannot be built in this image (no cargo/rustc, no network, needs
 * system ffmpeg): this is a restatement ("port"), not the reference binary.  Parity is PINNED by
 *   (1) the reference's 13 known-answer unit tests, event_pixel_tree.rs:534-1259, restated in
 *       tests/test_oracle_kat.py, and
 *   (2) the reference's own fixture pair lake_scaled_hd_crop.mp4 -> lake_scaled_hd_out.adder
 *       (adder_simulproc.rs:169-268 `dark`), see tests/golden/ and tests/test_oracle_golden.py
 *       (soft golden: video-decoder rounding differs for a minority of pixels).
 * One view mode of the display byte (FramedViewMode::D) depends on fast_math::log2_raw from the
 * un-vendored crate fast-math 0.1 (adder-codec-rs/Cargo.toml:45, no Cargo.lock): for that view
 * mode only, parity is UNPINNED (the published polynomial is restated in oracle_log2_raw()).
 *
 * All citations are relative to /root/reference/.
 */
#ifndef ADDER_ORACLE_H
#define ADDER_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/adder_b200.h" /* adder_event_t, enums, constants: types only */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- one pixel: PixelArena, adder-codec-rs/src/transcoder/event_pixel_tree.rs:53-66 ---------- */
typedef struct oracle_node {
  float integration; /* PixelState.integration :37 */
  float delta_t;     /* PixelState.delta_t     :38 */
  float best_delta_t;/* best_event.delta_t     :20 */
  uint8_t d;         /* PixelState.d           :36 */
  uint8_t has_best;  /* best_event.is_some()   :48 */
  uint8_t best_d;    /* best_event.d           :19 */
  uint8_t alt;       /* alt.is_some()          :45 (debug-assert only) */
} oracle_node;

#define ORACLE_INLINE_NODES 6 /* SmallVec<[PixelNode; 6]> :61 */

typedef struct oracle_px {
  uint16_t x, y;
  uint8_t c; /* ADDER_C_NONE = None */
  uint8_t time_mode;
  uint8_t base_val;
  uint8_t need_to_pop_top;
  uint8_t c_thresh;
  uint8_t c_increase_counter;
  uint8_t dtm_reached;
  uint8_t popped_dtm;
  float last_fired_t;
  float running_t;
  uint32_t length;    /* live nodes */
  uint32_t arena_len; /* SmallVec len (high-water mark) */
  uint32_t arena_cap; /* heap capacity when spilled, else ORACLE_INLINE_NODES */
  oracle_node* heap;  /* NULL while inline */
  oracle_node inl[ORACLE_INLINE_NODES];
} oracle_px;

void oracle_px_init(oracle_px* px, float start_intensity, uint16_t x, uint16_t y, uint8_t c); /* :69-87 */
void oracle_px_free(oracle_px* px);
oracle_px* oracle_px_new(float start_intensity, uint16_t x, uint16_t y, uint8_t c); /* heap-allocated, for ctypes */
void oracle_px_delete(oracle_px* px);
void oracle_px_time_mode(oracle_px* px, int time_mode); /* :89-93; <0 = None */
const oracle_node* oracle_px_node(const oracle_px* px, uint32_t idx);
uint32_t oracle_px_length(const oracle_px* px);
int oracle_px_need_to_pop_top(const oracle_px* px);

uint8_t oracle_get_d_from_intensity(float intensity); /* :482-499 */
/* :139-148 (+ :151-210, :113-137) */
adder_event_t oracle_px_pop_top_event(oracle_px* px, float next_intensity, int mode, uint32_t ref_time);
/* :213-287; appends to out[*n..cap), returns number appended (or -1 if cap too small) */
int oracle_px_pop_best_events(oracle_px* px, adder_event_t* out, size_t cap, int mode, int multi_mode,
                              uint32_t ref_time, float intensity);
/* :289-312; returns 1 and fills *ev if an event is produced */
int oracle_px_set_d_for_continuous(oracle_px* px, float next_intensity, uint32_t ref_time, adder_event_t* ev);
/* :317-413 */
void oracle_px_integrate(oracle_px* px, float intensity, float time, int mode, uint32_t dtm, uint32_t ref_time,
                         uint8_t c_thresh_max, uint8_t c_increase_velocity, int multi_mode);

/* ---- growable event vector (stands for Vec<Event>) ------------------------------------------- */
typedef struct oracle_evec {
  adder_event_t* data;
  size_t len, cap;
} oracle_evec;

/* integrate_for_px, adder-codec-rs/src/transcoder/source/video.rs:1317-1380 */
int oracle_integrate_for_px(oracle_px* px, uint8_t* base_val, uint8_t frame_val, float intensity, float time_spanned,
                            oracle_evec* buffer, int pixel_tree_mode, int pixel_multi_mode, uint32_t delta_t_max,
                            uint32_t ref_time, uint8_t c_thresh_max, uint8_t c_increase_velocity);

/* u8::get_frame_value, adder-codec-rs/src/framer/scale_intensity.rs:58-104 (+ :262-270) */
uint8_t oracle_get_frame_value_u8(uint8_t d, uint32_t t, double tpf, float practical_d_max, uint32_t delta_t_max,
                                  int view_mode, uint32_t sae_running_t, uint32_t sae_last_fired_t);
float oracle_log2_raw(float x); /* fast-math 0.1 log2_raw restated (UNPINNED) */

/* ---- the video: Video<W>, video.rs:322-345, state :186-243 ----------------------------------- */
typedef struct oracle_video oracle_video;

oracle_video* oracle_video_new(uint16_t w, uint16_t h, uint8_t c, int pixel_tree_mode); /* :350-438 */
void oracle_video_delete(oracle_video* v);
void oracle_video_chunk_rows(oracle_video* v, uint32_t chunk_rows);                         /* :473-481 */
int oracle_video_time_parameters(oracle_video* v, uint32_t tps, uint32_t ref_time, uint32_t dtm, int time_mode); /* :493-537; returns 1 if applied */
void oracle_video_write_out(oracle_video* v, int time_mode, int pixel_multi_mode);          /* :546-636 */
void oracle_video_update_crf(oracle_video* v, uint8_t crf);                                  /* :1241-1251 */
void oracle_video_update_quality_manual(oracle_video* v, uint8_t c_base, uint8_t c_max, uint32_t dtm_mult,
                                        uint8_t velocity, float radius);                    /* :1264-1287 */
void oracle_video_set_crf_parameters(oracle_video* v, const adder_crf_parameters_t* p);     /* :1289-1291, :553 */
void oracle_video_update_delta_t_max(oracle_video* v, uint32_t dtm);                        /* :819-822 */
void oracle_video_c_thresh_pos(oracle_video* v, uint8_t c);                                  /* :445-455 */
void oracle_video_set_c_thresh_rect(oracle_video* v, uint16_t x0, uint16_t y0, uint16_t x1, uint16_t y1, uint8_t value); /* :865-881 */
void oracle_video_set_view_mode(oracle_video* v, int view_mode);
void oracle_video_set_in_interval_count(oracle_video* v, uint32_t n);
uint32_t oracle_video_in_interval_count(const oracle_video* v);
uint32_t oracle_video_n_chunks(const oracle_video* v);
void oracle_crf_parameters(uint8_t crf, uint16_t w, uint16_t h, adder_crf_parameters_t* out); /* rate_controller.rs:55-70 */

/* Video::integrate_matrix, video.rs:651-778 (without the encoder / features / roi tail).
 * frame: (H,W,C) u8 dense.  Events go to the video's internal per-chunk vectors; the call returns
 * the total count.  n_threads: 1 = serial; >1 = OpenMP over chunks, the reference's rayon
 * decomposition (:677-692).  */
size_t oracle_video_integrate_matrix(oracle_video* v, const uint8_t* frame, float time_spanned, int n_threads);
/* Results of the last integrate_matrix: per-chunk lengths and the concatenation (raster order). */
void oracle_video_chunk_counts(const oracle_video* v, uint32_t* counts);
size_t oracle_video_copy_events(const oracle_video* v, adder_event_t* out, size_t cap);
const uint8_t* oracle_video_running_intensities(const oracle_video* v); /* (H,W,C) u8 */
/* Statistics for the roofline's data-dependent terms: live nodes summed over
 * px at entry and exit of the last frame, max live nodes ever, max events of one px in a frame. */
void oracle_video_stats(const oracle_video* v, uint64_t* live_nodes_entry, uint64_t* live_nodes_exit,
                        uint32_t* max_live_nodes, uint32_t* max_px_events);
/* Direct access for state-parity tests. */
const oracle_px* oracle_video_px(const oracle_video* v, size_t index);

int oracle_max_threads(void);
int oracle_num_procs(void); /* host cores available to this process, whatever OMP_NUM_THREADS says */
/* synthetic bench / test frames, identical to tests/synth.py and the device generator */
void oracle_synth_frame(int kind, uint64_t seed, uint32_t f, uint32_t w, uint32_t c, uint64_t i0, uint64_t n, uint8_t* out,
                        int n_threads);

/* ---- feature detection inside integrate_matrix ---------------------------
 * is_feature, utils/cv.rs:22-212 (FAST 9_16 on channel 0; parity UNPINNED by reference tests — none exist;
 * tests/test_features_cpu.py compares it with OpenCV's FAST, of which the reference says it is a port). */
int oracle_is_feature(const uint8_t* img, uint16_t w, uint16_t h, uint8_t channels, uint16_t x, uint16_t y, uint8_t c);
/* Video::update_detect_features, video.rs:825-837; handle_features (:883-1113) then runs at the end of every
 * integrate_matrix: feature sets per chunk, newly found features, c_thresh reset around them (:1077-1104). */
void oracle_video_update_detect_features(oracle_video* v, int detect_features, int feature_rate_adjustment);
size_t oracle_video_new_features(const oracle_video* v, uint16_t* xy_out, size_t cap);
const uint8_t* oracle_video_feature_mask(const oracle_video* v);

/* ---- INSTANTANEOUS framer, events -> u8 frames (framer_oracle.c) ----------- */
typedef struct oracle_framer oracle_framer;
/* FramerBuilder::new(plane, chunk_rows).codec_version(v, time_mode).time_parameters(tps, ref, dtm, output_fps)
 * .mode(INSTANTANEOUS).view_mode(..).source(U8, source_camera).buffer_limit(..).finish::<u8>(), driver.rs:36-147, :300-399 */
oracle_framer* oracle_framer_new(uint16_t w, uint16_t h, uint8_t c, uint32_t chunk_rows, uint8_t codec_version, int time_mode,
                                 uint32_t tps, uint32_t ref_interval, uint32_t delta_t_max, float output_fps, int view_mode,
                                 uint32_t source_camera, int64_t buffer_limit);
void oracle_framer_delete(oracle_framer* f);
int oracle_framer_ingest_event(oracle_framer* f, adder_event_t e);                       /* driver.rs:437-562 */
int oracle_framer_ingest_events_events(oracle_framer* f, const adder_event_t* ev, const uint32_t* chunk_counts, uint32_t n_counts); /* :564-626 */
int oracle_framer_write_multi_frame_bytes(oracle_framer* f, uint8_t* out, size_t cap, size_t* n_bytes); /* :971-982 */
int oracle_framer_flush_frame_buffer(oracle_framer* f);                                 /* :633-680 */
int oracle_framer_bad(const oracle_framer* f);
int64_t oracle_framer_frames_written(const oracle_framer* f);
uint32_t oracle_framer_tpf(const oracle_framer* f);

/* handle_color, adder-codec-rs/src/utils/cv.rs:215-232: (ch0*0.114 + ch1*0.587 + ch2*0.299) as u8 in f64,
 * evaluated left to right, truncating and saturating.  rgb: n_px * 3 bytes, out: n_px bytes. */
void oracle_handle_color(const uint8_t* rgb, size_t n_px, uint8_t* out);

/* ---- raw .adder wire format ---------------------------------
 * bincode 1.3 fixint big-endian, adder-codec-core/src/codec/encoder.rs:64-66, raw/stream.rs:34-36. */
/* EventStreamHeader + extensions V0..V3: codec/header.rs:14-85, encoder.rs:170-229.
 * Returns the bytes written: 25 (v0), 29 (v1), 33 (v2), 37 (v3); 0 for an unknown version. */
size_t oracle_raw_header(uint8_t* out, int compressed, uint8_t version, uint16_t width, uint16_t height, uint32_t tps,
                         uint32_t ref_interval, uint32_t delta_t_max, uint8_t channels, uint32_t source_camera,
                         uint32_t time_mode, uint32_t adu_interval);
/* RawOutput::ingest_event for n events, raw/stream.rs:100-120: EventSingle (9 bytes) when the plane
 * has one channel, Event with c = Some (11 bytes) otherwise.  Returns the bytes written. */
size_t oracle_raw_encode(const adder_event_t* ev, size_t n, uint8_t channels, uint8_t* out);
/* RawOutput::into_writer's EOF event, raw/stream.rs:79-92: always the 11-byte form. */
size_t oracle_raw_eof(uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
