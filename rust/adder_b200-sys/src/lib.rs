//! Raw bindings to `libadder_b200.so` — one declaration per entry point of `include/adder_b200.h`, in the header's order
//! (kept in step with the header by `tests/test_abi_cpu.py::test_rust_bindings_cover_the_header`).
//!
//! NOT COMPILED in this repository's build image (no cargo / rustc there).  The same declarations are exercised from
//! plain C by `tests/c_abi/consumer.c` and from Python/ctypes by `adder_codec_rs_b200/binding.py`; the struct layouts
//! below are checked against the header by `tests/test_abi_cpu.py::test_rust_bindings_cover_the_header`.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_float, c_int, c_void};

pub const ADDER_B200_ABI_VERSION: c_int = 1;
pub const ADDER_C_NONE: u8 = 0xFF;
pub const ADDER_COMM_BLOB_BYTES: usize = 256;

// adder_status
pub const ADDER_OK: c_int = 0;
pub const ADDER_ERR_BAD_PARAMS: c_int = 1;
pub const ADDER_ERR_NO_DEVICE: c_int = 2;
pub const ADDER_ERR_CUDA: c_int = 3;
pub const ADDER_ERR_CAPACITY: c_int = 4;
pub const ADDER_ERR_ARENA_DEPTH: c_int = 5;
pub const ADDER_ERR_UNSUPPORTED: c_int = 6;
pub const ADDER_ERR_INTERNAL: c_int = 7;
pub const ADDER_ERR_NOMEM: c_int = 8;
// modes: numeric values of the reference's enums (adder-codec-core/src/lib.rs:72-83, :196-213; video.rs:140-158)
pub const ADDER_MODE_FRAME_PERFECT: c_int = 0;
pub const ADDER_MODE_CONTINUOUS: c_int = 1;
pub const ADDER_MULTI_NORMAL: c_int = 0;
pub const ADDER_MULTI_COLLAPSE: c_int = 1;
pub const ADDER_TIME_DELTA_T: c_int = 0;
pub const ADDER_TIME_ABSOLUTE_T: c_int = 1;
pub const ADDER_TIME_MIXED: c_int = 2;
pub const ADDER_VIEW_INTENSITY: c_int = 0;
pub const ADDER_VIEW_D: c_int = 1;
pub const ADDER_VIEW_DELTA_T: c_int = 2;
pub const ADDER_VIEW_SAE: c_int = 3;

/// 12-byte record the boundary uses instead of the packed `Event` (whose `Option<u8>` is not a stable ABI).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default, PartialEq, Eq)]
pub struct adder_event_t {
    pub x: u16,
    pub y: u16,
    pub c: u8, // ADDER_C_NONE <=> Coord.c == None
    pub d: u8,
    pub reserved: u16,
    pub t: u32,
}

/// CrfParameters, adder-codec-core/src/codec/rate_controller.rs:40-53
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct adder_crf_parameters_t {
    pub c_thresh_baseline: u8,
    pub c_thresh_max: u8,
    pub c_increase_velocity: u8,
    pub reserved: u8,
    pub feature_c_radius: u16,
    pub reserved2: u16,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct adder_b200_video_info_t {
    pub width: u16,
    pub height: u16,
    pub channels: u8,
    pub pixel_tree_mode: u8,
    pub pixel_multi_mode: u8,
    pub time_mode: u8,
    pub view_mode: u8,
    /// 0 = every level holds its own integration / delta_t, 1 = offset form (see include/adder_b200.h)
    pub state_form: u8,
    pub reserved: [u8; 2],
    pub chunk_rows: u32,
    pub n_chunks: u32,
    pub in_interval_count: u32,
    pub tps: u32,
    pub ref_time: u32,
    pub delta_t_max: u32,
    pub crf: adder_crf_parameters_t,
    pub max_depth: u32,
    pub device: u32,
    pub state_bytes: u64,
    pub events_capacity: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct adder_b200_px_node_t {
    pub integration: f32,
    pub delta_t: f32,
    pub best_delta_t: f32,
    pub d: u8,
    pub best_d: u8,
    pub has_best: u8,
    pub reserved: u8,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct adder_b200_px_state_t {
    pub last_fired_t: f32,
    pub running_t: f32,
    pub base_val: u8,
    pub c_thresh: u8,
    pub c_increase_counter: u8,
    pub length: u8,
    pub dtm_reached: u8,
    pub popped_dtm: u8,
    pub time_mode: u8,
    pub reserved: u8,
    pub nodes: [adder_b200_px_node_t; 31],
}

#[repr(C)]
pub struct adder_b200_video {
    _private: [u8; 0],
}
#[repr(C)]
pub struct adder_b200_framer {
    _private: [u8; 0],
}
#[repr(C)]
pub struct adder_b200_comm {
    _private: [u8; 0],
}

extern "C" {
    pub fn adder_b200_abi_version() -> c_int;
    pub fn adder_b200_last_error() -> *const c_char;
    pub fn adder_b200_device_count() -> c_int;
    pub fn adder_b200_crf_parameters(crf: u8, plane_w: u16, plane_h: u16, out_: *mut adder_crf_parameters_t) -> c_int;
    pub fn adder_b200_video_create(width: u16, height: u16, channels: u8, pixel_tree_mode: c_int, device: c_int, max_depth: u32, out_: *mut *mut adder_b200_video) -> c_int;
    pub fn adder_b200_video_destroy(v: *mut adder_b200_video);
    pub fn adder_b200_video_chunk_rows(v: *mut adder_b200_video, chunk_rows: u32) -> c_int;
    pub fn adder_b200_video_time_parameters(v: *mut adder_b200_video, tps: u32, ref_time: u32, delta_t_max: u32, time_mode: c_int, applied: *mut c_int) -> c_int;
    pub fn adder_b200_video_write_out(v: *mut adder_b200_video, time_mode: c_int, pixel_multi_mode: c_int) -> c_int;
    pub fn adder_b200_video_update_crf(v: *mut adder_b200_video, crf: u8) -> c_int;
    pub fn adder_b200_video_update_quality_manual(v: *mut adder_b200_video, c_thresh_baseline: u8, c_thresh_max: u8, delta_t_max_multiplier: u32, c_increase_velocity: u8, feature_c_radius: c_float) -> c_int;
    pub fn adder_b200_video_set_crf_parameters(v: *mut adder_b200_video, params: *const adder_crf_parameters_t) -> c_int;
    pub fn adder_b200_video_update_delta_t_max(v: *mut adder_b200_video, delta_t_max: u32) -> c_int;
    pub fn adder_b200_video_c_thresh_pos(v: *mut adder_b200_video, c: u8) -> c_int;
    pub fn adder_b200_video_set_c_thresh_rect(v: *mut adder_b200_video, x0: u16, y0: u16, x1: u16, y1: u16, value: u8) -> c_int;
    pub fn adder_b200_video_set_view_mode(v: *mut adder_b200_video, view_mode: c_int) -> c_int;
    pub fn adder_b200_video_set_in_interval_count(v: *mut adder_b200_video, n: u32) -> c_int;
    pub fn adder_b200_video_set_row_offset(v: *mut adder_b200_video, row0: u16) -> c_int;
    pub fn adder_b200_video_set_counting(v: *mut adder_b200_video, on: c_int) -> c_int;
    pub fn adder_b200_video_read_counters(v: *mut adder_b200_video, out_: *mut u64) -> c_int;
    pub fn adder_b200_video_get_info(v: *const adder_b200_video, out_: *mut adder_b200_video_info_t) -> c_int;
    pub fn adder_b200_video_integrate_matrix(v: *mut adder_b200_video, frame: *const u8, row_pitch: usize, time_spanned: c_float, events_out: *mut adder_event_t, events_cap: usize, chunk_counts: *mut u32, n_events: *mut u64) -> c_int;
    pub fn adder_b200_video_fetch_events(v: *mut adder_b200_video, events_out: *mut adder_event_t, events_cap: usize, chunk_counts: *mut u32, n_events: *mut u64) -> c_int;
    pub fn adder_b200_video_running_intensities(v: *mut adder_b200_video, out_: *mut u8) -> c_int;
    pub fn adder_b200_video_integrate_frames_device(v: *mut adder_b200_video, d_frames: *const u8, frame_stride: usize, n_frames: u32, time_spanned: c_float, d_events: *mut adder_event_t, events_stride: usize, d_chunk_offsets: *mut u32) -> c_int;
    pub fn adder_b200_video_sync(v: *mut adder_b200_video) -> c_int;
    pub fn adder_b200_video_stream(v: *mut adder_b200_video) -> *mut c_void;
    pub fn adder_b200_video_launch_count(v: *const adder_b200_video) -> u64;
    pub fn adder_b200_video_events_emitted(v: *mut adder_b200_video, out_: *mut u64) -> c_int;
    pub fn adder_b200_video_integrate_frames_host(v: *mut adder_b200_video, frames: *const u8, frame_stride: usize, n_frames: u32, time_spanned: c_float, events_out: *mut adder_event_t, events_cap: usize, frame_counts: *mut u64, chunk_counts: *mut u32, n_events: *mut u64, frames_done: *mut u32) -> c_int;
    pub fn adder_b200_video_update_detect_features(v: *mut adder_b200_video, detect_features: c_int, feature_rate_adjustment: c_int) -> c_int;
    pub fn adder_b200_video_new_features(v: *mut adder_b200_video, xy_out: *mut u16, cap: usize, n: *mut u32) -> c_int;
    pub fn adder_b200_video_feature_mask(v: *mut adder_b200_video, out_: *mut u8) -> c_int;
    pub fn adder_b200_video_set_source_channels(v: *mut adder_b200_video, source_channels: u8) -> c_int;
    pub fn adder_b200_video_input_frame(v: *mut adder_b200_video, out_: *mut u8) -> c_int;
    pub fn adder_b200_video_raw_header(v: *const adder_b200_video, version: u8, source_camera: u32, adu_interval: u32, out_: *mut u8, cap: usize, n_bytes: *mut usize) -> c_int;
    pub fn adder_b200_raw_eof(out_: *mut u8, cap: usize, n_bytes: *mut usize) -> c_int;
    pub fn adder_b200_video_raw_event_size(v: *const adder_b200_video) -> c_int;
    pub fn adder_b200_video_raw_encode_device(v: *mut adder_b200_video, d_events: *const adder_event_t, d_n_events: *const u32, n_events_max: u64, d_out: *mut u8) -> c_int;
    pub fn adder_b200_video_integrate_frames_host_raw(v: *mut adder_b200_video, frames: *const u8, frame_stride: usize, n_frames: u32, time_spanned: c_float, bytes_out: *mut u8, bytes_cap: usize, frame_counts: *mut u64, chunk_counts: *mut u32, n_bytes: *mut u64, frames_done: *mut u32) -> c_int;
    pub fn adder_b200_framer_create(width: u16, height: u16, channels: u8, chunk_rows: u32, codec_version: u8, time_mode: c_int, tps: u32, ref_interval: u32, delta_t_max: u32, output_fps: c_float, view_mode: c_int, source_camera: u32, buffer_limit: i64, ring_frames: u32, device: c_int, out_: *mut *mut adder_b200_framer) -> c_int;
    pub fn adder_b200_framer_destroy(f: *mut adder_b200_framer);
    pub fn adder_b200_framer_ingest_events_device(f: *mut adder_b200_framer, d_events: *const adder_event_t, d_chunk_offsets: *const u32, frame_ready: *mut c_int) -> c_int;
    pub fn adder_b200_framer_ingest_events_device_async(f: *mut adder_b200_framer, d_events: *const adder_event_t, d_chunk_offsets: *const u32) -> c_int;
    pub fn adder_b200_framer_frame_ready(f: *mut adder_b200_framer, frame_ready: *mut c_int) -> c_int;
    pub fn adder_b200_framer_ingest_events_host(f: *mut adder_b200_framer, events: *const adder_event_t, chunk_counts: *const u32, frame_ready: *mut c_int) -> c_int;
    pub fn adder_b200_framer_write_multi_frame_bytes(f: *mut adder_b200_framer, frames_out: *mut u8, max_frames: u32, n_frames: *mut u32) -> c_int;
    pub fn adder_b200_framer_flush_frame_buffer(f: *mut adder_b200_framer, frame_ready: *mut c_int) -> c_int;
    pub fn adder_b200_framer_state(f: *const adder_b200_framer, frames_written: *mut i64, tpf: *mut u32) -> c_int;
    pub fn adder_b200_video_reset_state(v: *mut adder_b200_video) -> c_int;
    pub fn adder_b200_video_read_px(v: *mut adder_b200_video, index: usize, out_: *mut adder_b200_px_state_t) -> c_int;
    pub fn adder_b200_host_alloc(bytes: usize, out_: *mut *mut c_void) -> c_int;
    pub fn adder_b200_host_free(p: *mut c_void) -> c_int;
    pub fn adder_b200_device_alloc(v: *mut adder_b200_video, bytes: usize, out_: *mut *mut c_void) -> c_int;
    pub fn adder_b200_device_free(v: *mut adder_b200_video, p: *mut c_void) -> c_int;
    pub fn adder_b200_copy_to_device(v: *mut adder_b200_video, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn adder_b200_copy_to_host(v: *mut adder_b200_video, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn adder_b200_video_timer_start(v: *mut adder_b200_video) -> c_int;
    pub fn adder_b200_video_timer_stop(v: *mut adder_b200_video, ms: *mut c_float) -> c_int;
    pub fn adder_b200_synth_frames(v: *mut adder_b200_video, d_frames: *mut u8, frame_stride: usize, f0: u32, n_frames: u32, kind: c_int, seed: u64) -> c_int;
    pub fn adder_b200_video_integrate_frames_host_compact(v: *mut adder_b200_video, frames: *const u8, frame_stride: usize, n_frames: u32, time_spanned: c_float, bytes_out: *mut u8, bytes_cap: usize, frame_counts: *mut u64, chunk_counts: *mut u32, n_bytes: *mut u64, frames_done: *mut u32) -> c_int;
    pub fn adder_b200_compact_frame_bytes(n_px: u64, n_events: u64) -> u64;
    pub fn adder_b200_expand_compact(width: u16, rows: u16, channels: u8, row0: u16, block: *const u8, n_events: u64, events_out: *mut adder_event_t, n_threads: u32) -> c_int;
    pub fn adder_b200_comm_create(v: *mut adder_b200_video, world: u32, total_chunks: u32, slots: u32, out_stride: usize, out_: *mut *mut adder_b200_comm) -> c_int;
    pub fn adder_b200_comm_export(c: *mut adder_b200_comm, blob: *mut u8, cap: usize) -> c_int;
    pub fn adder_b200_comm_open(v: *mut adder_b200_video, blob: *const u8, blob_bytes: usize, out_: *mut *mut adder_b200_comm) -> c_int;
    pub fn adder_b200_comm_attach(v: *mut adder_b200_video, consumer: *mut adder_b200_comm, out_: *mut *mut adder_b200_comm) -> c_int;
    pub fn adder_b200_comm_destroy(c: *mut adder_b200_comm);
    pub fn adder_b200_comm_push_frames(c: *mut adder_b200_comm, band: u32, chunk0: u32, d_events: *const adder_event_t, events_stride: usize, d_chunk_offsets: *const u32, n_frames: u32, frame_seq0: u64) -> c_int;
    pub fn adder_b200_comm_wait_frames(c: *mut adder_b200_comm, frame_seq0: u64, n_frames: u32) -> c_int;
    pub fn adder_b200_comm_frame(c: *mut adder_b200_comm, frame_seq: u64, d_events: *mut *mut adder_event_t, d_chunk_offsets: *mut *mut u32) -> c_int;
    pub fn adder_b200_comm_release_frames(c: *mut adder_b200_comm, upto_seq: u64) -> c_int;
    pub fn adder_b200_comm_sync(c: *mut adder_b200_comm) -> c_int;
    pub fn adder_b200_comm_stream(c: *mut adder_b200_comm) -> *mut c_void;
}
