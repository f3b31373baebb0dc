//! The replacement for the per-pixel part of `Video<W>` (adder-codec-rs/src/transcoder/source/video.rs).
//!
//! NOT COMPILED in this repository's build image (no cargo / rustc there).  It is what a maintainer of
//! ac-freeman/adder-codec-rs would add as `src/transcoder/source/video_b200.rs`: `event_pixel_trees:
//! Array3<PixelArena>` (video.rs:325) becomes `gpu: B200Video`, the body of `integrate_matrix` (video.rs:651-778)
//! becomes `B200Video::integrate_matrix`, and every setter that touches per-pixel state forwards one call.
//! `Source<W>` (video.rs:1418-1442), `VideoBuilder<W>` (:272-317), `Framed` (framed.rs), adder-viz and SimulProcessor
//! keep their signatures.  tests/c_abi/consumer.c does the same work from C and is checked against the oracle.

use adder_b200_sys as ffi;
use adder_codec_core::{Coord, Event, PlaneSize};
use rayon::prelude::*; // already a dependency of adder-codec-rs (video.rs:677-734)
use std::ffi::CStr;

use crate::transcoder::source::video::SourceError;

/// Owned device handle: stands for `Array3<PixelArena>` + the parts of `VideoState` the kernels read.
pub struct B200Video {
    raw: *mut ffi::adder_b200_video,
    n_chunks: usize,
    /// page-locked staging for one frame's records (`adder_b200_host_alloc`): 1.6 ms instead of 7.0 ms per
    /// 1080p RGB call (profiles/r01t_consume_latency.txt)
    ev_buf: *mut ffi::adder_event_t,
    ev_cap: usize,
    counts: Vec<u32>,
}

// `Video: Send` (video.rs:346) stays valid: a handle is used by one thread at a time (`consume(&mut self)`).
unsafe impl Send for B200Video {}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::adder_b200_last_error()) }.to_string_lossy().into_owned()
}

fn check(rc: i32) -> Result<(), SourceError> {
    if rc == ffi::ADDER_OK {
        Ok(())
    } else if rc == ffi::ADDER_ERR_BAD_PARAMS {
        Err(SourceError::BadParams(last_error()))
    } else {
        Err(SourceError::VisionError(last_error()))
    }
}

impl B200Video {
    /// Video::new, video.rs:350-438: every pixel starts as PixelArena::new(1.0, coord).
    pub fn new(plane: PlaneSize, chunk_rows: usize, device: i32) -> Result<Self, SourceError> {
        let mut raw = std::ptr::null_mut();
        check(unsafe {
            ffi::adder_b200_video_create(plane.w(), plane.h(), plane.c(), ffi::ADDER_MODE_FRAME_PERFECT, device, 0, &mut raw)
        })?;
        check(unsafe { ffi::adder_b200_video_chunk_rows(raw, chunk_rows as u32) })?;
        let n_chunks = (plane.h_usize() + chunk_rows - 1) / chunk_rows;
        let ev_cap = plane.volume() * 2;
        let mut p = std::ptr::null_mut();
        check(unsafe { ffi::adder_b200_host_alloc(ev_cap * std::mem::size_of::<ffi::adder_event_t>(), &mut p) })?;
        Ok(Self { raw, n_chunks, ev_buf: p as *mut ffi::adder_event_t, ev_cap, counts: vec![0; n_chunks] })
    }

    /// The body of Video::integrate_matrix (video.rs:651-735): one frame in, the reference's Vec<Vec<Event>> out
    /// (one vector per chunk — the framer asserts that, framer/driver.rs:566).  `set_initial_d` (:656-658) and
    /// `in_interval_count += 1` (:662) happen inside the library; the caller mirrors the counter.
    pub fn integrate_matrix(&mut self, frame: &[u8], time_spanned: f32) -> Result<Vec<Vec<Event>>, SourceError> {
        let mut n: u64 = 0;
        let mut rc = unsafe {
            ffi::adder_b200_video_integrate_matrix(self.raw, frame.as_ptr(), 0, time_spanned, self.ev_buf, self.ev_cap,
                                                   self.counts.as_mut_ptr(), &mut n)
        };
        if rc == ffi::ADDER_ERR_CAPACITY {
            // nothing is lost: the frame's events are still on the device; grow the staging buffer and re-read
            self.grow(n as usize)?;
            rc = unsafe { ffi::adder_b200_video_fetch_events(self.raw, self.ev_buf, self.ev_cap, self.counts.as_mut_ptr(), &mut n) };
        }
        check(rc)?;
        let records = unsafe { std::slice::from_raw_parts(self.ev_buf, n as usize) };
        // 12-byte records -> Event; the packed Option<u8> layout of Event is never aliased.  One vector per chunk, built on
        // the rayon pool the reference already runs this loop on (video.rs:677-734): the chunks' slices are independent.
        // Measured by the C consumer at 1080p RGB (5.6 M events): 13.3 ms on one thread, 2.2 ms on a pool of 16
        // (profiles/r02x_c_consumer_1080p.json).
        let mut first = Vec::with_capacity(self.n_chunks);
        let mut at = 0usize;
        for &len in &self.counts {
            first.push(at);
            at += len as usize;
        }
        let counts = &self.counts;
        let out: Vec<Vec<Event>> = (0..self.n_chunks).into_par_iter().map(|ci| {
            records[first[ci]..first[ci] + counts[ci] as usize].iter().map(|e| Event {
                coord: Coord { x: e.x, y: e.y, c: if e.c == ffi::ADDER_C_NONE { None } else { Some(e.c) } },
                d: e.d,
                t: e.t,
            }).collect()
        }).collect();
        Ok(out)
    }

    /// state.running_intensities (video.rs:212), filled on the device by the same kernel (video.rs:713-730)
    pub fn running_intensities(&mut self, out: &mut [u8]) -> Result<(), SourceError> {
        check(unsafe { ffi::adder_b200_video_running_intensities(self.raw, out.as_mut_ptr()) })
    }

    // ---- setters: each replaces the `par_map_inplace` over event_pixel_trees in the method of the same name ----
    pub fn update_crf(&mut self, crf: u8) -> Result<(), SourceError> { check(unsafe { ffi::adder_b200_video_update_crf(self.raw, crf) }) }                       // video.rs:1241-1251
    pub fn update_quality_manual(&mut self, c_base: u8, c_max: u8, dtm_mult: u32, velocity: u8, radius: f32) -> Result<(), SourceError> {
        check(unsafe { ffi::adder_b200_video_update_quality_manual(self.raw, c_base, c_max, dtm_mult, velocity, radius) })                                   // :1264-1287
    }
    pub fn time_parameters(&mut self, tps: u32, ref_time: u32, dtm: u32, time_mode: Option<i32>) -> Result<bool, SourceError> {
        let mut applied = 0;
        check(unsafe { ffi::adder_b200_video_time_parameters(self.raw, tps, ref_time, dtm, time_mode.unwrap_or(-1), &mut applied) })?;                      // :493-537
        Ok(applied != 0)
    }
    pub fn write_out(&mut self, time_mode: Option<i32>, multi_mode: Option<i32>) -> Result<(), SourceError> {
        check(unsafe { ffi::adder_b200_video_write_out(self.raw, time_mode.unwrap_or(-1), multi_mode.unwrap_or(-1)) })                                       // :546-636
    }
    pub fn update_delta_t_max(&mut self, dtm: u32) -> Result<(), SourceError> { check(unsafe { ffi::adder_b200_video_update_delta_t_max(self.raw, dtm) }) }   // :819-822
    pub fn c_thresh_pos(&mut self, c: u8) -> Result<(), SourceError> { check(unsafe { ffi::adder_b200_video_c_thresh_pos(self.raw, c) }) }                   // :445-455
    pub fn set_c_thresh_rect(&mut self, x0: u16, y0: u16, x1: u16, y1: u16, c: u8) -> Result<(), SourceError> {
        check(unsafe { ffi::adder_b200_video_set_c_thresh_rect(self.raw, x0, y0, x1, y1, c) })                                                              // handle_roi :865-881, feature radius :1089-1104
    }
    pub fn set_in_interval_count(&mut self, n: u32) -> Result<(), SourceError> { check(unsafe { ffi::adder_b200_video_set_in_interval_count(self.raw, n) }) } // adder-viz restart, adder.rs:155-166
    pub fn update_detect_features(&mut self, on: bool, rate_adjustment: bool) -> Result<(), SourceError> {
        check(unsafe { ffi::adder_b200_video_update_detect_features(self.raw, on as i32, rate_adjustment as i32) })                                         // :825-837
    }

    fn grow(&mut self, need: usize) -> Result<(), SourceError> {
        unsafe { ffi::adder_b200_host_free(self.ev_buf as *mut _) };
        self.ev_cap = need.next_power_of_two();
        let mut p = std::ptr::null_mut();
        check(unsafe { ffi::adder_b200_host_alloc(self.ev_cap * std::mem::size_of::<ffi::adder_event_t>(), &mut p) })?;
        self.ev_buf = p as *mut ffi::adder_event_t;
        Ok(())
    }
}

impl Drop for B200Video {
    fn drop(&mut self) {
        unsafe {
            ffi::adder_b200_host_free(self.ev_buf as *mut _);
            ffi::adder_b200_video_destroy(self.raw);
        }
    }
}

// ---- inside `impl<W: Write + 'static> Video<W>` (video.rs), the function the GPU path replaces -------------------
//
// pub(crate) fn integrate_matrix(&mut self, matrix: Frame, time_spanned: f32) -> Result<Vec<Vec<Event>>, SourceError> {
//     self.state.in_interval_count += 1;                                              // video.rs:662 (mirror)
//     let frame = matrix.as_slice().expect("standard layout (H,W,C)");                // row-major, framed.rs:65-71
//     let big_buffer = self.gpu.integrate_matrix(frame, time_spanned)?;               // was :665-734
//     for events in &big_buffer { for e in events { self.encoder.ingest_event(*e)?; } } // :736-740 unchanged
//     self.gpu.running_intensities(self.state.running_intensities.as_slice_mut().unwrap())?;
//     self.display_frame_features = self.state.running_intensities.clone();           // :742 unchanged
//     self.handle_features(&big_buffer)?;                                             // :744, or update_detect_features on the device
//     Ok(big_buffer)
// }
