"""Framed — host-side mirror of the reference's `Framed<W>` source
(adder-codec-rs/src/transcoder/source/framed.rs:22-280) over the CUDA `Video`.

Same builder / Source method names and argument meaning as the reference
(`VideoBuilder`: video.rs:272-317, `Source`: video.rs:1418-1442).  What is NOT here, by scope
(SURVEY.md §2 #3): the ffmpeg decode — frames come from any iterator of (H, W, 3) or (H, W, C) u8
arrays instead of a file path — and the encoder behind `write_out` (only its effects on the
transcode state are kept: time mode, pixel multi mode, CRF parameter set).
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional

import numpy as np

from . import binding as B


class SourceError(Exception):
    """SourceError, video.rs:55-122."""


class BadParams(SourceError):
    pass


class StartOutOfBounds(SourceError):
    pass


class NoData(SourceError):
    """The frame source is exhausted (the reference surfaces the decoder's EOF error)."""


class Framed:
    def __init__(self, frames: Iterable[np.ndarray], width: int, height: int, color_input: bool,
                 source_fps: float = 30.0, frame_count: Optional[int] = None, device: int = 0, max_depth: int = 0):
        """Framed::new, framed.rs:44-78 (decoder replaced by `frames`)."""
        self._frames_src = frames
        self._it: Iterator[np.ndarray] = iter(frames)
        self.frame_idx_start = 0
        self.source_fps = float(source_fps)
        self.color_input = bool(color_input)
        self.frame_count = frame_count if frame_count is not None else (len(frames) if hasattr(frames, "__len__") else None)
        self._input_frame = np.zeros((height, width, 3 if color_input else 1), dtype=np.uint8)
        self._input_on_device = False
        self.video = B.Video(width, height, 3 if color_input else 1, B.MODE_FRAME_PERFECT, device, max_depth)

    # ---- Framed's own builder methods ----
    def frame_start(self, frame_idx_start: int) -> "Framed":
        """framed.rs:81-91"""
        if self.frame_count is not None and frame_idx_start >= self.frame_count:
            raise StartOutOfBounds(frame_idx_start)
        self._it = iter(self._frames_src)
        for _ in range(frame_idx_start):
            next(self._it)
        self.frame_idx_start = frame_idx_start
        return self

    def auto_time_parameters(self, ref_time: int, delta_t_max: int, time_mode: Optional[int] = None) -> "Framed":
        """framed.rs:94-111: an error (not a warning) unless delta_t_max % ref_time == 0."""
        if delta_t_max % ref_time != 0:
            raise BadParams("delta_t_max must be a multiple of ref_time")
        tps = int(np.float32(ref_time) * np.float32(self.source_fps))
        self.video.time_parameters(tps, ref_time, delta_t_max, time_mode)
        return self

    def get_ref_time(self) -> int:
        return self.video.info().ref_time

    @property
    def input_frame(self) -> np.ndarray:
        """framed.rs:129: the frame after handle_color (fetched from the device when it was made there)."""
        if self._input_on_device:
            self._input_frame = self.video.input_frame()
            self._input_on_device = False
        return self._input_frame

    def get_last_input_frame(self) -> np.ndarray:
        return self.input_frame

    # ---- VideoBuilder, framed.rs:189-279 ----
    def crf(self, crf: int) -> "Framed":
        self.video.update_crf(crf)
        return self

    def quality_manual(self, c_thresh_baseline, c_thresh_max, delta_t_max_multiplier, c_increase_velocity,
                       feature_c_radius_denom) -> "Framed":
        self.video.update_quality_manual(c_thresh_baseline, c_thresh_max, delta_t_max_multiplier, c_increase_velocity,
                                         feature_c_radius_denom)
        return self

    def chunk_rows(self, chunk_rows: int) -> "Framed":
        self.video.chunk_rows(chunk_rows)
        return self

    def time_parameters(self, tps, ref_time, delta_t_max, time_mode=None) -> "Framed":
        """framed.rs:216-231: a bad multiple only warns and keeps the old values."""
        if delta_t_max % ref_time == 0:
            self.video.time_parameters(tps, ref_time, delta_t_max, time_mode)
        else:
            print("delta_t_max must be a multiple of ref_time")
        return self

    def write_out(self, time_mode: int, pixel_multi_mode: int, crf_parameters=None) -> "Framed":
        """framed.rs:233-253 -> video.rs:546-636, state effects only; `crf_parameters`
        (c_base, c_max, velocity[, radius]) stands for the `encoder_options.crf` argument."""
        self.video.write_out(time_mode, pixel_multi_mode)
        if crf_parameters is not None:
            self.video.set_crf_parameters(*crf_parameters)
        return self

    # ---- Source, framed.rs:124-186 ----
    def next_frame(self) -> np.ndarray:
        """The decode half of consume() (framed.rs:128), with handle_color (:129) routed to the device: returns the
        frame exactly as the library wants it (three channels when a colour source feeds a gray transcode)."""
        try:
            frame = next(self._it)
        except StopIteration:
            raise NoData("end of input") from None
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        if frame.ndim == 2:
            frame = frame[..., None]
        want_src = 3 if (frame.shape[-1] == 3 and not self.color_input) else self.video.c
        if want_src != self.video.src_c:
            self.video.set_source_channels(want_src)
        self._input_frame, self._input_on_device = frame, want_src != self.video.c
        return frame

    def consume(self) -> List[np.ndarray]:
        """One input frame -> Vec<Vec<Event>>: a list of n_chunks event arrays (framed.rs:127-157)."""
        frame = self.next_frame()
        ref_time = self.video.info().ref_time
        events, counts = self.video.integrate_matrix(frame, float(ref_time))
        bounds = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)])
        return [events[bounds[i]:bounds[i + 1]] for i in range(len(counts))]

    def get_video_mut(self) -> B.Video:
        return self.video

    def get_video_ref(self) -> B.Video:
        return self.video

    def get_input(self) -> np.ndarray:
        return self.get_last_input_frame()

    def get_running_input_bitrate(self) -> float:
        """framed.rs:180-185"""
        i = self.video.info()
        return i.tps / i.ref_time * (i.width * i.height * i.channels) * 8.0


class RawAdderWriter:
    """The file side of `Encoder::new_raw` + `RawOutput` (codec/encoder.rs:170-229, raw/stream.rs:79-120):
    header at construction, event bytes as they come from `Video.integrate_frames_host_raw`, the 11-byte
    EOF event on close.  The bytes themselves are produced on the device; this only lays them out."""

    def __init__(self, fileobj, video: B.Video, version: int = 3, source_camera: int = 0, adu_interval: int = 0):
        self.f = fileobj
        self.event_size = video.raw_event_size
        self.header = video.raw_header(version, source_camera, adu_interval)
        self.f.write(self.header)
        self.n_events = 0

    def write_body(self, body) -> None:
        b = memoryview(np.ascontiguousarray(body)).cast("B")
        assert len(b) % self.event_size == 0
        self.f.write(b)
        self.n_events += len(b) // self.event_size

    def close(self) -> None:
        self.f.write(B.raw_eof())
        self.f.flush()
