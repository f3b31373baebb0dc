"""TEST DRIVER, not part of the product package (SURVEY.md §2 #17 marks the reference's SimulProcessor out of scope):
used by tests/test_gpu_framer.py to run the transcoder and the framer in lock step the way the reference's caller does.

SimulProcessor — host-side mirror of the reference's transcode-and-frame driver
(adder-codec-rs/src/utils/simulproc.rs:87-278): a `Framed` source feeds the INSTANTANEOUS framer frame by frame and
the reconstructed frames are written out as they fill.  Both halves run on the device; between them the events stay in
HBM (the reference hands a Vec<Vec<Event>> across an mpsc channel, simulproc.rs:235).  Optionally the raw .adder
stream is written as well (what `Framed::write_out` with EncoderType::Raw would have produced).
"""
from __future__ import annotations

from typing import BinaryIO, Optional

import numpy as np

from adder_codec_rs_b200 import binding as B
from adder_codec_rs_b200.framed import Framed, NoData, RawAdderWriter


class SimulProcessor:
    def __init__(self, source: Framed, ref_time: int, output: BinaryIO, frame_max: int = 0, codec_version: int = 3,
                 time_mode: Optional[int] = None, view_mode: int = B.VIEW_INTENSITY, raw_output: Optional[BinaryIO] = None,
                 ring_frames: int = 0):
        """SimulProcessor::new, simulproc.rs:113-225."""
        self.source = source
        self.video = source.get_video_ref()
        info = self.video.info()
        if time_mode is None:  # the framer must read the stream in the time mode the transcoder writes it in
            time_mode = info.time_mode
        assert time_mode == info.time_mode, "the framer's time mode must be the source video's"
        reconstructed_frame_rate = source.source_fps
        # For instantaneous reconstruction the frame rate must match the source rate (simulproc.rs:142-146)
        assert info.tps // ref_time == int(reconstructed_frame_rate), "tps / ref_time must equal the source frame rate"
        self.framer = B.Framer(info.width, info.height, info.channels, info.chunk_rows, codec_version, time_mode, info.tps,
                               ref_time, info.delta_t_max, output_fps=reconstructed_frame_rate, view_mode=view_mode,
                               source_camera=0, ring_frames=ring_frames, device=info.device)
        self.output = output
        self.frame_max = frame_max
        self.frames_out = 0
        self.raw = RawAdderWriter(raw_output, self.video, codec_version) if raw_output is not None else None
        self._P_out = info.width * info.height * info.channels
        self._d_frame = self.video.device_alloc(info.width * info.height * max(3, info.channels))  # a colour source may feed a gray transcode
        self._cap = self._P_out * (info.max_depth + 2)  # the library's worst case: 1 + max(L, 2) + 1 events per pixel per frame
        self._d_events = self.video.device_alloc(self._cap * 12)
        self._d_off = self.video.device_alloc((info.n_chunks + 1) * 4)
        self._d_raw = self.video.device_alloc(self._cap * 11) if self.raw else None
        self._n_chunks = info.n_chunks

    def run(self, frame_max: int = 0) -> int:
        """SimulProcessor::run, simulproc.rs:229-277.  Returns the number of input frames consumed."""
        consumed = 0
        while True:
            try:
                frame = self.source.next_frame()  # the decode half of consume(), framed.rs:128
            except NoData:
                break
            self._d_frame.from_host(frame)
            ref_time = self.video.info().ref_time
            self.video.integrate_frames_device(self._d_frame.ptr, 0, 1, float(ref_time), self._d_events.ptr, self._cap, self._d_off.ptr)
            if self.raw:
                self.video.raw_encode_device(self._d_events.ptr, self._d_off.ptr + self._n_chunks * 4, self._cap, self._d_raw.ptr)
            self.video.sync()
            consumed += 1
            if self.raw:
                n = int(self._d_off.to_host(np.uint32, nbytes=4, offset=self._n_chunks * 4)[0])
                self.raw.write_body(self._d_raw.to_host(nbytes=n * self.raw.event_size))
            if self.framer.ingest_events_device(self._d_events.ptr, self._d_off.ptr):  # simulproc.rs:178
                frames = self.framer.write_multi_frame_bytes()
                if len(frames) == 0:
                    raise RuntimeError("Should have frame, but didn't")  # simulproc.rs:180-183
                self.output.write(frames.tobytes())
                self.frames_out += len(frames)
            if self.frame_max > 0 and self.frames_out + 1 >= self.frame_max:  # frame_count starts at 1, simulproc.rs:173, :208
                break
            if frame_max > 0 and self.video.in_interval_count >= frame_max:  # :262-265
                break
        if self.raw:
            self.raw.close()  # end_write_stream, video.rs:641-648
        self.output.flush()
        return consumed
