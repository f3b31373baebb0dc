"""adder_codec_rs_b200 — B200-native framed→ADΔER per-pixel transcode path.

The product is `libadder_b200.so` (hand-written sm_100a CUDA behind the C ABI in include/adder_b200.h).
This package is the thin Python host side used by the tests and the bench: a ctypes binding and a
`Video` class that mirrors the transcode-state surface of the reference's `Video<W>`
(adder-codec-rs/src/transcoder/source/video.rs:322-345) name for name.  There is no CPU fallback:
importing works anywhere, but creating a `Video` without a CUDA device raises.
"""
from .binding import (  # noqa: F401
    EVENT_DTYPE,
    Exchange,
    Framer,
    AdderError,
    MODE_CONTINUOUS,
    MODE_FRAME_PERFECT,
    MULTI_COLLAPSE,
    MULTI_NORMAL,
    TIME_ABSOLUTE_T,
    TIME_DELTA_T,
    TIME_MIXED,
    VIEW_D,
    VIEW_DELTA_T,
    VIEW_INTENSITY,
    VIEW_SAE,
    Video,
    build,
    compact_frame_bytes,
    expand_compact,
    crf_parameters,
    device_count,
    lib,
    pinned_empty,
    raw_eof,
)
from .framed import Framed, RawAdderWriter  # noqa: F401
