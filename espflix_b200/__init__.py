"""espflix_b200 — B200-native drop-in for the espflix hot path: batched MPEG-1 decode of
352x192 I+P streams into the reference's striped YUV frame store, and NTSC/PAL composite field
synthesis. The product is libespflix_b200.so (C-ABI in include/espflix_b200.h); this package is
the thin ctypes host layer the tests and bench.py use."""
from .capi import (Context, EspflixError, audio_decode, audio_demux_ts, build_video_idx, lib_path, load_library,  # noqa: F401
                   tsidx_samples, tsidx_scan)

__all__ = ["Context", "EspflixError", "lib_path", "load_library", "tsidx_scan", "tsidx_samples", "build_video_idx", "audio_demux_ts", "audio_decode"]
