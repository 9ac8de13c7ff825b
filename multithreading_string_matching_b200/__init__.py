"""B200-native KMP packet matcher: drop-in for the hot path of Lemnon95/multithreading_string_matching.

Everything that computes lives in libkmpb200.so (C host code + sm_100a CUDA kernels, C ABI in
include/kmpb200.h); this package is the thin ctypes layer tests and bench.py drive it through."""
from ._lib import KmpbError, LIB_PATH, SIGNATURES, lib  # noqa: F401
from .matcher import (ENGINE_AUTO, ENGINE_PERPAT, ENGINE_UNION, PROTO_TCP, PROTO_UDP, Matcher, PayloadBatch,  # noqa: F401
                      Synth, device_count, extract_payload, format_report, load_patterns, pack_patterns, shard_range)

__version__ = "0.1.0"
