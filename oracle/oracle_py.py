"""ctypes view of oracle/liboracle.so (the CPU restatement, oracle/kmp_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs as the checker or the timed CPU baseline -- never by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PROTO_UDP, PROTO_TCP = 0, 1
_u8p = ctypes.POINTER(ctypes.c_uint8)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int)


def build():
    """Compile the restatement (and oracle/_ref when the reference checkout is present)."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        src = os.path.join(HERE, "kmp_oracle.c")
        if not os.path.isfile(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        L = ctypes.CDLL(path)
        L.orc_kmp_prefix.argtypes = [_u8p, ctypes.c_int, _i32p]
        L.orc_kmp_count.restype = ctypes.c_int64
        L.orc_kmp_count.argtypes = [_u8p, ctypes.c_int64, _u8p, ctypes.c_int, _i32p]
        L.orc_text_len.restype = ctypes.c_int64
        L.orc_text_len.argtypes = [_u8p, ctypes.c_int64]
        L.orc_udp_payload.argtypes = [_u8p, ctypes.c_uint32, _u32p, _u32p]
        L.orc_tcp_payload.argtypes = [_u8p, ctypes.c_uint32, _u32p, _u32p]
        L.orc_load_patterns.argtypes = [ctypes.c_char_p, ctypes.POINTER(_u8p), ctypes.POINTER(_u32p), _u32p]
        L.orc_load_pcap_csr.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(_u8p),
                                        ctypes.POINTER(_u64p), _u64p, _u64p]
        L.orc_count_csr.restype = None
        L.orc_count_csr.argtypes = [_u8p, _u64p, ctypes.c_uint64, _u8p, _u32p, ctypes.c_uint32, _i64p, ctypes.c_int]
        L.orc_format_report.restype = ctypes.c_void_p
        L.orc_format_report.argtypes = [_u8p, _u32p, ctypes.c_uint32, _i64p]
        L.orc_free.argtypes = [ctypes.c_void_p]
        _LIB = L
    return _LIB


def _ptr(a, ty):
    return a.ctypes.data_as(ty)


def _bytes_arr(b):
    a = np.frombuffer(bytes(b), dtype=np.uint8)
    return a if a.size else np.zeros(1, dtype=np.uint8)


def kmp_prefix(pattern):
    p = _bytes_arr(pattern)
    pi = np.zeros(max(len(pattern), 1), dtype=np.int32)
    lib().orc_kmp_prefix(_ptr(p, _u8p), len(pattern), _ptr(pi, _i32p))
    return pi[: len(pattern)].tolist()


def kmp_count(text, pattern):
    """Reference semantics for one (payload, pattern) pair, including the NUL rule."""
    t, p = _bytes_arr(text), _bytes_arr(pattern)
    pi = np.zeros(max(len(pattern), 1), dtype=np.int32)
    L = lib()
    L.orc_kmp_prefix(_ptr(p, _u8p), len(pattern), _ptr(pi, _i32p))
    n = L.orc_text_len(_ptr(t, _u8p), len(text))
    return int(L.orc_kmp_count(_ptr(t, _u8p), n, _ptr(p, _u8p), len(pattern), _ptr(pi, _i32p)))


def extract(frame, proto):
    f = _bytes_arr(frame)
    off, plen = ctypes.c_uint32(0), ctypes.c_uint32(0)
    fn = lib().orc_tcp_payload if proto in (PROTO_TCP, "tcp") else lib().orc_udp_payload
    ok = fn(_ptr(f, _u8p), len(frame), ctypes.byref(off), ctypes.byref(plen))
    return (off.value, plen.value) if ok else None


def load_patterns(path):
    """-> list of bytes tokens in file order (duplicates kept), per serial.c:54-87."""
    blob, off, n = _u8p(), _u32p(), ctypes.c_uint32(0)
    rc = lib().orc_load_patterns(os.fsencode(path), ctypes.byref(blob), ctypes.byref(off), ctypes.byref(n))
    if rc != 0:
        raise OSError("orc_load_patterns(%s) -> %d" % (path, rc))
    offs = [off[i] for i in range(n.value + 1)]
    data = ctypes.string_at(blob, offs[-1]) if offs[-1] else b""
    lib().orc_free(blob)
    lib().orc_free(off)
    return [data[offs[i]:offs[i + 1]] for i in range(n.value)]


def load_pcap_csr(path, proto=PROTO_UDP):
    """-> (bytes uint8[total], offsets uint64[n+1], n_frames), per serial.c:91-141."""
    proto = {"udp": PROTO_UDP, "tcp": PROTO_TCP}.get(proto, proto)
    b, o = _u8p(), _u64p()
    n, frames = ctypes.c_uint64(0), ctypes.c_uint64(0)
    rc = lib().orc_load_pcap_csr(os.fsencode(path), proto, ctypes.byref(b), ctypes.byref(o),
                                 ctypes.byref(n), ctypes.byref(frames))
    if rc != 0:
        raise OSError("orc_load_pcap_csr(%s) -> %d" % (path, rc))
    offsets = np.ctypeslib.as_array(o, shape=(n.value + 1,)).copy()
    total = int(offsets[-1])
    data = np.ctypeslib.as_array(b, shape=(max(total, 1),))[:total].copy()
    lib().orc_free(b)
    lib().orc_free(o)
    return data, offsets, frames.value


def pack_patterns(patterns):
    blob = np.frombuffer(b"".join(patterns), dtype=np.uint8).copy()
    if blob.size == 0:
        blob = np.zeros(1, dtype=np.uint8)
    off = np.zeros(len(patterns) + 1, dtype=np.uint32)
    np.cumsum([len(p) for p in patterns], out=off[1:])
    return blob, off


def count_csr(data, offsets, patterns, threads=0):
    """Per-pattern counts (list of int, file order) of the oracle on a CSR batch."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    if data.size == 0:
        data = np.zeros(1, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    blob, off = pack_patterns(patterns)
    counts = np.zeros(max(len(patterns), 1), dtype=np.int64)
    if threads <= 0:
        threads = os.cpu_count() or 1
    lib().orc_count_csr(_ptr(data, _u8p), _ptr(offsets, _u64p), len(offsets) - 1, _ptr(blob, _u8p),
                        _ptr(off, _u32p), len(patterns), _ptr(counts, _i64p), threads)
    return counts[: len(patterns)].tolist()


def format_report(patterns, counts):
    blob, off = pack_patterns(patterns)
    c = np.asarray(list(counts) + [0], dtype=np.int64)
    p = lib().orc_format_report(_ptr(blob, _u8p), _ptr(off, _u32p), len(patterns), _ptr(c, _i64p))
    s = ctypes.string_at(p)
    lib().orc_free(p)
    return s
