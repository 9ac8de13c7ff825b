#!/usr/bin/env python3
"""Regenerate tests/golden/ from the UNMODIFIED reference compiled into oracle/_ref/.

Run in the build container (where /root/reference exists), after `make -C oracle`:

    python tests/golden/make_golden.py

It writes
  data/                 the reference's bundled DATA fixtures (pcaps + strings.txt; no source code),
                        copied so the GPU box -- which has no /root/reference -- can run config[1]
  expected/<run>.txt    stdout of oracle/_ref/serial (minus its Elapsed line) for every bundled pcap
                        in udp and tcp mode, after checking that serial (-O2), serial_doc (documented
                        flags, serial.c:2), openmp_data and openmp_data_doc at 1/3/8 threads all agree
  kmp_vectors.json      the reference's own kmp_prefix / kmp_matcher (serial.c:190-238, linked from
                        oracle/_ref/libserial_kmp.so) on seeded random patterns and NUL-terminated texts
  extract_vectors.json  the reference's own dump_UDP_packet / dump_TCP_packet (packet_dumping.h:87-188)
                        on seeded random frames (TCP only inside its defined domain: frame long enough
                        for the headers it announces)

Nothing here is imported by the product.
"""
import ctypes
import json
import os
import random
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("KMP_REFERENCE_DIR", "/root/reference")
REFBIN = os.path.join(ROOT, "oracle", "_ref")
PCAPS = ["udp.pcap", "udp_1000.pcap", "big_udp.pcap", "very_big_udp.pcap", "tcp.pcap"]
# Runs left out because the reference's answer is not a function of its input there: the TCP parser
# (packet_dumping.h:150-188, no protocol or bounds check) applied to very_big_udp.pcap's UDP frames
# reads payloads whose strlen() runs off the heap block; openmp_data_doc was seen to print different
# lines from one process to the next.
UB_RUNS = {("very_big_udp.pcap", "tcp")}


def run(binary, *args):
    env = {k: v for k, v in os.environ.items() if not k.startswith("MALLOC_")}
    out = subprocess.run([os.path.join(REFBIN, binary), *args], capture_output=True, env=env, check=True).stdout
    lines = out.decode("latin-1").splitlines(keepends=True)
    assert lines[-1].startswith("Elapsed time = "), lines[-1]
    return "".join(lines[:-1])


def golden_outputs():
    os.makedirs(os.path.join(HERE, "expected"), exist_ok=True)
    strings = os.path.join(REF, "strings.txt")
    for pcap in PCAPS:
        path = os.path.join(REF, pcap)
        for proto in ("udp", "tcp"):
            if (pcap, proto) in UB_RUNS:
                continue
            base = run("serial", path, strings, proto)
            assert run("serial_doc", path, strings, proto) == base, (pcap, proto, "serial_doc")
            for variant in ("openmp_data", "openmp_data_doc"):
                for threads in ("1", "3", "8"):
                    got = run(variant, path, strings, threads, proto)
                    assert got == base, (pcap, proto, variant, threads)
            if proto == "udp":
                assert run("serial", path, strings) == base  # default protocol is udp (serial.c:31)
            name = "%s.%s.txt" % (pcap[:-5], proto)
            with open(os.path.join(HERE, "expected", name), "w", encoding="latin-1") as f:
                f.write(base)
            print("expected/%s: %d pattern lines" % (name, base.count("\n") - 1))


def kmp_vectors():
    lib = ctypes.CDLL(os.path.join(REFBIN, "libserial_kmp.so"))
    lib.kmp_prefix.restype = ctypes.POINTER(ctypes.c_int)
    lib.kmp_prefix.argtypes = [ctypes.c_char_p]
    lib.kmp_matcher.restype = ctypes.c_int
    lib.kmp_matcher.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]
    rng = random.Random(0xB200)
    alphabets = [b"ab", b"abc", b"aA-_", bytes(range(0x20, 0x7F)), bytes(range(1, 256))]
    fixed = [b"aabaaab", b"abcabc", b"aaaa", b"content-list", b"icmp_seq", b"a", b"aa", b"id", b"rr"]
    vectors = []
    for i in range(600):
        alpha = alphabets[i % len(alphabets)]
        if i < len(fixed):
            pat = fixed[i]
        else:
            pat = bytes(rng.choice(alpha) for _ in range(rng.choice([1, 1, 2, 2, 3, 4, 5, 8, 12, 33, 99])))
        n = rng.choice([0, 1, 2, 3, 7, 16, 64, 257, 1400])
        text = bytearray(rng.choice(alpha) for _ in range(n))
        if n and i % 3 == 0:  # plant the pattern a few times, overlapping allowed
            for _ in range(rng.randint(1, 4)):
                at = rng.randrange(0, n)
                text[at:at + len(pat)] = pat[: max(0, n - at)]
        text = bytes(text[:n])
        pi_ptr = lib.kmp_prefix(pat)
        pi = [pi_ptr[k] for k in range(len(pat))]
        count = lib.kmp_matcher(text, pat, pi_ptr)  # c_char_p adds the terminating NUL
        libc.free(pi_ptr)
        vectors.append({"pattern": pat.hex(), "text": text.hex(), "pi": pi, "count": count})
    with open(os.path.join(HERE, "kmp_vectors.json"), "w") as f:
        json.dump({"source": "oracle/_ref/libserial_kmp.so (serial.c:190-238)", "vectors": vectors}, f)
    print("kmp_vectors.json: %d vectors, %d with hits" % (len(vectors), sum(v["count"] > 0 for v in vectors)))


def extract_vectors():
    lib = ctypes.CDLL(os.path.join(REFBIN, "libserial_kmp.so"))
    for fn in (lib.dump_UDP_packet, lib.dump_TCP_packet):
        fn.restype = ctypes.c_void_p
        fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint), ctypes.c_uint]
    rng = random.Random(0x0DF0)
    out = []
    for i in range(400):
        proto = "tcp" if i % 2 else "udp"
        length = rng.choice([0, 5, 13, 14, 20, 33, 34, 41, 42, 43, 60, 90, 200, 1442])
        frame = bytearray(rng.randrange(256) for _ in range(length))
        if length > 23 and i % 4 < 3:  # mostly plausible IPv4 headers, with hostile IHL / offsets mixed in
            frame[12:14] = b"\x08\x00"
            frame[14] = 0x40 | rng.choice([5, 5, 5, 6, 15, 0, 4, 2])
            frame[23] = rng.choice([17, 17, 6, 6, 1])
            tcp_at = 14 + (frame[14] & 15) * 4
            if tcp_at + 12 < length:
                frame[tcp_at + 12] = rng.choice([5, 5, 8, 15, 4, 0]) << 4
        if proto == "tcp":
            # defined domain only: the reference reads and wraps outside shorter frames
            if length < 15:
                continue
            size_ip = (frame[14] & 15) * 4
            if size_ip >= 20:
                if length < 14 + size_ip + 13:
                    continue
                size_tcp = (frame[14 + size_ip + 12] >> 4) * 4
                if size_tcp >= 20 and length < 14 + size_ip + size_tcp:
                    continue
        buf = ctypes.create_string_buffer(bytes(frame) + b"\0" * 256, length + 256)
        plen = ctypes.c_uint(0xFFFFFFFF)
        fn = lib.dump_TCP_packet if proto == "tcp" else lib.dump_UDP_packet
        ptr = fn(ctypes.addressof(buf), ctypes.byref(plen), length)
        rec = {"proto": proto, "frame": bytes(frame).hex(), "ok": ptr is not None}
        if ptr is not None:
            rec["off"] = ptr - ctypes.addressof(buf)
            rec["len"] = plen.value
        out.append(rec)
    with open(os.path.join(HERE, "extract_vectors.json"), "w") as f:
        json.dump({"source": "oracle/_ref/libserial_kmp.so (packet_dumping.h:87-188)", "vectors": out}, f)
    print("extract_vectors.json: %d frames, %d accepted" % (len(out), sum(v["ok"] for v in out)))


def copy_data():
    os.makedirs(os.path.join(HERE, "data"), exist_ok=True)
    for name in PCAPS + ["strings.txt"]:
        shutil.copyfile(os.path.join(REF, name), os.path.join(HERE, "data", name))
    print("data/: %d files" % (len(PCAPS) + 1))


if __name__ == "__main__":
    if not os.path.isdir(REF) or not os.path.isfile(os.path.join(REFBIN, "serial")):
        sys.exit("need %s and oracle/_ref (make -C oracle)" % REF)
    copy_data()
    golden_outputs()
    kmp_vectors()
    extract_vectors()
