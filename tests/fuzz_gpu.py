#!/usr/bin/env python3
"""Randomised differential run of the CUDA path against the oracle (not collected by pytest: run it by
hand on a B200, `python tests/fuzz_gpu.py [seconds] [seed]`).  Small alphabets so that matches, overlaps,
NULs and packet boundaries are dense; packet lengths from 0 to a few rows; every few rounds a batch large
enough to span many work items."""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multithreading_string_matching_b200 as kmp  # noqa: E402
from oracle import oracle_py  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
m = kmp.Matcher(0, engine="union")
t0, rounds, total_bytes = time.time(), 0, 0
while time.time() - t0 < budget:
    alpha = bytes(rng.sample(range(1, 256), rng.choice((2, 3, 5, 26, 95))))
    p_nul = rng.choice((0.0, 0.0005, 0.01, 0.2, 0.7, 0.97))
    n_pat = rng.choice((1, 3, 17, 97, 300))
    max_len = rng.choice((2, 4, 7, 13, 40, 99))
    pats = [bytes(rng.choice(alpha) for _ in range(rng.randint(1, max_len))) for _ in range(n_pat)]
    big = rounds % 4 == 3
    n_pk = rng.randint(200, 1500) if big else rng.randint(1, 120)
    top = rng.choice((0, 5, 70, 1500, 4000))
    packets = []
    for _ in range(n_pk):
        ln = rng.randint(0, top) if rng.random() < 0.9 else rng.choice((0, 1, 31, 32, 33, 1023, 1024, 1025))
        b = bytearray(rng.choice(alpha) for _ in range(ln))
        for i in range(ln):
            if p_nul and rng.random() < p_nul:
                b[i] = 0
        if ln and rng.random() < 0.3:  # plant patterns, also across the packet's end
            for _ in range(rng.randint(1, 4)):
                p = rng.choice(pats)
                at = rng.randint(0, ln)
                b[at:at + len(p)] = p[: max(0, ln - at)]
        packets.append(bytes(b[:ln]))
    offsets = np.zeros(len(packets) + 1, dtype=np.uint64)
    np.cumsum([len(p) for p in packets], out=offsets[1:])
    data = np.frombuffer(b"".join(packets) + b"\0" * 64, dtype=np.uint8)[: int(offsets[-1])]
    m.set_patterns(pats)
    got = m.count_host(data, offsets)
    want = oracle_py.count_csr(data, offsets, pats)
    if got != want:
        bad = [(pats[i], got[i], want[i]) for i in range(len(pats)) if got[i] != want[i]][:5]
        print("MISMATCH in round", rounds, "alphabet", len(alpha), "p_nul", p_nul, "patterns", n_pat, "max_len", max_len,
              "packets", n_pk, "top", top, bad)
        sys.exit(1)
    rounds += 1
    total_bytes += int(offsets[-1])
print("fuzz ok: %d rounds, %.1f MB, %.0f s" % (rounds, total_bytes / 1e6, time.time() - t0))
