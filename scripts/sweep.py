#!/usr/bin/env python3
"""BASELINE.json configs 4 and 5 on one GPU (kernel-only GB/s by CUDA events; parity for these shapes is
in tests/test_gpu_parity.py).  Prints one JSON line per case.

  config 4  pattern sweep: n in {1..256} patterns of length {4..64}, random [a-z0-9], over 1 GB of
            synthetic payload (714 286 x 1400 B), union engine (perpat for the small sets)
  config 5  mixed 64 / 576 / 1400 / 9000-byte payloads: device-resident and through kmpb_count_host
"""
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multithreading_string_matching_b200 as kmp  # noqa: E402

ALPHA = b"abcdefghijklmnopqrstuvwxyz0123456789"


def time_device(m, d_bytes, d_off, n, nbytes, n_pat, reps=5):
    d_counts = torch.zeros(max(n_pat, 1), dtype=torch.int64, device="cuda:0")
    stream = torch.cuda.current_stream()
    m.set_profile(True)
    ms = []
    for i in range(reps + 2):
        d_counts.zero_()
        m.count_device(d_bytes.data_ptr(), d_off.data_ptr(), n, d_counts.data_ptr(), span=(0, nbytes), stream=stream.cuda_stream)
        if i >= 2:
            ms.append(m.last_kernel_ms())
    m.set_profile(False)
    torch.cuda.synchronize()
    return float(np.mean(ms)), int(d_counts.sum().item())


def main():
    strings = kmp.load_patterns(os.path.join(ROOT, "tests", "golden", "data", "strings.txt"))
    out = []
    # ---- config 4 ----------------------------------------------------------------------------
    n = 714_286
    m = kmp.Matcher(0, engine="union")
    m.set_patterns(strings)
    synth = kmp.Synth(seed=0xB200, payload_len=1400, plants=0, plant_patterns=())
    nbytes = synth.nbytes(0, n)
    d_bytes = torch.zeros(nbytes + 4096, dtype=torch.uint8, device="cuda:0")
    d_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda:0")
    synth.fill_device(m, 0, n, d_bytes.data_ptr(), d_off.data_ptr())
    torch.cuda.synchronize()
    rng = random.Random(4)
    for length in (4, 8, 16, 32, 64):
        for n_pat in (1, 2, 4, 8, 16, 32, 64, 128, 256):
            pats = set()
            while len(pats) < n_pat:
                pats.add(bytes(rng.choice(ALPHA) for _ in range(length)))
            pats = sorted(pats)
            for engine in (["union", "perpat"] if n_pat <= 4 else ["union"]):
                m.set_engine(engine)
                m.set_patterns(pats)
                ms, hits = time_device(m, d_bytes, d_off, n, nbytes, n_pat, reps=3 if engine == "perpat" else 5)
                line = {"config": 4, "engine": engine, "patterns": n_pat, "length": length, "payload_GB": nbytes / 1e9,
                        "kernel_ms": ms, "GBps": nbytes / ms / 1e6, "matches": hits}
                print(json.dumps(line), flush=True)
                out.append(line)
    m.set_engine("union")
    del d_bytes, d_off
    # ---- config 5 ----------------------------------------------------------------------------
    m.set_patterns(strings)
    synth = kmp.Synth(seed=11, len_mode=1, plants=2, plant_patterns=strings)
    n = 1_500_000
    nbytes = synth.nbytes(0, n)
    d_bytes = torch.zeros(nbytes + 4096, dtype=torch.uint8, device="cuda:0")
    d_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda:0")
    synth.fill_device(m, 0, n, d_bytes.data_ptr(), d_off.data_ptr())
    torch.cuda.synchronize()
    ms, hits = time_device(m, d_bytes, d_off, n, nbytes, len(strings))
    line = {"config": 5, "engine": "union", "packets": n, "payload_GB": nbytes / 1e9, "kernel_ms": ms,
            "GBps": nbytes / ms / 1e6, "matches": hits, "mode": "device-resident"}
    print(json.dumps(line), flush=True)
    h_bytes = torch.empty(nbytes + 4096, dtype=torch.uint8, pin_memory=True)
    h_off = torch.empty(n + 1, dtype=torch.int64, pin_memory=True)
    h_bytes[:nbytes].copy_(d_bytes[:nbytes])
    h_off.copy_(d_off)
    torch.cuda.synchronize()
    m.count_host_ptr(h_bytes.data_ptr(), h_off.data_ptr(), n)
    t0 = time.perf_counter()
    c = m.count_host_ptr(h_bytes.data_ptr(), h_off.data_ptr(), n)
    dt = time.perf_counter() - t0
    line = {"config": 5, "engine": "union", "packets": n, "payload_GB": nbytes / 1e9, "wall_ms": dt * 1e3,
            "GBps": nbytes / dt / 1e9, "matches": int(sum(c)), "mode": "end to end from pinned host memory (H2D inside)"}
    print(json.dumps(line), flush=True)
    m.close()


if __name__ == "__main__":
    main()
