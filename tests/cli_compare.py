#!/usr/bin/env python3
"""Checker-side tool (under tests/ because it executes the reference binaries of oracle/_ref).
Process-level comparison on one synthetic pcap (the reference's surface: <pcap> <strings.txt> ...):
bin/kmp_match vs the unmodified reference programs built under oracle/_ref.  Prints wall time of each
whole process, the time each program reports itself, and whether the count lines are identical."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import multithreading_string_matching_b200 as kmp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
strings = os.path.join(ROOT, "tests", "golden", "data", "strings.txt")
pats = kmp.load_patterns(strings)
synth = kmp.Synth(seed=0xB200, payload_len=1400, plants=2, plant_patterns=pats)
data, off = synth.fill_host(0, n)
tmp = tempfile.mkdtemp(prefix="kmpb_cli_")
pcap = os.path.join(tmp, "sample.pcap")
bench.write_pcap(pcap, data, off)
print("pcap: %d packets x 1400 B, %.1f MB" % (n, os.path.getsize(pcap) / 1e6))


def run(cmd, unlimited_stack=False):
    import resource

    def pre():
        if unlimited_stack:
            resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
    t0 = time.perf_counter()
    out = subprocess.run(cmd, capture_output=True, preexec_fn=pre).stdout.decode("latin-1").splitlines()
    wall = time.perf_counter() - t0
    return wall, out[-1] if out else "", "\n".join(out[:-1])


cores = os.cpu_count()
results = []
exe = os.path.join(ROOT, "multithreading_string_matching_b200", "bin", "kmp_match")
for label, cmd, stack in (("kmp_match (1 GPU), first run", [exe, pcap, strings], False),
                          ("kmp_match (1 GPU), second run", [exe, pcap, strings], False),
                          ("openmp_data -O2, %d threads" % cores, [os.path.join(ROOT, "oracle", "_ref", "openmp_data"), pcap, strings, str(cores)], True),
                          ("serial -O2", [os.path.join(ROOT, "oracle", "_ref", "serial"), pcap, strings], False)):
    if not os.path.isfile(cmd[0]):
        print("%-34s missing" % label)
        continue
    if label.startswith("serial") and n > 100_000:
        print("%-34s skipped above 100k packets (minutes of CPU time)" % label)
        continue
    wall, last, lines = run(cmd, stack)
    results.append(lines)
    print("%-34s wall %8.3f s   self-reported: %s" % (label, wall, last.strip()))
print("count lines identical:", all(r == results[0] for r in results))
