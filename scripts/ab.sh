#!/bin/bash
# A/B on one box: GPU parity suite + device-resident bench of the working tree's library, then the same bench with
# the libraries under ab/ (built from earlier commits).  Usage: scripts/ab.sh [ab/libX.so ...]
PKG=multithreading_string_matching_b200
mkdir -p gpurun_out
run() { timeout 300 python bench.py --no-cpu --no-e2e --steps 10 ${BENCH_ARGS} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GB/s=%.1f kernel_ms=%.3f matches=%d'%(d['value'],d['roofline']['kernel_ms'],d['matches_per_step']))"; }
if [ -z "$SKIP_TESTS" ]; then timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
echo "tree: $(run)"
cp $PKG/libkmpb200.so /tmp/tree.so
for lib in "$@"; do cp "$lib" $PKG/libkmpb200.so; echo "$lib: $(run)"; done
cp /tmp/tree.so $PKG/libkmpb200.so
echo "tree again: $(run)"
