#!/bin/bash
# Build kernel variants on the GPU box and bench each (device-resident only).
# Usage: gpu_variants.sh "<flags1>" "<flags2>" ...   (a flag set may start with SRC=<file> to swap the kernel source in;
# BASE=<lib.so> benches that prebuilt library first and last, to normalise between boxes)
PKG=multithreading_string_matching_b200
run() { timeout 300 python bench.py --no-cpu --no-e2e --steps 10 ${BENCH_ARGS} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GB/s=%.1f kernel_ms=%.3f matches=%d'%(d['value'],d['roofline']['kernel_ms'],d['matches_per_step']))" 2>&1 | tail -1; }
cp $PKG/csrc/cuda/union_kernel.cu /tmp/union_kernel.orig.cu
if [ -n "$BASE" ]; then cp $PKG/libkmpb200.so /tmp/tree.so; cp "$BASE" $PKG/libkmpb200.so; echo "BASE $BASE | $(run)"; fi
for flags in "$@"; do
  src=/tmp/union_kernel.orig.cu
  case "$flags" in SRC=*) src="${flags%% *}"; src="${src#SRC=}"; flags="${flags#SRC=* }"; [ "$flags" = "SRC=$src" ] && flags="";; esac
  cp "$src" $PKG/csrc/cuda/union_kernel.cu
  rm -f $PKG/build/union_kernel.o $PKG/build/tables.o $PKG/libkmpb200.so  # tables.cu uploads the filter words (KMPB_FILTER6)
  make -s -C $PKG EXTRA_NVFLAGS="$flags" >/dev/null 2>&1 || { echo "BUILD FAILED: $src $flags"; continue; }
  regs=$(grep -A2 "kmpb_union_kernel" $PKG/build/union_kernel.ptxas.log | grep -o "Used [0-9]* registers" | head -1)
  spill=$(grep -A1 "Function properties for _Z17kmpb_union_kernel" $PKG/build/union_kernel.ptxas.log | tail -1 | tr -s ' ')
  if [ -n "$VARIANT_TESTS" ]; then timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1; fi
  echo "$(basename $src) $flags | $regs |$spill | $(run) $([ -n "$EXTRA_PY" ] && python $EXTRA_PY 2>&1 | tail -1)"
done
cp /tmp/union_kernel.orig.cu $PKG/csrc/cuda/union_kernel.cu
if [ -n "$BASE" ]; then cp "$BASE" $PKG/libkmpb200.so; echo "BASE again | $(run)"; fi
