#!/bin/bash
# Build kernel variants on the GPU box and bench each (device-resident only). Usage: gpu_variants.sh "<flags1>" "<flags2>" ...
PKG=multithreading_string_matching_b200
for flags in "$@"; do
  rm -f $PKG/build/union_kernel.o $PKG/libkmpb200.so
  make -s -C $PKG EXTRA_NVFLAGS="$flags" >/dev/null 2>&1 || { echo "BUILD FAILED: $flags"; continue; }
  regs=$(grep -A2 "kmpb_union_kernel" $PKG/build/union_kernel.ptxas.log | grep -o "Used [0-9]* registers" | head -1)
  spill=$(grep -A1 "Function properties for _Z17kmpb_union_kernel" $PKG/build/union_kernel.ptxas.log | tail -1 | tr -s ' ')
  out=$(timeout 300 python bench.py --no-cpu --no-e2e --steps 10 ${BENCH_ARGS} 2>&1 | tail -1)
  echo "$flags | $regs |$spill | $(echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('GB/s=%.1f kernel_ms=%.3f matches=%d'%(d['value'],d['roofline']['kernel_ms'],d['matches_per_step']))" 2>&1 | tail -1)"
done
