#!/bin/bash
# Build kernel variants HERE (nvcc cross-compiles) into ab/libkmpb200_<name>.so; they travel to the GPU box with the
# snapshot and scripts/ab.sh benches them.  Usage: scripts/build_variants.sh name1:"-DFLAG=1 ..." name2:"..."
# (a variant rebuilds union_kernel.cu and tables.cu only; the working tree's library is restored at the end)
PKG=multithreading_string_matching_b200
mkdir -p ab
make -s -C $PKG >/dev/null 2>&1 || { echo "tree build failed"; exit 1; }
cp $PKG/libkmpb200.so /tmp/tree.so; cp $PKG/build/union_kernel.o /tmp/tree_union.o; cp $PKG/build/tables.o /tmp/tree_tables.o
for v in "$@"; do
  name="${v%%:*}"; flags="${v#*:}"
  rm -f $PKG/build/union_kernel.o $PKG/build/tables.o $PKG/libkmpb200.so
  if make -s -C $PKG EXTRA_NVFLAGS="$flags" >/dev/null 2>&1; then
    cp $PKG/libkmpb200.so ab/libkmpb200_$name.so
    echo "$name [$flags]: $(grep -A2 'kmpb_union_kernel' $PKG/build/union_kernel.ptxas.log | grep -o 'Used [0-9]* registers' | head -1), $(grep -A1 'Function properties for _Z17kmpb_union_kernel' $PKG/build/union_kernel.ptxas.log | tail -1 | tr -s ' ')"
  else echo "$name: BUILD FAILED"; tail -5 $PKG/build/union_kernel.ptxas.log; fi
done
cp /tmp/tree_union.o $PKG/build/union_kernel.o; cp /tmp/tree_tables.o $PKG/build/tables.o; cp /tmp/tree.so $PKG/libkmpb200.so
touch $PKG/build/*.o $PKG/libkmpb200.so $PKG/bin/kmp_match
