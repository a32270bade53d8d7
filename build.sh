#!/bin/bash
# Build libsfk.so in-tree for sm_100a. Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")"
SRC=semisupervisedobjectdetection_b200/csrc
OUT=semisupervisedobjectdetection_b200/libsfk.so
OBJDIR=build/obj
mkdir -p $OBJDIR
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
pids=()
for f in $SRC/*.cu; do
  o=$OBJDIR/$(basename ${f%.cu}).o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find $SRC include -newer "$o" \( -name '*.h' -o -name '*.cuh' \) | head -1)" ]; then
    nvcc $FLAGS ${SFK_NVCC_EXTRA} -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $OBJDIR/*.o
echo "built $OUT"
