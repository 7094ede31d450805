#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG ..." unit1.cu unit2.cu ... : an A/B build of libcpimh (.ab/libcpimh_NAME.so) in which the listed
# translation units are recompiled with extra flags and everything else comes from the standard object files.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; FLAGS=$2; shift 2
mkdir -p $ROOT/.ab/obj_$NAME
cd $ROOT/unbiased_pimh_b200/csrc
for u in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr $FLAGS -c $u -o $ROOT/.ab/obj_$NAME/${u%.cu}.o &
done
wait
OBJS=""
for o in build/*.o; do b=$(basename $o); if [ -f $ROOT/.ab/obj_$NAME/$b ]; then OBJS="$OBJS $ROOT/.ab/obj_$NAME/$b"; else OBJS="$OBJS $o"; fi; done
nvcc -shared -o $ROOT/.ab/libcpimh_$NAME.so $OBJS -gencode arch=compute_100a,code=sm_100a
echo $ROOT/.ab/libcpimh_$NAME.so
