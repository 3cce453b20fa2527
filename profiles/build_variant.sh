#!/bin/bash
# build an experiment variant of the library: profiles/build_variant.sh NAME -DFLAG...  ->  visual_foresight_b200/_variants/libvfengine_NAME.so
# (select it with VF_ENGINE_LIB=...; the product library is built by __graft_entry__.build())
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
OUT=visual_foresight_b200/_variants; mkdir -p $OUT/obj_$NAME
for f in conv_simt pointwise cdna_cost cem conv_mma engine; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" \
       -c visual_foresight_b200/csrc/$f.cu -o $OUT/obj_$NAME/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libvfengine_$NAME.so $OUT/obj_$NAME/*.o
rm -rf $OUT/obj_$NAME
echo built $OUT/libvfengine_$NAME.so
