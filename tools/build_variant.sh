#!/bin/bash
# build_variant.sh <name> <file.cu-to-substitute> : links a variant library build/variants/<name>.so that
# differs from the in-tree build only in one translation unit (kernel A/B experiments on one box)
set -e
NAME=$1; SRC=$2; BASE=$(basename $SRC .cu)
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC,-fvisibility=hidden \
  -I include -I uda_poseestimation_b200/csrc -c $SRC -o build/variants/${NAME}_${BASE}.o
OBJS=""
for o in build/udape_obj/*.o; do b=$(basename $o .o); if [ "$b" != "$BASE" ]; then OBJS="$OBJS $o"; fi; done
nvcc -shared -o build/variants/${NAME}.so $OBJS build/variants/${NAME}_${BASE}.o -cudart static -gencode arch=compute_100a,code=sm_100a
echo build/variants/${NAME}.so
