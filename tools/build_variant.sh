#!/bin/bash
# usage: tools/build_variant.sh <name> <file.cu> "<-D flags>"   -> build/variants/lib_<name>.so (other objects reused)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
C=$ROOT/ctr-design-and-path-plan_b200/csrc
mkdir -p $ROOT/build/variants $ROOT/build/obj
name=$1; src=$2; flags=$3
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$ROOT/include -I$C --expt-relaxed-constexpr $flags -c $C/$src -o $ROOT/build/obj/${name}.o
objs=""
for f in ctrb_api ctrb_model ctrb_collide ctrb_static ctrb_static_fast ctrb_knn ctrb_bvh; do
  if [ "$f.cu" == "$src" ]; then objs="$objs $ROOT/build/obj/${name}.o"; else objs="$objs $C/$f.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/build/variants/lib_${name}.so $objs -lcudart_static -lpthread -ldl -lrt
echo built lib_${name}.so
