#!/bin/bash
# build_variant.sh <name> <file.cu> [-Dmacro=..]...   ->  _ab/<name>.so (A/B builds of one translation unit)
set -e
name=$1; src=$2; shift 2
cd "$(dirname "$0")/../sln_amodal_b200/csrc"
mkdir -p ../../_ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -diag-suppress 177 "$@" -c $src -o ../../_ab/$name.o
objs=""
for f in lib crop nms proposal semdist detection rle unmold rpn; do
  if [ "$f.cu" == "$src" ]; then objs="$objs ../../_ab/$name.o"; else objs="$objs _obj/$f.o"; fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../_ab/$name.so $objs
rm -f ../../_ab/$name.o
echo built _ab/$name.so
