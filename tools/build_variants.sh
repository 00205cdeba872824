#!/usr/bin/env bash
# A/B builds of libgbcodec.so with -D switches for one source file:  tools/build_variants.sh step_tile.cu name1 "-DX=0" name2 "-DX=1 -DY=0" ...
# -> tools/variants/<name>/libgbcodec.so (objects of the other sources come from the package's _build/)
set -eu
src=$1; shift
pkg=infantposeestimation_gaussianbias_b200
base=${src%.cu}
while [[ $# -gt 1 ]]; do
  name=$1; flags=$2; shift 2
  mkdir -p tools/variants/$name
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags -c $pkg/csrc/$src -o tools/variants/$name/$base.o
  objs=$(ls $pkg/_build/*.o | grep -v "/$base.o")
  nvcc -shared -o tools/variants/$name/libgbcodec.so $objs tools/variants/$name/$base.o
  rm tools/variants/$name/$base.o
  echo built $name: $flags
done
