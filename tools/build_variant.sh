#!/bin/bash
# Build an A/B variant of the library: tools/build_variant.sh <name> <file.cu> <extra nvcc flags...>
# Recompiles ONE source with the extra flags, links it with the regular objects -> ab/lib_<name>.so
# (select it at run time with RP_LIB_PATH=ab/lib_<name>.so).
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
python -m repurpose_b200.build >/dev/null
mkdir -p ab/$name
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
  --expt-relaxed-constexpr -Xptxas -v "$@" -c repurpose_b200/csrc/$src -o ab/$name/$src.o 2> ab/$name/ptxas.log
objs=""
for f in repurpose_b200/_build/*.cu.o; do
  b=$(basename $f)
  if [ "$b" == "$src.o" ]; then objs="$objs ab/$name/$src.o"; else objs="$objs $f"; fi
done
/usr/local/cuda/bin/nvcc -shared -o ab/lib_$name.so $objs -cudart static -gencode arch=compute_100a,code=sm_100a
echo ab/lib_$name.so
