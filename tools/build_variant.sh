#!/bin/bash
# A/B builds: bash tools/build_variant.sh <name> <extra nvcc flags...>  ->  hyslam_b200/lib/libhyorb_<name>.so  (use with HYORB_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../hyslam_b200/csrc"
mkdir -p ../build_$name ../lib
for f in tables tma preprocess pyramid level fast quadtree blur describe match stereo api; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off "$@" -c $f.cu -o ../build_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libhyorb_$name.so ../build_$name/*.o -cudart static
echo built ../lib/libhyorb_$name.so
