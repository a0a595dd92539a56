#!/bin/bash
# Build a variant of libsdc_b200.so with extra -D flags into sdc_b200/lib/variants/<name>.so (for A/B timing):
#   tools/build_variant.sh lb8 -DSDCB200_SORT_LB=8
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/sdc_b200/lib/variants
mkdir -p $out/obj_$name
for f in $root/sdc_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -Xcompiler -fPIC -I $root/include "$@" -c $f -o $out/obj_$name/$b.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o $out/$name.so $out/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
rm -rf $out/obj_$name
echo $out/$name.so
