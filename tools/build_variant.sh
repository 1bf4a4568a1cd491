#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags]  -> astrild_b200/lib/variants/lib<name>.so (A/B kernel builds; tools, not product)
set -e
cd "$(dirname "$0")/../astrild_b200/csrc"
name=$1; shift
mkdir -p ../lib/variants /tmp/apk_var_$name
for f in api bin_power deposit_atomic deposit_sorted mesh_ops route; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off \
    -I../../include -I. --expt-relaxed-constexpr "$@" -c $f.cu -o /tmp/apk_var_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/lib$name.so /tmp/apk_var_$name/*.o -lcufft -lcudart
echo built ../lib/variants/lib$name.so
