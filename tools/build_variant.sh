#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags]  -> build/variants/libapk_<name>.so (A/B kernel builds; tools, not product;
# run with ASTRILD_PK_LIB=$PWD/build/variants/libapk_<name>.so)
set -e
cd "$(dirname "$0")/../astrild_b200/csrc"
name=$1; shift
mkdir -p ../../build/variants /tmp/apk_var_$name
for f in api bin_kmu bin_power deposit_atomic deposit_sorted ingest mesh_ops power route; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off \
    -I../../include -I. --expt-relaxed-constexpr "$@" -c $f.cu -o /tmp/apk_var_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/variants/libapk_$name.so /tmp/apk_var_$name/*.o -lcufft -lcudart
echo built build/variants/libapk_$name.so
