#!/bin/bash
# usage: tools/ab.sh "<workloads>" <variant> <variant> ...   ("main" = the in-tree library); tools, not product
wls=$1; shift
for v in "$@"; do
  if [ "$v" != main ]; then export ASTRILD_PK_LIB=$PWD/build/variants/libapk_$v.so; else unset ASTRILD_PK_LIB; fi
  echo "== variant $v"
  for w in $wls; do bash tools/bench_summary.sh $w 3 --no-e2e; done
done
