#!/bin/bash
# usage: tools/build_variant.sh NAME [extra nvcc flags...] : builds the working tree's csrc in a scratch copy with
# extra flags and drops the library at calypso-gap_b200/lib/libgapcu_NAME.so (for A/B runs via GAPCU_LIB)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=/tmp/gapcu_variant_$name
rm -rf $tmp && mkdir -p $tmp
cp -r $root/calypso-gap_b200 $root/include $tmp/
rm -rf $tmp/calypso-gap_b200/build $tmp/calypso-gap_b200/lib
if [ $# -gt 0 ]; then
  flags=$(printf '"%s", ' "$@")
  sed -i "s|\"-O3\", \"-std=c++17\"|${flags}\"-O3\", \"-std=c++17\"|" $tmp/calypso-gap_b200/build.py
fi
(cd $tmp/calypso-gap_b200 && python -c "import build; build.build_libgapcu()" > $tmp/build.log 2>&1) || { tail -20 $tmp/build.log; exit 1; }
cp $tmp/calypso-gap_b200/lib/libgapcu.so $root/calypso-gap_b200/lib/libgapcu_$name.so
echo built $root/calypso-gap_b200/lib/libgapcu_$name.so
