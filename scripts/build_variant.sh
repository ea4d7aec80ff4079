#!/bin/bash
# Builds an alternative libcpk_<tag>.so from the current sources (or from a git revision) with extra
# nvcc flags, in a scratch directory so that the objects of the default build stay untouched.
#   scripts/build_variant.sh <tag> "<extra nvcc flags>" [git-rev]
# A/B runs on ONE box then pick the build through CPK_LIB_PATH=$PWD/cpkrylov_b200/libcpk_<tag>.so
set -e
tag=$1; extra=$2; rev=$3
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d /tmp/cpk_build_XXXX)
mkdir -p $tmp/cpkrylov_b200 $tmp/include
if [ -n "$rev" ]; then
  (cd $root && git archive $rev cpkrylov_b200/csrc include) | tar -x -C $tmp
else
  cp -r $root/cpkrylov_b200/csrc $tmp/cpkrylov_b200/ && cp $root/include/*.h $tmp/include/
  rm -f $tmp/cpkrylov_b200/csrc/*.o
fi
make -C $tmp/cpkrylov_b200/csrc -j8 EXTRA="$extra" OUT=$root/cpkrylov_b200/libcpk_$tag.so > $tmp/log 2>&1 || { tail -20 $tmp/log; exit 1; }
rm -rf $tmp
echo "built cpkrylov_b200/libcpk_$tag.so ($extra)"
