#!/bin/bash
# Builds A/B variants of libpbf_b200.so under fluidsimulator_b200/lib/variants/<name>/ (tools/variants.sh
# runs tools/explore.py against each on the GPU box).  Usage: tools/build_variants.sh name "EXTRA flags" ...
set -e
cd "$(dirname "$0")/../fluidsimulator_b200/csrc"
while [ $# -ge 2 ]; do
  name=$1; extra=$2; shift 2
  make -s -j8 LIBDIR=../lib/variants/$name OBJDIR=../build/variants/$name EXTRA="$extra" ../lib/variants/$name/libpbf_b200.so >/dev/null
  echo "built $name ($extra): $(grep -h -A2 'k_lambdaILb1' ../build/variants/$name/kernels/solve.ptxas.log | grep -o 'Used [0-9]* registers' | head -1)"
done
