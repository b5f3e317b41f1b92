#!/bin/bash
# Builds tuning variants of libmagi_attn.so into magi-1_b200/variants/ (used by tools/bench_variants.py on the GPU box).
# usage: tools/build_variants.sh name1="-DFLAG=.." name2="..."
set -e
cd "$(dirname "$0")/../magi-1_b200/csrc"
mkdir -p ../variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  make -s BUILD="build_$name" OUT="../variants/lib_$name.so" EXTRA="$flags" -j4
  echo "built variants/lib_$name.so  [$flags]"
done
