#!/bin/bash
# Developer tool: build kernel variants (compile-time knobs) as separate libraries under variants/ for A/B timing.
# usage: tools/variants.sh name1 "FLAGS1" name2 "FLAGS2" ...
set -e
cd "$(dirname "$0")/../raytracer-group27_b200"
mkdir -p ../variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  make -s -j8 EXTRA="$flags" OUT=../variants/lib_$name.so BUILD=build_$name > /dev/null
  echo "built variants/lib_$name.so  ($flags)"
done
