#!/bin/bash
# Developer tool (GPU box): time every variants/lib_*.so on the given probe configs.
cd "$(dirname "$0")/.."
for lib in variants/lib_*.so; do
  echo "== $lib"
  RTB200_LIB=$PWD/$lib timeout 120 python tools/perf_probe.py "$@" 2>&1 | grep -E "sah|lbvh" | cut -c1-330
done
