#!/bin/bash
# Developer tool (GPU box): tools/ab_probe.py for the default library and every variants/lib_*.so.  usage: tools/run_ab.sh [c3 c2 c4 ...]
cd "$(dirname "$0")/.."
timeout 300 python tools/ab_probe.py "$@" 2>&1 | tail -1
for lib in variants/lib_*.so; do
  RTB200_LIB=$PWD/$lib timeout 300 python tools/ab_probe.py "$@" 2>&1 | tail -1
done
