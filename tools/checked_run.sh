#!/bin/bash
# The GPU suite against the default library, then against the bounds-checked build ("The checked build").
# Under gpurun:  bash tools/checked_run.sh [tag]   -> gpurun_out/<tag>_*.log / .json
cd "$(dirname "$0")/.." || exit 1
TAG=${1:-r02h}
mkdir -p gpurun_out
CHK=$PWD/raytracer-group27_b200/librtb200_checked.so
t0=$(date +%s)
timeout -k 5 ${T1:-235} python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/${TAG}_default_suite.log 2>&1
echo "default suite rc=$? after $(( $(date +%s) - t0 )) s"; tail -n 3 gpurun_out/${TAG}_default_suite.log
t1=$(date +%s)
RTB200_LIB=$CHK RTB200_VIOLATIONS_OUT=gpurun_out/${TAG}_checked_suite.json timeout -k 5 ${T2:-170} python -m pytest tests -m gpu -q --durations=12 \
  --deselect "tests/test_gpu_parity.py::test_c5_lattice_full_mesh_subset_equals_exhaustive" > gpurun_out/${TAG}_checked_suite.log 2>&1
echo "checked suite rc=$? after $(( $(date +%s) - t1 )) s"; tail -n 3 gpurun_out/${TAG}_checked_suite.log; cat gpurun_out/${TAG}_checked_suite.json 2>/dev/null; echo
t2=$(date +%s)
RTB200_LIB=$CHK RTB200_VIOLATIONS_OUT=gpurun_out/${TAG}_checked_c5.json timeout -k 5 ${T3:-45} python -m pytest -q \
  "tests/test_gpu_parity.py::test_c5_lattice_full_mesh_subset_equals_exhaustive" > gpurun_out/${TAG}_checked_c5.log 2>&1
echo "checked C5 rc=$? after $(( $(date +%s) - t2 )) s"; tail -n 2 gpurun_out/${TAG}_checked_c5.log; cat gpurun_out/${TAG}_checked_c5.json 2>/dev/null; echo
exit 0
