#!/bin/bash
# Run the CUDA kernel tests under compute-sanitizer (memcheck / racecheck / synccheck / initcheck).
# The reference wires no sanitizer into its CI; this script is the equivalent of a TSAN/ASAN job
# for our hand-written kernels.  Usage: scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [pytest -k expr]
set -euo pipefail
tool=${1:-memcheck}
expr=${2:-"rms or rope or gemm_tc or fmha or sample or moe"}
cd "$(dirname "$0")/.."
export VLLM_B200_PDL=0          # programmatic dependent launch hides launch boundaries from racecheck
exec compute-sanitizer --tool "$tool" --error-exitcode 1 --launch-timeout 120 \
  python -m pytest tests/test_kernels_gpu.py -x -q -k "$expr"
