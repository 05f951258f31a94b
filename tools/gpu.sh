#!/bin/bash
# Build the CUDA library (and check it exports every symbol the binding declares), then run a command on the GPU box.
# usage: tools/gpu.sh <timeout_s> '<command>' [logfile]
set -e
cd "$(dirname "$0")/.."
python tensor-train-interior-point-method_b200/build.py > /dev/null
python - <<'PY'
import sys
sys.path.insert(0, "tensor-train-interior-point-method_b200")
from ttipm_b200 import _cabi
_cabi.load()
PY
mkdir -p gpurun_out
gpurun --timeout "$1" -- "$2" > "${3:-gpurun_out/last.log}" 2>&1
