#!/bin/bash
# Build the CUDA library, check every declared export is present, then hand the command to gpurun.
#   tools/launch_gpu.sh [--gpus N] [--timeout S] -- <command>
set -e
cd "$(dirname "$0")/.."
bash mmhand_b200/csrc/build.sh 2>&1 | grep -E "error|built"
python -c "
from mmhand_b200 import lib as L
lib = L.load(); missing = [s for s in L.EXPORTS if not hasattr(lib, s)]; assert not missing, missing"
exec /usr/local/graft/bin/gpurun "$@"
