#!/bin/bash
# build in-tree, then run a command on a B200 box.  usage: tools/gpu.sh [--timeout S] [--gpus N] -- '<cmd>'
set -e
cd "$(dirname "$0")/.."
python "monocular-depth-estimation-cil_b200/build.py" >/dev/null
exec /usr/local/graft/bin/gpurun "$@"
