#!/bin/bash
# tools/gpu/run.sh <script> [gpurun options]: rebuild every native artefact HERE (the in-tree .so files travel with the
# snapshot; a stale one silently measures old kernels), then run the script on the GPU box.
set -e
cd "$(dirname "$0")/../.."
python __graft_entry__.py > /dev/null
s="$1"; shift
n=$(basename "$s" .sh)
exec /usr/local/graft/bin/gpurun "$@" -- "bash $s > gpurun_out/$n.log 2>&1; tail -60 gpurun_out/$n.log"
