#!/bin/bash
# local wrapper: ALWAYS rebuild the in-tree libraries before snapshotting the repo to the GPU box (a stale .so travels silently)
set -e
cd "$(dirname "$0")/.."
make -C live_ekf_slam_b200/csrc -j8 2>&1 | grep -E "error|Error|warning: v" && exit 1
make -C oracle >/dev/null
exec /usr/local/graft/bin/gpurun "$@"
