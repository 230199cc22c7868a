#!/bin/bash
# build locally (so the shipped .so is never stale), then run a command on the GPU box
set -e
cd /root/repo
make -s -j8 -C /root/repo/dcfa-yolo_b200/csrc
make -s -C /root/repo/oracle
T=${GPU_TIMEOUT:-1200}
exec /usr/local/graft/bin/gpurun --timeout $T -- "$@"
