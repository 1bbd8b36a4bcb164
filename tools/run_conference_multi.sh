#!/bin/bash
# Conference on the whole node: 720p30 encode+decode streams paced at 30 fps, sharded over N GPUs.
# usage: run_conference_multi.sh <n_gpus> "<streams list>"    (run under gpurun --gpus N)
set -e
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
N=${1:-1}
g++ -O2 -std=c++17 -Iinclude tools/conference_bench.cpp -o /tmp/conference_bench -Lkvazzup_b200 -lb200media -Wl,-rpath,"$PWD/kvazzup_b200" -lpthread
python - <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from kvazzup_b200 import synth
np.concatenate([synth.camera_i420(1280, 720, t) for t in range(30)]).tofile('/tmp/conf_720p.yuv')
PY
for n in ${2:-$((8*N)) $((30*N))}; do
  /tmp/conference_bench /tmp/conf_720p.yuv 1280 720 30 $n 90 0 1 30 $N || true
done
nproc
