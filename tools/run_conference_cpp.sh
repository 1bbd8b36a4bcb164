#!/bin/bash
# Builds tools/conference_bench.cpp against the in-tree library and runs the 720p conference sweep.
set -e
cd "$(dirname "$0")/.."
export CUDA_DEVICE_MAX_CONNECTIONS=32
g++ -O2 -std=c++17 -Iinclude tools/conference_bench.cpp -o /tmp/conference_bench -Lkvazzup_b200 -lb200media -Wl,-rpath,"$PWD/kvazzup_b200" -lpthread
python - <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from kvazzup_b200 import synth
np.concatenate([synth.camera_i420(1280, 720, t) for t in range(30)]).tofile('/tmp/conf_720p.yuv')
PY
for cfg in "8 90 0 1" "32 90 0 1" "64 60 0 1" "96 45 0 1" "32 90 0 4" "32 90 1 1" "96 60 1 1"; do
  /tmp/conference_bench /tmp/conf_720p.yuv 1280 720 30 $cfg || true
done
