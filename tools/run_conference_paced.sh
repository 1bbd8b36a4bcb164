#!/bin/bash
# Live-call shape: N 720p streams paced at 30 fps on one GPU, nothing in flight; reports the
# encode -> packets' worth of NALs -> decode -> host latency per picture under that load.
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
for n in ${1:-8 30 60}; do
  /tmp/conference_bench /tmp/conf_720p.yuv 1280 720 30 $n 90 0 1 30 || true
done
