#!/bin/bash
# compute-sanitizer passes over the small GPU parity cases: memcheck (global /
# shared out-of-bounds, misaligned accesses) and racecheck (shared-memory hazards).  Run on the GPU box.
cd "$(dirname "$0")/.."
SEL='64-64 or 192-136 or 128-72 or 200-136 or 72-200 or 64-8 or 72-40 or 24-16 or committed_frames or corrupted_peer'
for tool in memcheck racecheck; do
  echo "==== $tool ===="
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_enc_gpu.py tests/test_dec_gpu.py tests/test_conv_gpu.py -x -q -k "$SEL" -p no:cacheprovider 2>&1 | grep -v "^$" | tail -25
  echo "exit: ${PIPESTATUS[0]}"
done
