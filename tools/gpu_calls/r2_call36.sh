#!/bin/bash
# Round 2, call 36: CUPTI in-situ per-kernel breakdown of the final C2 step (warm, two streams) + timeline
mkdir -p gpurun_out
timeout 600 python tools/warm_profile.py > gpurun_out/r2c36_warm_profile.txt 2>&1
echo "rc=$?"; head -60 gpurun_out/r2c36_warm_profile.txt
