#!/bin/bash
# Round 2, call 3: per-kernel timeline of one warm step (graph replay) for the overlap analysis.
mkdir -p gpurun_out
timeout 300 python tools/step_trace.py r2c3 1 2>&1 | tail -3
timeout 300 python tools/step_trace.py r2c3_eager 0 2>&1 | tail -3
