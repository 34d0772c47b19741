#!/bin/bash
# GPU-box pass for the MotionMixer path: parity tests, per-kernel timings (v2 warp variant vs v1), bench line.
tag=${1:-r1c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$tag.log
echo "--- quick_bench warp variant"; timeout 120 python tools/quick_bench.py | tee gpurun_out/quick_$tag.json
echo "--- quick_bench generic (MMX_MLP_V1=1)"; MMX_MLP_V1=1 timeout 120 python tools/quick_bench.py | tee gpurun_out/quick_v1_$tag.json
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
