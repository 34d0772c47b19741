#!/bin/bash
# GPU-box pass for the ConvMixer path: parity tests (all -m gpu) + a quick per-kernel timing of the K1 / K3 shapes.
tag=${1:-r1b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_$tag.log
timeout 300 python tools/conv_bench.py > gpurun_out/conv_bench_$tag.json 2> gpurun_out/conv_bench_$tag.err; echo "conv_bench rc=$?"; cat gpurun_out/conv_bench_$tag.json; tail -3 gpurun_out/conv_bench_$tag.err
