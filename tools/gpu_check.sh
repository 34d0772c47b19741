#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), then the ncu launch list of the same bench command.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tag]
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$tag.log
tail -5 gpurun_out/pytest_$tag.log
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; cat gpurun_out/bench_ref_$tag.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"
python bench.py --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32_$tag.json 2>> gpurun_out/bench_$tag.err; cat gpurun_out/bench_tf32_$tag.json
