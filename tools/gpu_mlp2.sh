#!/bin/bash
tag=${1:-r1e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
echo "--- quick_bench warp variant"; timeout 120 python tools/quick_bench.py | tee gpurun_out/quick_$tag.json
echo "--- quick_bench warp variant, fwd occ2"; MMX_MLP_FWD_OCC=2 timeout 120 python tools/quick_bench.py | tee gpurun_out/quick_occ2_$tag.json
for w in 7 6; do echo "--- warps $w"; MMX_MLP_WARPS=$w timeout 120 python tools/quick_bench.py; done
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_$tag.json
