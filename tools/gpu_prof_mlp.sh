#!/bin/bash
# ncu --set full capture of the MixerBlock warp-variant kernels (one launch each) via tools/quick_bench.py
tag=${1:-r1c}
mkdir -p gpurun_out
export PDROP=${PDROP:-0.0} ITERS=2 WARM=1
python tools/quick_bench.py > gpurun_out/plain_prof_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:Warp -s 2 -c 3 -f -o gpurun_out/prof_mlp_$tag python tools/quick_bench.py > gpurun_out/ncu_prof_$tag.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_prof_$tag.log
