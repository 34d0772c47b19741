#!/bin/bash
# quick GPU pass for the tensor-core MixerBlock kernels: block-level check vs the FP32 kernels + kernel timings
for cfg in "B=100 H=50" "B=301 H=32 CH=40 RR=2 ACT=gelu" "B=4096 H=50"; do echo "== $cfg"; env $cfg timeout 120 python tools/tc_check.py 2>&1 | tr "\n" " "; echo; done
env B=4096 PDROP=0.1 MMX_PRECISION=tf32 timeout 100 python tools/quick_bench.py
env B=4096 PDROP=0.0 MMX_PRECISION=tf32 timeout 100 python tools/quick_bench.py
env B=16384 PDROP=0.1 MMX_PRECISION=tf32 timeout 100 python tools/quick_bench.py
