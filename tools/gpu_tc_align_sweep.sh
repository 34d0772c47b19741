#!/bin/bash
# tensor-core MixerBlock backward: which CTA re-alignment points pay (MMX_TC_ALIGN_MASK, decimal; bit i = TC_ALIGN(i))
for m in ${MASKS:-32767 0 2 8 512 8192 16384 8704 514 522 128 640}; do
  for B in ${BS:-4096 16384}; do
    echo -n "mask=$m B=$B  "; env MMX_PRECISION=tf32 MMX_TC_ALIGN_MASK=$m B=$B PDROP=0.1 ITERS=30 python tools/quick_bench.py | python -c "import json,sys; d=json.load(sys.stdin)['p0.1']; print('fwd %.1f us  bwd %.1f us' % (1e3*d['fwd_ms'], 1e3*d['bwd_ms']))"
  done
done
