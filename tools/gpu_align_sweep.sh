#!/bin/bash
# SIMT MixerBlock backward: which CTA re-alignment points pay (MMX_MLP_ALIGN_MASK bit i = point i, bit 12 = token k-loop; decimal)
for m in ${MASKS:-3 5 9 17 33 65 129 257 73 72 8 64 75 77 89 105 201 329 4169}; do
  for B in ${BS:-4096}; do
    echo -n "mask=$m B=$B  "; env MMX_MLP_ALIGN_MASK=$m B=$B PDROP=0.1 ITERS=30 python tools/quick_bench.py | python -c "import json,sys; d=json.load(sys.stdin)['p0.1']; print('fwd %.1f us  bwd %.1f us' % (1e3*d['fwd_ms'], 1e3*d['bwd_ms']))"
  done
done
