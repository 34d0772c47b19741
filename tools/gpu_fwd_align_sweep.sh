#!/bin/bash
# forward kernels: CTA re-alignment points (decimal masks)
for m in ${MASKS_SIMT:-0 1 15 2 4 8}; do
  for B in ${BS:-4096 16384}; do
    echo -n "SIMT fwd mask=$m B=$B  "; env MMX_MLP_ALIGN_MASK_FWD=$m B=$B PDROP=0.1 ITERS=30 python tools/quick_bench.py | python -c "import json,sys; d=json.load(sys.stdin)['p0.1']; print('fwd %.1f us  bwd %.1f us' % (1e3*d['fwd_ms'], 1e3*d['bwd_ms']))"
  done
done
for m in ${MASKS_TC:-0 1 127 8 64 9}; do
  for B in ${BS:-4096 16384}; do
    echo -n "TC fwd mask=$m B=$B  "; env MMX_PRECISION=tf32 MMX_TC_ALIGN_MASK_FWD=$m B=$B PDROP=0.1 ITERS=30 python tools/quick_bench.py | python -c "import json,sys; d=json.load(sys.stdin)['p0.1']; print('fwd %.1f us  bwd %.1f us' % (1e3*d['fwd_ms'], 1e3*d['bwd_ms']))"
  done
done
