#!/bin/bash
# A/B of one env knob on the whole K2 training step (device-resident, FP32 mode): bash tools/gpu_step_ab.sh VAR=VALUE
for e in "X=1" "$@"; do echo -n "$e: "; env $e python bench.py --steps 50 --no-cpu-baseline --no-alt-precision 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('%.4f ms/step  %.0f seq/s' % (d['ms_per_step'], d['value']))"; done
