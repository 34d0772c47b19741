#!/bin/bash
tag=${1:-r1f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$tag.log
for wl in ${WLS:-k3}; do
  timeout 600 python bench.py --workload $wl --steps 20 > gpurun_out/bench_${wl}_$tag.json 2> gpurun_out/bench_${wl}_$tag.err; echo "bench $wl rc=$?"; cut -c1-250 gpurun_out/bench_${wl}_$tag.json; tail -2 gpurun_out/bench_${wl}_$tag.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${wl}_$tag.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"], d.get("cpu_baseline"))
PY
done
