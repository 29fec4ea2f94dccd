#!/bin/bash
# quick iteration: parity tests then a short bench (no secondary / cpu legs)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 --no-cpu $BENCH_ARGS > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print("value %.3e  ms/step %.1f  e2e %.3e  roofline %.3f  kernel_ms %.1f"%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['kernel_ms']))
if 'secondary' in d: print("secondary: %.3e node-evals/s  %.1f ms"%(d['secondary']['value'], d['secondary']['ms_per_step']))
PY
