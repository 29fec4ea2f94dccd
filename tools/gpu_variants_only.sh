#!/bin/bash
# A/B timing of the tuning builds under variants/ (no tests); two interleaved rounds on the same box
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for round in 1 2; do
for v in "" $(ls variants/*.so); do
  MTGP_LIB=${v:+$PWD/$v} python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open('gpurun_out/v.json'))
print("%-28s value %.3e  ms/step %.1f  kernel_ms %.1f  sm_mhz %s"%(sys.argv[1] or "default", d['value'],d['ms_per_step'],d['roofline']['kernel_ms'], d['clocks'].get('sm_mhz')))
PY
done
done
