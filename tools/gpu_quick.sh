#!/bin/bash
# perf iteration: bench the default build and every variants/*.so (no tests), then an ncu capture of the default
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "" $(ls variants/*.so 2>/dev/null); do
  MTGP_LIB=${v:+$PWD/$v} python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open('gpurun_out/v.json'))
print("%-28s value %.3e  ms/step %.1f  kernel_ms %.1f"%(sys.argv[1] or "default", d['value'],d['ms_per_step'],d['roofline']['kernel_ms']))
PY
done
[ "$1" == "noprof" ] || bash tools/gpu_prof.sh
