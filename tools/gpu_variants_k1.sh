#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "" $(ls variants/*.so); do
  MTGP_LIB=${v:+$PWD/$v} python bench.py --only-secondary --steps 3 > gpurun_out/v.json 2> gpurun_out/v.err || tail -3 gpurun_out/v.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open('gpurun_out/v.json'))
print("%-28s %.3e node-evals/s  %.2f ms"%(sys.argv[1] or "default", d['value'],d['ms_per_step']))
PY
done
