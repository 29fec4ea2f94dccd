#!/bin/bash
# pointwise variants: secondary bench with the default build and every variants/*.so
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "" $(ls variants/*.so 2>/dev/null); do
  MTGP_LIB=${v:+$PWD/$v} python bench.py --only-secondary --steps 3 > gpurun_out/k1.json 2> gpurun_out/k1.err || tail -3 gpurun_out/k1.err
  python - "$v" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/k1.json').read().strip().splitlines()[-1])
print("%-30s C4 %.2f ms  wide %.2f ms"%(sys.argv[1] or "default", d['ms_per_step'], d['wide_variant']['ms_per_step']))
PY
done
