#!/bin/bash
# A/B timing of tuning builds on the notebook configurations (adaptive kernel)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "" $(ls variants/*.so); do
  MTGP_LIB=${v:+$PWD/$v} python bench.py --only-notebooks > gpurun_out/vnb.json 2> gpurun_out/vnb.err || tail -3 gpurun_out/vnb.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open('gpurun_out/vnb.json'))
print("%-30s "%(sys.argv[1] or "default") + "  ".join("%s %.2f"%(k,v['ms_per_evaluation']) for k,v in d.items()))
PY
done
