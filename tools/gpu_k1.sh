#!/bin/bash
# pointwise kernel iteration: GPU tests, then the secondary bench alone (+ optional ncu capture)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python bench.py --only-secondary --steps 3 > gpurun_out/k1.json 2> gpurun_out/k1.err; echo "k1 rc=$?"; tail -2 gpurun_out/k1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/k1.json').read().strip().splitlines()[-1])
print("C4: %.2f ms  %.3e live node-evals/s  wide(n_vars=8): %s"%(d['ms_per_step'], d['value'], d.get('wide_variant')))
PY
if [ "$1" == "prof" ]; then
  python bench.py --only-secondary --steps 2 > gpurun_out/plain3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_pointwise -s 6 -c 1 -f -o gpurun_out/prof_pointwise \
      python bench.py --only-secondary --steps 2 > gpurun_out/ncu_full2.log 2>&1
  echo "ncu pointwise rc=$?"
fi
